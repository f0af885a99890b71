// fsweep_inst.cu — explicit instantiations of the sweep kernels for one group size.
// Compiled once per G with -DFSWEEP_G=<1|2|4|8|16|32> so the six widths build in parallel.
#include "fsweep_loop.cuh"

#ifndef FSWEEP_G
#error "compile with -DFSWEEP_G=<group size>"
#endif

namespace fsweep {

template <typename T, int G, int CC>
static cudaError_t launch_fwd_t(const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A) {
  auto k = fsweep_fwd_kernel<T, G, CC>;
  if (cfg.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
  }
  k<<<dim3((unsigned)cfg.grid, (unsigned)cfg.items), BLOCK, cfg.smem, cfg.stream>>>(P, A);
  return cudaGetLastError();
}

template <typename T, int G, int CC>
static cudaError_t launch_bwd_t(const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A) {
  auto k = fsweep_bwd_kernel<T, G, CC>;
  if (cfg.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
  }
  k<<<dim3((unsigned)cfg.grid, (unsigned)cfg.items), BLOCK, cfg.smem, cfg.stream>>>(P, A);
  return cudaGetLastError();
}

template <typename T, int G, int CC>
static cudaError_t occ_t(bool bwd, size_t smem, int* n) {
  if (bwd) {
    auto k = fsweep_bwd_kernel<T, G, CC>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, k, BLOCK, smem);
  }
  auto k = fsweep_fwd_kernel<T, G, CC>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, k, BLOCK, smem);
}

#if FSWEEP_G <= 32
#define FSWEEP_DISPATCH(fn, ...)                                         \
  if (dtype == FSWEEP_C64) {                                             \
    if (cc == 1) return fn<float, FSWEEP_G, 1>(__VA_ARGS__);             \
    return fn<float, FSWEEP_G, 4>(__VA_ARGS__);                          \
  } else {                                                               \
    if (cc == 1) return fn<double, FSWEEP_G, 1>(__VA_ARGS__);            \
    return fn<double, FSWEEP_G, 4>(__VA_ARGS__);                         \
  }
#else  // two warps per bin: float32 only (a row of 64 complex128 does not fit the register file)
#define FSWEEP_DISPATCH(fn, ...)                                         \
  if (dtype == FSWEEP_C64) {                                             \
    if (cc == 1) return fn<float, FSWEEP_G, 1>(__VA_ARGS__);             \
    return fn<float, FSWEEP_G, 4>(__VA_ARGS__);                          \
  }                                                                      \
  return cudaErrorNotSupported;
#endif

template <>
cudaError_t launch_fwd<FSWEEP_G>(int dtype, int cc, const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A) {
  FSWEEP_DISPATCH(launch_fwd_t, cfg, P, A)
}
template <>
cudaError_t launch_bwd<FSWEEP_G>(int dtype, int cc, const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A) {
  FSWEEP_DISPATCH(launch_bwd_t, cfg, P, A)
}
template <>
cudaError_t occupancy<FSWEEP_G>(int dtype, int cc, bool bwd, size_t smem, int* blocks_per_sm) {
  FSWEEP_DISPATCH(occ_t, bwd, smem, blocks_per_sm)
}

// ---- pattern-specialised FDN-loop kernels (float for every G <= 32; double up to G = 16: registers)
#if FSWEEP_G <= 32
constexpr bool kLoopDouble = FSWEEP_G <= 16;

template <typename T, int G, bool BWD>
static cudaError_t launch_loop_t(const LaunchCfg& cfg, const ProgK& P, const LoopInfo& L, const SweepArgs& A) {
  auto k = BWD ? fsweep_loop_bwd_kernel<T, G> : fsweep_loop_fwd_kernel<T, G>;
  if (cfg.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
  }
  k<<<cfg.grid, BLOCK, cfg.smem, cfg.stream>>>(P, L, A);
  return cudaGetLastError();
}

template <typename T, int G, bool BWD>
static cudaError_t occ_loop_t(size_t smem, int* n) {
  auto k = BWD ? fsweep_loop_bwd_kernel<T, G> : fsweep_loop_fwd_kernel<T, G>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, k, BLOCK, smem);
}

template <>
cudaError_t launch_loop_fwd<FSWEEP_G>(int dtype, const LaunchCfg& cfg, const ProgK& P, const LoopInfo& L,
                                      const SweepArgs& A) {
  if (dtype == FSWEEP_C64) return launch_loop_t<float, FSWEEP_G, false>(cfg, P, L, A);
  if constexpr (kLoopDouble) return launch_loop_t<double, FSWEEP_G, false>(cfg, P, L, A);
  return cudaErrorNotSupported;
}
template <>
cudaError_t launch_loop_bwd<FSWEEP_G>(int dtype, const LaunchCfg& cfg, const ProgK& P, const LoopInfo& L,
                                      const SweepArgs& A) {
  if (dtype == FSWEEP_C64) return launch_loop_t<float, FSWEEP_G, true>(cfg, P, L, A);
  if constexpr (kLoopDouble) return launch_loop_t<double, FSWEEP_G, true>(cfg, P, L, A);
  return cudaErrorNotSupported;
}
template <>
cudaError_t occupancy_loop<FSWEEP_G>(int dtype, bool bwd, size_t smem, int* n) {
  if (dtype == FSWEEP_C64) return bwd ? occ_loop_t<float, FSWEEP_G, true>(smem, n) : occ_loop_t<float, FSWEEP_G, false>(smem, n);
  if constexpr (kLoopDouble)
    return bwd ? occ_loop_t<double, FSWEEP_G, true>(smem, n) : occ_loop_t<double, FSWEEP_G, false>(smem, n);
  return cudaErrorNotSupported;
}
#else
template <>
cudaError_t launch_loop_fwd<FSWEEP_G>(int, const LaunchCfg&, const ProgK&, const LoopInfo&, const SweepArgs&) {
  return cudaErrorNotSupported;
}
template <>
cudaError_t launch_loop_bwd<FSWEEP_G>(int, const LaunchCfg&, const ProgK&, const LoopInfo&, const SweepArgs&) {
  return cudaErrorNotSupported;
}
template <>
cudaError_t occupancy_loop<FSWEEP_G>(int, bool, size_t, int*) {
  return cudaErrorNotSupported;
}
#endif

}  // namespace fsweep
