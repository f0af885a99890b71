// fsweep_inst.cu — explicit instantiations of the sweep kernels for one group size.
// Compiled once per G with -DFSWEEP_G=<1|2|4|8|16|32> so the six widths build in parallel.
#include "fsweep_kernels.cuh"

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
  k<<<cfg.grid, BLOCK, cfg.smem, cfg.stream>>>(P, A);
  return cudaGetLastError();
}

template <typename T, int G, int CC>
static cudaError_t launch_bwd_t(const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A) {
  auto k = fsweep_bwd_kernel<T, G, CC>;
  if (cfg.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
  }
  k<<<cfg.grid, BLOCK, cfg.smem, cfg.stream>>>(P, A);
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

#define FSWEEP_DISPATCH(fn, ...)                                         \
  if (dtype == FSWEEP_C64) {                                             \
    if (cc == 1) return fn<float, FSWEEP_G, 1>(__VA_ARGS__);             \
    return fn<float, FSWEEP_G, 4>(__VA_ARGS__);                          \
  } else {                                                               \
    if (cc == 1) return fn<double, FSWEEP_G, 1>(__VA_ARGS__);            \
    return fn<double, FSWEEP_G, 4>(__VA_ARGS__);                         \
  }

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

}  // namespace fsweep
