// fsweep_tpb.cu — instantiations and launch thunks of the thread-per-bin kernels (fsweep_tpb.cuh).
#include "fsweep_tpb.cuh"

namespace fsweep {

cudaError_t launch_tpb_fwd(int np, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A) {
  if (np == 4)
    fsweep_tpb_fwd_kernel<4><<<grid, TPB_BLOCK, 0, st>>>(P, L, A);
  else
    fsweep_tpb_fwd_kernel<8><<<grid, TPB_BLOCK, 0, st>>>(P, L, A);
  return cudaGetLastError();
}

template <int NP>
static cudaError_t bwd_t(int grid, size_t smem, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                         int G) {
  auto k = fsweep_tpb_bwd_kernel<NP>;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, TPB_BLOCK, smem, st>>>(P, L, A, G);
  return cudaGetLastError();
}

cudaError_t launch_tpb_bwd(int np, int grid, size_t smem, cudaStream_t st, const ProgK& P, const LoopInfo& L,
                           const SweepArgs& A, int G) {
  return np == 4 ? bwd_t<4>(grid, smem, st, P, L, A, G) : bwd_t<8>(grid, smem, st, P, L, A, G);
}

}  // namespace fsweep
