// fsweep_tpr.cu — instantiations and the launch thunk of the register-matrix thread-per-bin kernels (fsweep_tpr.cuh).
#include "fsweep_tpr.cuh"

namespace fsweep {

template <int NP, bool BWD>
static cudaError_t tpr_t(int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G) {
  auto k = fsweep_tpr_kernel<NP, BWD>;
  const size_t smem = TprSmem<NP>::bytes;
  static bool configured = false;  // per instantiation; benign if two host threads race (same value)
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return launch_pdl(k, dim3(grid), dim3(TPR_BLOCK), smem, st, P, L, A, G);
}

cudaError_t launch_tpr(int np, bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                       int G) {
  (void)np;  // the 8-wide instantiation serves every width <= 8 (rows beyond the live width are identity rows)
  return bwd ? tpr_t<8, true>(grid, st, P, L, A, G) : tpr_t<8, false>(grid, st, P, L, A, G);
}

}  // namespace fsweep
