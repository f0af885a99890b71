// fsweep_tpc.cu — instantiations and the launch thunk of the compact thread-per-bin kernels (fsweep_tpc.cuh).
#include "fsweep_tpc.cuh"

namespace fsweep {

template <int NP, bool BWD>
static cudaError_t tpc_t(int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G) {
  auto k = fsweep_tpc_kernel<NP, BWD>;
  const size_t smem = tpc_smem_bytes(NP);
  static bool configured = false;  // per instantiation; benign if two host threads race (same value)
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k<<<grid, TPC_BLOCK, smem, st>>>(P, L, A, G);
  return cudaGetLastError();
}

cudaError_t launch_tpc(int np, bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                       int G) {
  if (np == 4) return bwd ? tpc_t<4, true>(grid, st, P, L, A, G) : tpc_t<4, false>(grid, st, P, L, A, G);
  return bwd ? tpc_t<8, true>(grid, st, P, L, A, G) : tpc_t<8, false>(grid, st, P, L, A, G);
}

}  // namespace fsweep
