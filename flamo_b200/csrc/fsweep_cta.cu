// fsweep_cta.cu — instantiations, launch thunk and occupancy query of the CTA-per-bin kernels (fsweep_cta.cuh):
// SIMT warp-pipeline elimination (256 threads) and the tensor-core elimination (128 threads, fsweep_tc.cuh).
#include "fsweep_cta.cuh"

namespace fsweep {

template <bool BWD, bool TC>
static cudaError_t configure() {
  static bool done = false;  // per instantiation; benign if two host threads race (same values)
  if (done) return cudaSuccess;
  auto k = fsweep_cta_kernel<BWD, TC>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem_bytes(BWD, TC));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

template <bool BWD, bool TC>
static cudaError_t launch(int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G) {
  cudaError_t e = configure<BWD, TC>();
  if (e != cudaSuccess) return e;
  fsweep_cta_kernel<BWD, TC><<<grid, TC ? tc::T : CTA_T, cta_smem_bytes(BWD, TC), st>>>(P, L, A, G);
  return cudaGetLastError();
}

template <bool BWD, bool TC>
static cudaError_t occupancy(int* blocks_per_sm) {
  cudaError_t e = configure<BWD, TC>();
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_cta_kernel<BWD, TC>, TC ? tc::T : CTA_T,
                                                       cta_smem_bytes(BWD, TC));
}

cudaError_t launch_cta(bool bwd, bool tc, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                       int G) {
  if (tc) return bwd ? launch<true, true>(grid, st, P, L, A, G) : launch<false, true>(grid, st, P, L, A, G);
  return bwd ? launch<true, false>(grid, st, P, L, A, G) : launch<false, false>(grid, st, P, L, A, G);
}

cudaError_t occupancy_cta(bool bwd, bool tc, int* blocks_per_sm) {
  if (tc) return bwd ? occupancy<true, true>(blocks_per_sm) : occupancy<false, true>(blocks_per_sm);
  return bwd ? occupancy<true, false>(blocks_per_sm) : occupancy<false, false>(blocks_per_sm);
}

}  // namespace fsweep
