// fsweep_cta.cu — instantiations, launch thunk and occupancy query of the CTA-per-bin kernels (fsweep_cta.cuh):
// SIMT warp-pipeline elimination (256 threads; float32 and float64) and the tensor-core elimination (128 threads,
// float32, fsweep_tc.cuh).
#include "fsweep_cta.cuh"

namespace fsweep {

template <typename T, bool BWD, bool TC>
static cudaError_t configure() {
  static bool done = false;  // per instantiation; benign if two host threads race (same values)
  if (done) return cudaSuccess;
  auto k = fsweep_cta_kernel<T, BWD, TC>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem_bytes(BWD, TC, sizeof(T)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

template <typename T, bool BWD, bool TC>
static cudaError_t launch(int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G) {
  cudaError_t e = configure<T, BWD, TC>();
  if (e != cudaSuccess) return e;
  fsweep_cta_kernel<T, BWD, TC><<<grid, TC ? tc::T : CTA_T, cta_smem_bytes(BWD, TC, sizeof(T)), st>>>(P, L, A, G);
  return cudaGetLastError();
}

template <typename T, bool BWD, bool TC>
static cudaError_t occupancy(int* blocks_per_sm) {
  cudaError_t e = configure<T, BWD, TC>();
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_cta_kernel<T, BWD, TC>, TC ? tc::T : CTA_T,
                                                       cta_smem_bytes(BWD, TC, sizeof(T)));
}

cudaError_t launch_cta(int dtype, bool bwd, bool tc, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L,
                       const SweepArgs& A, int G) {
  if (dtype == FSWEEP_C128)
    return bwd ? launch<double, true, false>(grid, st, P, L, A, G) : launch<double, false, false>(grid, st, P, L, A, G);
  if (tc) return bwd ? launch<float, true, true>(grid, st, P, L, A, G) : launch<float, false, true>(grid, st, P, L, A, G);
  return bwd ? launch<float, true, false>(grid, st, P, L, A, G) : launch<float, false, false>(grid, st, P, L, A, G);
}

cudaError_t occupancy_cta(int dtype, bool bwd, bool tc, int* blocks_per_sm) {
  if (dtype == FSWEEP_C128)
    return bwd ? occupancy<double, true, false>(blocks_per_sm) : occupancy<double, false, false>(blocks_per_sm);
  if (tc) return bwd ? occupancy<float, true, true>(blocks_per_sm) : occupancy<float, false, true>(blocks_per_sm);
  return bwd ? occupancy<float, true, false>(blocks_per_sm) : occupancy<float, false, false>(blocks_per_sm);
}

}  // namespace fsweep
