// fsweep_cta.cu — instantiations, launch thunk and occupancy query of the CTA-per-bin kernels (fsweep_cta.cuh).
#include "fsweep_cta.cuh"

namespace fsweep {

template <bool BWD>
static cudaError_t configure() {
  static bool done = false;  // per instantiation; benign if two host threads race (same values)
  if (done) return cudaSuccess;
  auto k = fsweep_cta_kernel<BWD>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cta_smem_bytes(BWD));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

cudaError_t launch_cta(bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G) {
  cudaError_t e = bwd ? configure<true>() : configure<false>();
  if (e != cudaSuccess) return e;
  if (bwd)
    fsweep_cta_kernel<true><<<grid, CTA_T, cta_smem_bytes(true), st>>>(P, L, A, G);
  else
    fsweep_cta_kernel<false><<<grid, CTA_T, cta_smem_bytes(false), st>>>(P, L, A, G);
  return cudaGetLastError();
}

cudaError_t occupancy_cta(bool bwd, int* blocks_per_sm) {
  cudaError_t e = bwd ? configure<true>() : configure<false>();
  if (e != cudaSuccess) return e;
  if (bwd) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_cta_kernel<true>, CTA_T, cta_smem_bytes(true));
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_cta_kernel<false>, CTA_T, cta_smem_bytes(false));
}

}  // namespace fsweep
