// fsweep_stream.cu — instantiations, launch thunk and occupancy query of the streaming table kernels (fsweep_stream.cuh).
#include "fsweep_stream.cuh"

namespace fsweep {

template <bool BWD, bool TMA>
static cudaError_t configure(size_t smem) {
  static size_t configured = 0;  // per instantiation: largest dynamic shared memory size opted into so far
  if (smem <= configured) return cudaSuccess;
  cudaError_t e =
      cudaFuncSetAttribute(fsweep_stream_kernel<BWD, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  configured = smem;
  return cudaSuccess;
}

template <bool BWD, bool TMA>
static cudaError_t launch(int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S, const SweepArgs& A,
                          int G) {
  cudaError_t e = configure<BWD, TMA>(smem);
  if (e != cudaSuccess) return e;
  fsweep_stream_kernel<BWD, TMA><<<grid, S.threads, smem, st>>>(P, S, A, G);
  return cudaGetLastError();
}

template <bool BWD, bool TMA>
static cudaError_t occupancy(int threads, size_t smem, int* blocks_per_sm) {
  cudaError_t e = configure<BWD, TMA>(smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_stream_kernel<BWD, TMA>, threads, smem);
}

cudaError_t launch_stream(bool bwd, bool tma, int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                          const SweepArgs& A, int G) {
  if (tma) return bwd ? launch<true, true>(grid, smem, st, P, S, A, G) : launch<false, true>(grid, smem, st, P, S, A, G);
  return bwd ? launch<true, false>(grid, smem, st, P, S, A, G) : launch<false, false>(grid, smem, st, P, S, A, G);
}

cudaError_t occupancy_stream(bool bwd, bool tma, int threads, size_t smem, int* blocks_per_sm) {
  if (tma) return bwd ? occupancy<true, true>(threads, smem, blocks_per_sm) : occupancy<false, true>(threads, smem, blocks_per_sm);
  return bwd ? occupancy<true, false>(threads, smem, blocks_per_sm) : occupancy<false, false>(threads, smem, blocks_per_sm);
}

}  // namespace fsweep
