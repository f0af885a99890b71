// fsweep_stream.cu — instantiations, launch thunk and occupancy query of the streaming table kernels (fsweep_stream.cuh).
#include "fsweep_stream.cuh"

namespace fsweep {

template <bool BWD>
static cudaError_t configure(size_t smem) {
  static size_t configured = 0;  // per instantiation: largest dynamic shared memory size opted into so far
  if (smem <= configured) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fsweep_stream_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  configured = smem;
  return cudaSuccess;
}

cudaError_t launch_stream(bool bwd, int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                          const SweepArgs& A, int G) {
  cudaError_t e = bwd ? configure<true>(smem) : configure<false>(smem);
  if (e != cudaSuccess) return e;
  if (bwd)
    fsweep_stream_kernel<true><<<grid, S.threads, smem, st>>>(P, S, A, G);
  else
    fsweep_stream_kernel<false><<<grid, S.threads, smem, st>>>(P, S, A, G);
  return cudaGetLastError();
}

cudaError_t occupancy_stream(bool bwd, int threads, size_t smem, int* blocks_per_sm) {
  cudaError_t e = bwd ? configure<true>(smem) : configure<false>(smem);
  if (e != cudaSuccess) return e;
  if (bwd) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_stream_kernel<true>, threads, smem);
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_stream_kernel<false>, threads, smem);
}

}  // namespace fsweep
