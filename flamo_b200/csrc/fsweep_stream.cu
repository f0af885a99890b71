// fsweep_stream.cu — instantiations, launch thunk and occupancy query of the streaming table kernels (fsweep_stream.cuh).
#include "fsweep_stream.cuh"
#include "fsweep_streamr.cuh"
#include "fsweep_streamw.cuh"

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

template <int W, bool TMA>
static cudaError_t configure_r(size_t smem) {
  static size_t configured = 0;
  if (smem <= configured) return cudaSuccess;
  cudaError_t e =
      cudaFuncSetAttribute(fsweep_streamr_kernel<W, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  configured = smem;
  return cudaSuccess;
}

template <int W, bool TMA>
static cudaError_t launch_r(int grid, int threads, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                            const StreamRInfo& R, const SweepArgs& A) {
  cudaError_t e = configure_r<W, TMA>(smem);
  if (e != cudaSuccess) return e;
  fsweep_streamr_kernel<W, TMA><<<grid, threads, smem, st>>>(P, S, R, A);
  return cudaGetLastError();
}

template <int W, bool TMA>
static cudaError_t occupancy_r(int threads, size_t smem, int* blocks_per_sm) {
  cudaError_t e = configure_r<W, TMA>(smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_streamr_kernel<W, TMA>, threads, smem);
}

template <bool TMA>
static cudaError_t launch_rw(int w, int grid, int threads, size_t smem, cudaStream_t st, const ProgK& P,
                             const StreamInfo& S, const StreamRInfo& R, const SweepArgs& A) {
  if (w <= 4) return launch_r<4, TMA>(grid, threads, smem, st, P, S, R, A);
  if (w <= 8) return launch_r<8, TMA>(grid, threads, smem, st, P, S, R, A);
  return launch_r<16, TMA>(grid, threads, smem, st, P, S, R, A);
}

cudaError_t launch_streamr(int w, bool tma, int grid, int threads, size_t smem, cudaStream_t st, const ProgK& P,
                           const StreamInfo& S, const StreamRInfo& R, const SweepArgs& A) {
  return tma ? launch_rw<true>(w, grid, threads, smem, st, P, S, R, A)
             : launch_rw<false>(w, grid, threads, smem, st, P, S, R, A);
}

cudaError_t occupancy_streamr(int w, bool tma, int threads, size_t smem, int* blocks_per_sm) {
  if (tma) {
    if (w <= 4) return occupancy_r<4, true>(threads, smem, blocks_per_sm);
    if (w <= 8) return occupancy_r<8, true>(threads, smem, blocks_per_sm);
    return occupancy_r<16, true>(threads, smem, blocks_per_sm);
  }
  if (w <= 4) return occupancy_r<4, false>(threads, smem, blocks_per_sm);
  if (w <= 8) return occupancy_r<8, false>(threads, smem, blocks_per_sm);
  return occupancy_r<16, false>(threads, smem, blocks_per_sm);
}

template <int QC, bool TMA>
static cudaError_t configure_w(size_t smem) {
  static size_t configured = 0;
  if (smem <= configured) return cudaSuccess;
  cudaError_t e =
      cudaFuncSetAttribute(fsweep_streamw_kernel<QC, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  configured = smem;
  return cudaSuccess;
}

template <int QC, bool TMA>
static cudaError_t launch_w(int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                            const StreamRInfo& R, const SweepArgs& A) {
  cudaError_t e = configure_w<QC, TMA>(smem);
  if (e != cudaSuccess) return e;
  fsweep_streamw_kernel<QC, TMA><<<grid, SWARP_THREADS, smem, st>>>(P, S, R, A);
  return cudaGetLastError();
}

template <int QC, bool TMA>
static cudaError_t occupancy_w(size_t smem, int* blocks_per_sm) {
  cudaError_t e = configure_w<QC, TMA>(smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fsweep_streamw_kernel<QC, TMA>, SWARP_THREADS,
                                                       smem);
}

#define FSWEEP_QC_SWITCH(QCV, TMAV, CALL)                          \
  switch (QCV) {                                                   \
    case 1: return TMAV ? CALL(1, true) : CALL(1, false);          \
    case 2: return TMAV ? CALL(2, true) : CALL(2, false);          \
    case 4: return TMAV ? CALL(4, true) : CALL(4, false);          \
    case 8: return TMAV ? CALL(8, true) : CALL(8, false);          \
    default: return TMAV ? CALL(16, true) : CALL(16, false);       \
  }

cudaError_t launch_streamw(int qc, bool tma, int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                           const StreamRInfo& R, const SweepArgs& A) {
#define FSWEEP_W_LAUNCH(Q, TM) launch_w<Q, TM>(grid, smem, st, P, S, R, A)
  FSWEEP_QC_SWITCH(qc, tma, FSWEEP_W_LAUNCH)
#undef FSWEEP_W_LAUNCH
}

cudaError_t occupancy_streamw(int qc, bool tma, size_t smem, int* blocks_per_sm) {
#define FSWEEP_W_OCC(Q, TM) occupancy_w<Q, TM>(smem, blocks_per_sm)
  FSWEEP_QC_SWITCH(qc, tma, FSWEEP_W_OCC)
#undef FSWEEP_W_OCC
}

}  // namespace fsweep
