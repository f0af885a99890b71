// fsweep_streamw.cuh — streaming sweep, FORWARD, one WARP per bin.
//
// The thread-per-(bin, column) kernels walk the whole op chain privately: fully unrolled register code is 30+ KB per
// tile pass (r02t capture of fsweep_streamr_kernel: 4 of every 6 stall cycles are instruction fetch) and rolled code
// needs the state in shared memory (fsweep_stream_kernel: half of the shared memory holds state, 13 warps per SM).
// Here the 32 lanes of a warp share ONE bin: lane = (row r, column c), c fastest, RP = 32 / columns rows per pass, and
// the signal of the bin is DISTRIBUTED over the lanes — row m of column c lives in slot m / RP of lane (m % RP, c):
//   * a dense op is a rolled loop over its inputs: two shuffles bring v[n][c] from the lane that holds it, one LDS.64
//     brings H[m][n] (the lanes of a pass read consecutive rows: no bank conflicts), four FFMAs accumulate;
//   * diagonal ops are one multiply per slot;
//   * nothing is unrolled beyond the slots: the whole kernel is a few hundred instructions (it lives in the L0
//     instruction cache), ~40 registers per thread, and shared memory holds only table tiles — two per block, filled
//     one tile ahead by the TMA bulk-copy engine (cp.async.bulk + mbarrier; plain cp.async granules when a bin shard
//     breaks the 16-byte alignment);
//   * x is read and y written with the lanes of a bin on consecutive addresses.
// Same programs as fsweep_stream_kernel (TABLE / PTABLE / GAIN / PGAIN, widths <= 16, batch * cols a power of two <= 16),
// forward only.
#pragma once
#include "fsweep_streamr.cuh"

namespace fsweep {

constexpr int SWARP_THREADS = 128;

template <int QC, bool TMA>
__global__ void __launch_bounds__(SWARP_THREADS) fsweep_streamw_kernel(const __grid_constant__ ProgK P,
                                                                     const __grid_constant__ StreamInfo S,
                                                                     const __grid_constant__ StreamRInfo R,
                                                                     const SweepArgs A) {
  constexpr int RP = 32 / QC;              // rows per pass
  constexpr int NS = (SW + RP - 1) / RP;   // slots per lane
  extern __shared__ __align__(128) unsigned char wsm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, n_warps = T >> 5;
  const int c = lane & (QC - 1), r = lane / QC;
  const int tb = S.tb;
  const long long n_tiles = (A.n_bins + tb - 1) / tb;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  const int in_ch = P.in_ch, out_ch = P.out_ch, n_ops = S.n_ops;
  const int b = (A.cols == 1) ? c : c / A.cols, cc = c - b * A.cols;
  const int tile_units = R.tile_units;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(wsm + (size_t)(TMA ? 2 : 1) * tile_units * 8);

  auto is_full = [&](long long tile) { return (tile + 1) * tb <= A.n_bins; };
  auto issue_tma = [&](long long tile, int stage) {  // one thread
    if (tile < n_tiles && is_full(tile)) {
      float2* dst = reinterpret_cast<float2*>(wsm) + (size_t)stage * tile_units;
      mbar_expect_tx(sBar + stage, (uint32_t)tile_units * 8u);
      for (int i = 0; i < n_ops; ++i) {
        if (R.pad_units[i] == 0) continue;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                   (size_t)(A.bin_begin + tile * tb) * S.row_bytes[i];
        tma_load_1d(dst + R.pad_off[i], src, (uint32_t)(tb * S.row_bytes[i]), sBar + stage);
      }
    }
  };
  if constexpr (TMA) {
    if (tid == 0) {
      mbar_init(sBar, 1);
      mbar_init(sBar + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) issue_tma(blockIdx.x, 0);
  }
  int stage = 0;
  uint32_t phases = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long b0 = tile * tb;
    const float2* sTab = reinterpret_cast<const float2*>(wsm) + (size_t)stage * tile_units;
    const int nb = (int)min((long long)tb, A.n_bins - b0);
    bool plain = !TMA;
    if constexpr (TMA) {
      if (tid == 0) issue_tma(tile + gridDim.x, stage ^ 1);  // (released by the barrier that closed the last iteration)
      plain = !is_full(tile);
    }
    if (plain) {
      float2* dstage = reinterpret_cast<float2*>(wsm) + (size_t)stage * tile_units;
      for (int i = 0; i < n_ops; ++i) {
        if (R.pad_units[i] == 0) continue;
        const float2* src = reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                                            (size_t)(A.bin_begin + b0) * S.row_bytes[i]);
        float2* dst = dstage + R.pad_off[i];
        const int n8 = nb * R.pad_units[i];
        for (int e = tid; e < n8; e += T) cp_async8(dst + e, src + e);
      }
      __pipeline_commit();
      __pipeline_wait_prior(0);
      __syncthreads();
    } else {
      mbar_wait(sBar + stage, (phases >> stage) & 1u);
      phases ^= 1u << stage;
    }

    for (int bi = warp; bi < nb; bi += n_warps) {
      const long long bl = b0 + bi;
      float2 s[NS];
#pragma unroll
      for (int p = 0; p < NS; ++p) {
        const int n = p * RP + r;
        s[p] = f2(0.f, 0.f);
        if (n < in_ch) {
          const cx<float> t = ld_cx(x + (size_t)b * A.xbs + ((size_t)bl * in_ch + n) * A.cols + cc);
          s[p] = f2(t.x, t.y);
        }
      }
      for (int i = 0; i < n_ops; ++i) {
        const OpK& op = P.ops[i];
        const int n_in = op.n_in, n_out = op.n_out;
        if (op.kind == FSWEEP_OP_TABLE || op.kind == FSWEEP_OP_GAIN) {
          const bool real = op.kind == FSWEEP_OP_GAIN;
          const float2* H = sTab + R.pad_off[i] + bi * R.pad_units[i];
          const float* Wg = reinterpret_cast<const float*>(op.coef);
          float2 o[NS];
#pragma unroll
          for (int p = 0; p < NS; ++p) {
            o[p] = f2(0.f, 0.f);
            if (p * RP < n_out) {  // (uniform)
              const int m = min(p * RP + r, n_out - 1);  // lanes beyond the last row repeat it (their result is unused)
              float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
#pragma unroll
              for (int sp = 0; sp < NS; ++sp) {
                const int n0 = sp * RP;
                if (n0 < n_in) {  // (uniform)
                  const int cnt = min(RP, n_in - n0);
                  if (real) {
                    const float* w = Wg + m * n_in + n0;
#pragma unroll 4
                    for (int nn = 0; nn < cnt; ++nn) {
                      const float vx = __shfl_sync(FULL, s[sp].x, nn * QC + c);
                      const float vy = __shfl_sync(FULL, s[sp].y, nn * QC + c);
                      const float wv = __ldg(w + nn);
                      ax = fmaf(wv, vx, ax);
                      ay = fmaf(wv, vy, ay);
                    }
                  } else {
                    const float2* h = H + m * n_in + n0;
#pragma unroll 4
                    for (int nn = 0; nn < cnt; ++nn) {
                      const float vx = __shfl_sync(FULL, s[sp].x, nn * QC + c);
                      const float vy = __shfl_sync(FULL, s[sp].y, nn * QC + c);
                      const float2 hh = h[nn];
                      ax = fmaf(hh.x, vx, ax);
                      bx = fmaf(-hh.y, vy, bx);
                      ay = fmaf(hh.x, vy, ay);
                      by = fmaf(hh.y, vx, by);
                    }
                  }
                }
              }
              o[p] = f2(ax + bx, ay + by);
            }
          }
#pragma unroll
          for (int p = 0; p < NS; ++p) s[p] = o[p];
        } else if (op.kind == FSWEEP_OP_PTABLE) {
          const float2* H = sTab + R.pad_off[i] + bi * R.pad_units[i];
#pragma unroll
          for (int p = 0; p < NS; ++p) {
            const int m = p * RP + r;
            if (m < n_out) {
              const float2 h = H[m], v = s[p];
              s[p] = f2(h.x * v.x - h.y * v.y, h.x * v.y + h.y * v.x);
            }
          }
        } else {  // PGAIN
          const float* Wg = reinterpret_cast<const float*>(op.coef);
#pragma unroll
          for (int p = 0; p < NS; ++p) {
            const int m = p * RP + r;
            if (m < n_out) {
              const float w = __ldg(Wg + m);
              s[p] = f2(w * s[p].x, w * s[p].y);
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < NS; ++p) {
        const int m = p * RP + r;
        if (m < out_ch) {
          const size_t off = (size_t)b * A.ybs + ((size_t)bl * out_ch + m) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS)
            reinterpret_cast<float*>(A.y)[off] = abs_t(s[p].x, s[p].y);
          else
            st_cx(reinterpret_cast<cx<float>*>(A.y) + off, mk<float>(s[p].x, s[p].y));
        }
      }
    }
    __syncthreads();  // everyone is done with the tile before its buffer is refilled
    if constexpr (TMA) stage ^= 1;
  }
}

cudaError_t launch_streamw(int qc, bool tma, int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                           const StreamRInfo& R, const SweepArgs& A);
cudaError_t occupancy_streamw(int qc, bool tma, size_t smem, int* blocks_per_sm);

}  // namespace fsweep
