// fsweep_stream.cuh — streaming sweep for TABLE-heavy programs (generic FIR filters): the HBM-bound case.
//
// A Series of dsp.Filter / dsp.parallelFilter (their responses are per-bin tables built once per call by cuFFT) with
// gains in between — the shape of the real examples/e8_active_acoustics.py path — has almost
// no arithmetic: per bin it reads sum(n_out*n_in*8 B) of tables and does as many complex multiply-adds.  The generic
// row-distributed interpreter ran it at 4 - 6 % of the HBM roofline (latency / issue bound: profiles/r01_notes.md).
// Here the tables of a TILE of consecutive bins — contiguous in memory per op — are brought into shared memory by an
// asynchronous multi-stage pipeline (cp.async, 8-byte granules: any complex64 row is aligned, no special cases for
// shards or tails), several tiles in flight per block, and the arithmetic runs from shared memory with one thread per
// (bin, column) walking the whole chain privately.  Table gradients are staged in shared memory and written back as
// one contiguous, fully coalesced block per op and tile.  (A first version with one thread per (bin, output row,
// column) and a barrier per op issued 1019 warp instructions per bin and was issue bound: profiles/r01l.)
//
// Two ways of bringing a tile in.  TMA (opt-in, FSWEEP_STREAM_TMA=1, when every table block of the call is 16-byte aligned): persistent
// blocks, a ring of S_TMA_STAGES tile buffers filled by the bulk-copy engine — ONE elected thread issues one
// cp.async.bulk per table op and tile (SASS: UBLKCP) against an mbarrier with the expected byte count, S_TMA_STAGES - 1
// tiles ahead of the arithmetic — and table gradients leave the block the same way (cp.async.bulk shared -> global,
// bulk_group).  Nobody spends an instruction on moving table bytes and the DRAM latency is hidden by the ring rather
// than by occupancy.  Default: every thread copies 8-byte granules with cp.async, one tile per small block — measured
// faster, because shared memory per bin (not DRAM latency) bounds this kernel (fsweep_api.cu, stream_tma_ok).
//
// Supported programs (host-checked, plan->stream): float32; no recursion; ops in {TABLE, PTABLE, GAIN, PGAIN};
// widths <= 16; batch*cols a power of two <= 16; coefficient gradients for TABLE / PTABLE (written per bin) and PGAIN
// (accumulated); dense GAIN only without gradient.  Everything else stays on the generic kernels.
#pragma once
#include <cuda_pipeline.h>

#include "fsweep_tma.cuh"
#include "fsweep_tpc.cuh"

namespace fsweep {

constexpr int SW = 16;        // row slots per (bin, column)
constexpr int S_STAGES = 1;   // tiles in flight per block: shared memory per bin bounds the warps per SM, and more resident blocks hide more latency than a deeper pipeline
constexpr int S_MAXST = 96;   // saved-state entries per (bin, column): sum of op input widths + last output width
constexpr int S_TMA_STAGES = 3;  // TMA path: tile buffers in the ring

struct StreamInfo {
  int n_ops;
  int tb;                     // bins per tile
  int qc;                     // columns (batch*cols), power of two
  int threads;                // tb * qc
  int bytes_per_bin;          // table bytes per bin, all table ops
  int tab_off[MAX_OPS];       // byte offset of op i's block inside a stage (block = tb * row_bytes[i]); -1: no table
  int row_bytes[MAX_OPS];
  int st_off[MAX_OPS + 1];    // offset of op i's input vector in the saved-state record; [n_ops]: final output
  int st_total;
  int n_pgain_acc;            // PGAIN gradient accumulators (sum of widths)
  int pg_off[MAX_OPS];        // offset of op i's PGAIN accumulators, -1: none
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) { __pipeline_memcpy_async(smem, gmem, 8); }

// ---- width-specialised inner loops.  The op chain is runtime data, but its widths are <= 16: one instantiation per
// width keeps the vector an op consumes in REGISTERS and unrolls the dot products — no index arithmetic, one LDS
// (the table entry) per complex multiply-add.  (The generic loops ran 13.9 M warp instructions for 4.8 M of use.)
template <int NIN>
__device__ __forceinline__ void table_rows(const float2* __restrict__ H, const float2* vin, float2* vout, int n_out,
                                           int T) {  // vout[m] = sum_n H[m][n] vin[n]
  float2 v[NIN];
#pragma unroll
  for (int n = 0; n < NIN; ++n) v[n] = vin[(size_t)n * T];
  // (rows are independent accumulation chains: four of them in flight per thread)
#pragma unroll 4
  for (int m = 0; m < n_out; ++m) {
    const float2* h = H + m * NIN;
    float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
#pragma unroll
    for (int n = 0; n < NIN; ++n) {
      const float2 hh = h[n];
      ax = fmaf(hh.x, v[n].x, ax);
      bx = fmaf(-hh.y, v[n].y, bx);
      ay = fmaf(hh.x, v[n].y, ay);
      by = fmaf(hh.y, v[n].x, by);
    }
    vout[(size_t)m * T] = f2(ax + bx, ay + by);
  }
}
template <int NOUT>
__device__ __forceinline__ void table_cols(const float2* __restrict__ H, const float2* go, float2* gi, int n_in,
                                           int T) {  // gi[n] = sum_m conj(H[m][n]) go[m]
  float2 g[NOUT];
#pragma unroll
  for (int m = 0; m < NOUT; ++m) g[m] = go[(size_t)m * T];
#pragma unroll 4
  for (int n = 0; n < n_in; ++n) {
    const float2* h = H + n;
    float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
#pragma unroll
    for (int m = 0; m < NOUT; ++m) {
      const float2 hh = h[m * n_in];
      ax = fmaf(hh.x, g[m].x, ax);
      bx = fmaf(hh.y, g[m].y, bx);
      ay = fmaf(hh.x, g[m].y, ay);
      by = fmaf(-hh.y, g[m].x, by);
    }
    gi[(size_t)n * T] = f2(ax + bx, ay + by);
  }
}
#define FSWEEP_WIDTH_SWITCH(W, CALL)                                                      \
  switch (W) {                                                                            \
    case 1: { constexpr int WW = 1; CALL; } break;                                        \
    case 2: { constexpr int WW = 2; CALL; } break;                                        \
    case 3: { constexpr int WW = 3; CALL; } break;                                        \
    case 4: { constexpr int WW = 4; CALL; } break;                                        \
    case 5: { constexpr int WW = 5; CALL; } break;                                        \
    case 6: { constexpr int WW = 6; CALL; } break;                                        \
    case 7: { constexpr int WW = 7; CALL; } break;                                        \
    case 8: { constexpr int WW = 8; CALL; } break;                                        \
    case 9: { constexpr int WW = 9; CALL; } break;                                        \
    case 10: { constexpr int WW = 10; CALL; } break;                                      \
    case 11: { constexpr int WW = 11; CALL; } break;                                      \
    case 12: { constexpr int WW = 12; CALL; } break;                                      \
    case 13: { constexpr int WW = 13; CALL; } break;                                      \
    case 14: { constexpr int WW = 14; CALL; } break;                                      \
    case 15: { constexpr int WW = 15; CALL; } break;                                      \
    default: { constexpr int WW = 16; CALL; } break;                                      \
  }

// shared memory: [S_STAGES][tb * bytes_per_bin] tables | state columns [n_state][T] | (BWD) gradient columns [2][SW][T] |
//                (BWD) [n_pgain_acc] block accumulators      (table gradients overwrite their tables in the tile)
// where n_state = st_total (BWD: every op input is kept for its gradient) or 2*SW (forward: ping-pong).
// ONE THREAD PER (bin, column): the whole op chain runs privately on thread-private shared-memory columns
// (element i of thread t at [i*T + t]: conflict-free), so there is no barrier between ops and no idle row slot; the
// only exchange is the sum over the columns of a bin in the table gradients (adjacent lanes, shuffles).
template <bool BWD, bool TMA>
__global__ void __launch_bounds__(128) fsweep_stream_kernel(const __grid_constant__ ProgK P,
                                                            const __grid_constant__ StreamInfo S, const SweepArgs A,
                                                            int G) {
  extern __shared__ __align__(128) unsigned char ssm[];
  constexpr int NST = TMA ? S_TMA_STAGES : S_STAGES;
  const int tid = threadIdx.x, T = blockDim.x;
  const int qc = S.qc, tb = S.tb;
  const int c = tid & (qc - 1);  // column (fastest: the lanes of one bin are adjacent)
  const int bi = tid / qc;       // bin inside the tile
  const size_t stage_bytes = (size_t)tb * S.bytes_per_bin;
  const int n_state = BWD ? S.st_total : 2 * SW;
  unsigned char* sTab = ssm;
  float2* sState = reinterpret_cast<float2*>(ssm + NST * stage_bytes) + tid;  // element i: sState[i * T]
  float2* sGrad = sState + (size_t)n_state * T;                                      // [2][SW] columns (BWD)
  // (table gradients have no buffer of their own: dL/dH of an op overwrites H in the tile once the op's input gradient
  // has been formed — a table is read exactly once in the reverse sweep — and leaves the block from there)
  float* sAcc = reinterpret_cast<float*>(sGrad - tid + (BWD ? (size_t)2 * SW * T : 0));
  uint64_t* sBar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sAcc) +
                                               (((BWD ? (size_t)S.n_pgain_acc * 4 : 0) + 15) / 16) * 16);  // [NST], TMA

  const long long n_tiles = (A.n_bins + tb - 1) / tb;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  const int in_ch = P.in_ch, out_ch = P.out_ch, n_ops = S.n_ops;
  const int b = (A.cols == 1) ? c : c / A.cols, cc = c - b * A.cols;
  if (BWD) {
    for (int i = tid; i < S.n_pgain_acc; i += T) sAcc[i] = 0.f;
  }

  if constexpr (TMA) {
    if (tid == 0) {
      for (int i = 0; i < NST; ++i) mbar_init(sBar + i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  // a tile is "full" when all tb bins exist: only full tiles go through the bulk-copy engine (sizes in multiples of 16)
  auto is_full = [&](long long tile) { return (tile + 1) * tb <= A.n_bins; };
  // ---- TMA tile loader: ONE thread, one bulk copy per table op, completion on the stage's mbarrier
  auto issue_tma = [&](long long tile, int stage) {
    if (tile < n_tiles && is_full(tile)) {
      unsigned char* dst = sTab + (size_t)stage * stage_bytes;
      uint32_t total = 0;
      for (int i = 0; i < n_ops; ++i)
        if (S.tab_off[i] >= 0) total += (uint32_t)(tb * S.row_bytes[i]);
      mbar_expect_tx(sBar + stage, total);
      for (int i = 0; i < n_ops; ++i) {
        if (S.tab_off[i] < 0) continue;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                   (size_t)(A.bin_begin + tile * tb) * S.row_bytes[i];
        tma_load_1d(dst + S.tab_off[i], src, (uint32_t)(tb * S.row_bytes[i]), sBar + stage);
      }
    }
  };
  // ---- asynchronous tile loader: every thread copies 8-byte granules of the tile's contiguous table blocks
  auto issue = [&](long long tile, int stage) {
    if (tile < n_tiles) {
      const long long b0 = tile * tb;
      const int nb = (int)min((long long)tb, A.n_bins - b0);
      unsigned char* dst = sTab + (size_t)stage * stage_bytes;
      for (int i = 0; i < n_ops; ++i) {
        if (S.tab_off[i] < 0) continue;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                   (size_t)(A.bin_begin + b0) * S.row_bytes[i];
        const int n8 = nb * S.row_bytes[i] / 8;
        for (int e = tid; e < n8; e += T) cp_async8(dst + S.tab_off[i] + (size_t)e * 8, src + (size_t)e * 8);
      }
    }
    __pipeline_commit();
  };

  long long tile = blockIdx.x;
  if constexpr (TMA) {
    if (tid == 0)
      for (int s = 0; s < NST - 1; ++s) issue_tma(tile + (long long)s * gridDim.x, s);
  } else {
    for (int s = 0; s < NST - 1; ++s) issue(tile + (long long)s * gridDim.x, s);
  }
  int stage = 0;
  uint32_t phases = 0;  // TMA: bit s = parity the next wait on stage s expects
  for (; tile < n_tiles; tile += gridDim.x) {
    if constexpr (TMA) {
      // the stage refilled here was consumed in the previous iteration (its closing __syncthreads is behind us)
      if (tid == 0) {
        // the stage refilled here held the previous tile, whose table gradients leave the block from that very
        // buffer (bulk stores): they must have been READ before the bulk loads overwrite it
        if (BWD) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        issue_tma(tile + (long long)(NST - 1) * gridDim.x, (stage + NST - 1) % NST);
      }
      if (is_full(tile)) {
        mbar_wait(sBar + stage, (phases >> stage) & 1u);
        phases ^= 1u << stage;
      } else {  // the one partial tile at the end of the range: plain loads
        const int nb = (int)(A.n_bins - tile * tb);
        unsigned char* dst = sTab + (size_t)stage * stage_bytes;
        for (int i = 0; i < n_ops; ++i) {
          if (S.tab_off[i] < 0) continue;
          const float2* src = reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                                             (size_t)(A.bin_begin + tile * tb) * S.row_bytes[i]);
          float2* d2 = reinterpret_cast<float2*>(dst + S.tab_off[i]);
          const int n8 = nb * S.row_bytes[i] / 8;
          for (int e = tid; e < n8; e += T) d2[e] = __ldg(src + e);
        }
        __syncthreads();
      }
    } else {
      issue(tile + (long long)(NST - 1) * gridDim.x, (stage + NST - 1) % NST);
      __pipeline_wait_prior(NST - 1);
      __syncthreads();
    }
    unsigned char* tab = sTab + (size_t)stage * stage_bytes;
    const long long bl = tile * tb + bi;
    const bool live = bl < A.n_bins;

    // ---- forward chain (BWD: every op input stays in its own slot of the state record)
    float2* vin = sState + (BWD ? (size_t)S.st_off[0] * T : 0);
    for (int n = 0; n < in_ch; ++n) {
      cx<float> v = mk<float>(0.f, 0.f);
      if (live) v = ld_cx(x + (size_t)b * A.xbs + ((size_t)bl * in_ch + n) * A.cols + cc);
      vin[(size_t)n * T] = f2(v.x, v.y);
    }
    for (int i = 0; i < n_ops; ++i) {
      const OpK& op = P.ops[i];
      float2* vout = BWD ? sState + (size_t)S.st_off[i + 1] * T : sState + (size_t)((i + 1) & 1) * SW * T;
      const int n_in = op.n_in, n_out = op.n_out;
      if (op.kind == FSWEEP_OP_TABLE) {
        const float2* H = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
        FSWEEP_WIDTH_SWITCH(n_in, table_rows<WW>(H, vin, vout, n_out, T))
      } else if (op.kind == FSWEEP_OP_PTABLE) {
        const float2* H = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
        for (int m = 0; m < n_out; ++m) {
          const float2 h = H[m], v = vin[(size_t)m * T];
          vout[(size_t)m * T] = f2(h.x * v.x - h.y * v.y, h.x * v.y + h.y * v.x);
        }
      } else if (op.kind == FSWEEP_OP_GAIN) {
        const float* W = reinterpret_cast<const float*>(op.coef);
        for (int m = 0; m < n_out; ++m) {
          float ax = 0.f, ay = 0.f;
          for (int n = 0; n < n_in; ++n) {
            const float w = __ldg(W + m * n_in + n);
            const float2 v = vin[(size_t)n * T];
            ax = fmaf(w, v.x, ax);
            ay = fmaf(w, v.y, ay);
          }
          vout[(size_t)m * T] = f2(ax, ay);
        }
      } else {  // PGAIN
        const float* W = reinterpret_cast<const float*>(op.coef);
        for (int m = 0; m < n_out; ++m) {
          const float w = __ldg(W + m);
          const float2 v = vin[(size_t)m * T];
          vout[(size_t)m * T] = f2(w * v.x, w * v.y);
        }
      }
      vin = vout;
    }
    // ---- output (forward) / output gradient (backward): vin now points at the final output
    if constexpr (!BWD) {
      if (live) {
        for (int m = 0; m < out_ch; ++m) {
          const float2 o = vin[(size_t)m * T];
          const size_t off = (size_t)b * A.ybs + ((size_t)bl * out_ch + m) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS)
            reinterpret_cast<float*>(A.y)[off] = abs_t(o.x, o.y);
          else
            st_cx(reinterpret_cast<cx<float>*>(A.y) + off, mk<float>(o.x, o.y));
        }
      }
    } else {
      float2* gout = sGrad;  // buffer 0
      for (int m = 0; m < out_ch; ++m) {
        float2 g = f2(0.f, 0.f);
        if (live) {
          const float2 o = vin[(size_t)m * T];
          const size_t off = (size_t)b * A.gybs + ((size_t)bl * out_ch + m) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS) {
            const float ga = __ldg(reinterpret_cast<const float*>(A.gy) + off);
            const float mag = abs_t(o.x, o.y);
            if (mag > 0.f) {
              const float sc = ga * rcp_t(mag);
              g = f2(sc * o.x, sc * o.y);
            }
          } else {
            const cx<float> gv = ld_cx(reinterpret_cast<const cx<float>*>(A.gy) + off);
            g = f2(gv.x, gv.y);
          }
        }
        gout[(size_t)m * T] = g;
      }
      if constexpr (TMA) __syncthreads();  // thread 0 has seen the last bulk stores read the staging buffer
      // ---- reverse sweep
      int buf = 0;
      for (int i = n_ops - 1; i >= 0; --i) {
        const OpK& op = P.ops[i];
        const int n_in = op.n_in, n_out = op.n_out;
        const float2* go = sGrad + (size_t)buf * SW * T;
        float2* gi = sGrad + (size_t)(buf ^ 1) * SW * T;
        const float2* v = sState + (size_t)S.st_off[i] * T;
        if (op.kind == FSWEEP_OP_TABLE) {
          const float2* H = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
          // g_in[n] = sum_m conj(H[m][n]) g[m]  (first: the table gradient below overwrites H)
          FSWEEP_WIDTH_SWITCH(n_out, table_cols<WW>(H, go, gi, n_in, T))
          if (op.acc_mode == ACC_TABLE) {  // dL/dH[m][n] = sum over the bin's columns of g[m] conj(v[n])
            // the qc threads of a bin split the entries among themselves and each sums over ALL columns, reading
            // the neighbours' (thread-private) g and v columns: no shuffle chain per entry
            float2* gt = reinterpret_cast<float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
            __syncwarp();  // every column of the bin has read H
            const float2* go0 = go - c;  // column 0 of this bin
            const float2* v0 = v - c;
            int m = c / n_in, n = c - m * n_in;  // (m, n) walk along with e: no division per entry
            for (int e = c; e < n_out * n_in; e += qc, n += qc) {
              while (n >= n_in) {
                n -= n_in;
                ++m;
              }
              float vx = 0.f, vy = 0.f;
              for (int cq = 0; cq < qc; ++cq) {
                const float2 gm = go0[(size_t)m * T + cq], vn = v0[(size_t)n * T + cq];
                vx = fmaf(gm.x, vn.x, fmaf(gm.y, vn.y, vx));
                vy = fmaf(gm.y, vn.x, fmaf(-gm.x, vn.y, vy));
              }
              gt[e] = f2(vx, vy);
            }
            __syncwarp();  // nobody overwrites a gradient column while a neighbour still reads it
          }
        } else if (op.kind == FSWEEP_OP_PTABLE) {
          const float2* H = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
          float2* gt = reinterpret_cast<float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
          for (int m = 0; m < n_out; ++m) {
            const float2 gm = go[(size_t)m * T], vn = v[(size_t)m * T], h = H[m];
            if (op.acc_mode == ACC_TABLE) {
              float vx = gm.x * vn.x + gm.y * vn.y, vy = gm.y * vn.x - gm.x * vn.y;
              for (int o = qc >> 1; o > 0; o >>= 1) {
                vx += __shfl_xor_sync(FULL, vx, o);
                vy += __shfl_xor_sync(FULL, vy, o);
              }
              if (c == 0) gt[m] = f2(vx, vy);
            }  // (diagonal tables: n entries only, the shuffle sum is cheap)
            gi[(size_t)m * T] = f2(h.x * gm.x + h.y * gm.y, h.x * gm.y - h.y * gm.x);
          }
        } else if (op.kind == FSWEEP_OP_GAIN) {
          const float* W = reinterpret_cast<const float*>(op.coef);
          for (int n = 0; n < n_in; ++n) {
            float ax = 0.f, ay = 0.f;
            for (int m = 0; m < n_out; ++m) {
              const float w = __ldg(W + m * n_in + n);
              const float2 gm = go[(size_t)m * T];
              ax = fmaf(w, gm.x, ax);
              ay = fmaf(w, gm.y, ay);
            }
            gi[(size_t)n * T] = f2(ax, ay);
          }
        } else {  // PGAIN
          const float* W = reinterpret_cast<const float*>(op.coef);
          for (int m = 0; m < n_out; ++m) {
            const float2 gm = go[(size_t)m * T];
            if (S.pg_off[i] >= 0) {
              const float2 vn = v[(size_t)m * T];
              float d = gm.x * vn.x + gm.y * vn.y;  // zero for dead bins (their state and gradient are zero)
              for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
              if ((tid & 31) == 0) atomicAdd(sAcc + S.pg_off[i] + m, d);
            }
            const float w = __ldg(W + m);
            gi[(size_t)m * T] = f2(w * gm.x, w * gm.y);
          }
        }
        buf ^= 1;
      }
      if (A.gx != nullptr && live) {
        const float2* gfin = sGrad + (size_t)buf * SW * T;
        for (int n = 0; n < in_ch; ++n) {
          const float2 gv = gfin[(size_t)n * T];
          st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)b * A.gxbs + ((size_t)bl * in_ch + n) * A.cols + cc,
                mk<float>(gv.x, gv.y));
        }
      }
      if constexpr (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging -> async proxy
      __syncthreads();
      // ---- table gradients of the tile: contiguous blocks (TMA: one bulk store per op; else coalesced 8-byte stores)
      {
        const long long b0 = tile * tb;
        const int nb = (int)min((long long)tb, A.n_bins - b0);
        const bool bulk = TMA && nb == tb;
        for (int i = 0; i < n_ops; ++i) {
          const OpK& op = P.ops[i];
          if (op.acc_mode != ACC_TABLE || S.tab_off[i] < 0) continue;
          unsigned char* gdst = reinterpret_cast<unsigned char*>(op.gtab) + (size_t)(A.bin_begin + b0) * S.row_bytes[i];
          if (bulk) {
            if (tid == 0) tma_store_1d(gdst, tab + S.tab_off[i], (uint32_t)(tb * S.row_bytes[i]));
          } else {
            float2* dst = reinterpret_cast<float2*>(gdst);
            const float2* src = reinterpret_cast<const float2*>(tab + S.tab_off[i]);
            const int n8 = nb * S.row_bytes[i] / 8;
            for (int e = tid; e < n8; e += T) dst[e] = src[e];
          }
        }
        if (bulk && tid == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    __syncthreads();  // everyone is done with this stage (and the staging buffer) before they are refilled
    stage = (stage + 1) % NST;
  }
  if constexpr (TMA) {
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    __pipeline_wait_prior(0);
  }

  if constexpr (BWD) {
    // PGAIN accumulators -> partial[(row_off + 0) * G + row]
    __syncthreads();
    float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    for (int i = 0; i < n_ops; ++i) {
      if (S.pg_off[i] < 0) continue;
      const OpK& op = P.ops[i];
      for (int m = tid; m < op.n_out; m += T) partial[op.row_off * G + m] = sAcc[S.pg_off[i] + m];
    }
  }
}

cudaError_t launch_stream(bool bwd, bool tma, int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                          const SweepArgs& A, int G);
cudaError_t occupancy_stream(bool bwd, bool tma, int threads, size_t smem, int* blocks_per_sm);

}  // namespace fsweep
