// fsweep_stream.cuh — streaming sweep for TABLE-heavy programs (generic FIR filters): the HBM-bound case.
//
// A Series of dsp.Filter / dsp.parallelFilter (their responses are per-bin tables built once per call by cuFFT) with
// gains in between — the shape of the real examples/e8_active_acoustics.py path — has almost
// no arithmetic: per bin it reads sum(n_out*n_in*8 B) of tables and does as many complex multiply-adds.  The generic
// row-distributed interpreter ran it at 4 - 6 % of the HBM roofline (latency / issue bound: profiles/r01_notes.md).
// Here the tables of a TILE of consecutive bins — contiguous in memory per op — are brought into shared memory by an
// asynchronous multi-stage pipeline (cp.async, 8-byte granules: any complex64 row is aligned, no special cases for
// shards or tails), several tiles in flight per block, and the arithmetic runs from shared memory with one thread per
// (bin, output row, column).  Table gradients are staged in shared memory and written back as one contiguous,
// fully coalesced block per op and tile.
//
// Supported programs (host-checked, plan->stream): float32; no recursion; ops in {TABLE, PTABLE, GAIN, PGAIN};
// widths <= 16; batch*cols a power of two <= 16; coefficient gradients for TABLE / PTABLE (written per bin) and PGAIN
// (accumulated); dense GAIN only without gradient.  Everything else stays on the generic kernels.
#pragma once
#include <cuda_pipeline.h>

#include "fsweep_tpc.cuh"

namespace fsweep {

constexpr int SW = 16;        // row slots per (bin, column)
constexpr int S_STAGES = 3;   // tiles in flight
constexpr int S_MAXST = 96;   // saved-state entries per (bin, column): sum of op input widths + last output width

struct StreamInfo {
  int n_ops;
  int tb;                     // bins per tile
  int qc;                     // columns (batch*cols), power of two
  int threads;                // tb * SW * qc
  int bytes_per_bin;          // table bytes per bin, all table ops
  int tab_off[MAX_OPS];       // byte offset of op i's block inside a stage (block = tb * row_bytes[i]); -1: no table
  int row_bytes[MAX_OPS];
  int st_off[MAX_OPS + 1];    // offset of op i's input vector in the saved-state record; [n_ops]: final output
  int st_total;
  int n_pgain_acc;            // PGAIN gradient accumulators (sum of widths)
  int pg_off[MAX_OPS];        // offset of op i's PGAIN accumulators, -1: none
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) { __pipeline_memcpy_async(smem, gmem, 8); }

// shared memory: [S_STAGES][tb * bytes_per_bin] tables | [tb*qc][st_total] states | [2][tb*qc][SW] gradients |
//                [tb * bytes_per_bin] table-gradient staging (BWD) | [n_pgain_acc] block accumulators (BWD)
template <bool BWD>
__global__ void __launch_bounds__(512) fsweep_stream_kernel(const __grid_constant__ ProgK P,
                                                            const __grid_constant__ StreamInfo S, const SweepArgs A,
                                                            int G) {
  extern __shared__ __align__(16) unsigned char ssm[];
  const int tid = threadIdx.x;
  const int qc = S.qc, tb = S.tb;
  const int c = tid & (qc - 1);          // column (fastest: the lanes of one (bin, row) are adjacent)
  const int r = (tid / qc) & (SW - 1);   // row slot
  const int bi = tid / (qc * SW);        // bin inside the tile
  const size_t stage_bytes = (size_t)tb * S.bytes_per_bin;
  unsigned char* sTab = ssm;
  float2* sState = reinterpret_cast<float2*>(ssm + S_STAGES * stage_bytes);
  float2* sGrad = sState + (size_t)tb * qc * S.st_total;
  unsigned char* sGTab = reinterpret_cast<unsigned char*>(sGrad + (size_t)2 * tb * qc * SW);
  float* sAcc = reinterpret_cast<float*>(sGTab + (BWD ? stage_bytes : 0));
  float2* myState = sState + (size_t)(bi * qc + c) * S.st_total;
  float2* myGrad = sGrad + (size_t)(bi * qc + c) * SW;  // + buf * tb*qc*SW

  const long long n_tiles = (A.n_bins + tb - 1) / tb;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  const int in_ch = P.in_ch, out_ch = P.out_ch;
  double lacc = 0.0;
  if (BWD) {
    for (int i = tid; i < S.n_pgain_acc; i += blockDim.x) sAcc[i] = 0.f;
  }

  // ---- asynchronous tile loader: every thread copies 8-byte granules of the tile's contiguous table blocks
  auto issue = [&](long long tile, int stage) {
    if (tile < n_tiles) {
      const long long b0 = tile * tb;
      const int nb = (int)min((long long)tb, A.n_bins - b0);
      unsigned char* dst = sTab + (size_t)stage * stage_bytes;
      for (int i = 0; i < S.n_ops; ++i) {
        if (S.tab_off[i] < 0) continue;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                   (size_t)(A.bin_begin + b0) * S.row_bytes[i];
        const int n8 = nb * S.row_bytes[i] / 8;
        for (int e = tid; e < n8; e += blockDim.x) cp_async8(dst + S.tab_off[i] + (size_t)e * 8, src + (size_t)e * 8);
      }
    }
    __pipeline_commit();
  };

  long long tile = blockIdx.x;
  for (int s = 0; s < S_STAGES - 1; ++s) issue(tile + (long long)s * gridDim.x, s);
  int stage = 0;
  for (; tile < n_tiles; tile += gridDim.x) {
    issue(tile + (long long)(S_STAGES - 1) * gridDim.x, (stage + S_STAGES - 1) % S_STAGES);
    __pipeline_wait_prior(S_STAGES - 1);
    __syncthreads();
    const unsigned char* tab = sTab + (size_t)stage * stage_bytes;
    const long long bl = tile * tb + bi;
    const bool live = bl < A.n_bins;
    const int b = (A.cols == 1) ? c : c / A.cols, cc = c - b * A.cols;

    // ---- forward chain, saving every op's input
    if (r < in_ch) {
      cx<float> v = mk<float>(0.f, 0.f);
      if (live) v = ld_cx(x + (size_t)b * A.xbs + ((size_t)bl * in_ch + r) * A.cols + cc);
      myState[S.st_off[0] + r] = f2(v.x, v.y);
    }
    for (int i = 0; i < S.n_ops; ++i) {
      __syncthreads();
      const OpK& op = P.ops[i];
      if (r < op.n_out) {
        const float2* vin = myState + S.st_off[i];
        float2 acc = f2(0.f, 0.f);
        if (op.kind == FSWEEP_OP_TABLE) {
          const float2* H = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]) + r * op.n_in;
          for (int n = 0; n < op.n_in; ++n) {
            const float2 h = H[n], v = vin[n];
            acc.x = fmaf(h.x, v.x, fmaf(-h.y, v.y, acc.x));
            acc.y = fmaf(h.x, v.y, fmaf(h.y, v.x, acc.y));
          }
        } else if (op.kind == FSWEEP_OP_PTABLE) {
          const float2 h = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i])[r];
          const float2 v = vin[r];
          acc = f2(h.x * v.x - h.y * v.y, h.x * v.y + h.y * v.x);
        } else if (op.kind == FSWEEP_OP_GAIN) {
          const float* W = reinterpret_cast<const float*>(op.coef) + r * op.n_in;
          for (int n = 0; n < op.n_in; ++n) {
            const float w = __ldg(W + n);
            acc.x = fmaf(w, vin[n].x, acc.x);
            acc.y = fmaf(w, vin[n].y, acc.y);
          }
        } else {  // PGAIN
          const float w = __ldg(reinterpret_cast<const float*>(op.coef) + r);
          acc = f2(w * vin[r].x, w * vin[r].y);
        }
        myState[S.st_off[i + 1] + r] = acc;
      }
    }
    // ---- output / output gradient
    float2 g = f2(0.f, 0.f);
    if (r < out_ch && live) {
      const float2 o = myState[S.st_off[S.n_ops] + r];
      const size_t off = ((size_t)bl * out_ch + r) * A.cols + cc;
      if constexpr (!BWD) {
        if (A.epilogue == FSWEEP_EPI_ABS)
          reinterpret_cast<float*>(A.y)[(size_t)b * A.ybs + off] = abs_t(o.x, o.y);
        else
          st_cx(reinterpret_cast<cx<float>*>(A.y) + (size_t)b * A.ybs + off, mk<float>(o.x, o.y));
      } else {
        if (A.epilogue == FSWEEP_EPI_ABS) {
          const float ga = __ldg(reinterpret_cast<const float*>(A.gy) + (size_t)b * A.gybs + off);
          const float mag = abs_t(o.x, o.y);
          if (mag > 0.f) {
            const float sc = ga * rcp_t(mag);
            g = f2(sc * o.x, sc * o.y);
          }
        } else {
          const cx<float> gv = ld_cx(reinterpret_cast<const cx<float>*>(A.gy) + (size_t)b * A.gybs + off);
          g = f2(gv.x, gv.y);
        }
      }
    }
    if constexpr (BWD) {
      // ---- reverse sweep: g lives in sGrad[buf] (row slots), buf flips per op
      int buf = 0;
      myGrad[r] = g;  // rows >= out_ch and dead bins: zero
      for (int i = S.n_ops - 1; i >= 0; --i) {
        __syncthreads();
        const OpK& op = P.ops[i];
        const float2* gout = myGrad + (size_t)buf * tb * qc * SW;
        float2* gin = myGrad + (size_t)(buf ^ 1) * tb * qc * SW;
        const float2* vin = myState + S.st_off[i];
        const bool dense = op.kind == FSWEEP_OP_TABLE || op.kind == FSWEEP_OP_GAIN;
        // coefficient gradient
        if (op.acc_mode == ACC_TABLE) {
          float2* gt = reinterpret_cast<float2*>(sGTab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]);
          if (op.kind == FSWEEP_OP_TABLE) {
            const float2 go = r < op.n_out ? gout[r] : f2(0.f, 0.f);
            for (int n = 0; n < op.n_in; ++n) {
              const float2 v = vin[n];
              float vx = go.x * v.x + go.y * v.y, vy = go.y * v.x - go.x * v.y;  // g conj(v)
              for (int o = qc >> 1; o > 0; o >>= 1) {
                vx += __shfl_xor_sync(FULL, vx, o);
                vy += __shfl_xor_sync(FULL, vy, o);
              }
              if (c == 0 && r < op.n_out) gt[r * op.n_in + n] = f2(vx, vy);
            }
          } else {  // PTABLE
            const float2 go = r < op.n_out ? gout[r] : f2(0.f, 0.f);
            const float2 v = r < op.n_out ? vin[r] : f2(0.f, 0.f);
            float vx = go.x * v.x + go.y * v.y, vy = go.y * v.x - go.x * v.y;
            for (int o = qc >> 1; o > 0; o >>= 1) {
              vx += __shfl_xor_sync(FULL, vx, o);
              vy += __shfl_xor_sync(FULL, vy, o);
            }
            if (c == 0 && r < op.n_out) gt[r] = f2(vx, vy);
          }
        } else if (op.kind == FSWEEP_OP_PGAIN && S.pg_off[i] >= 0 && r < op.n_out && live) {
          atomicAdd(sAcc + S.pg_off[i] + r, gout[r].x * vin[r].x + gout[r].y * vin[r].y);
        }
        // gradient to the op input: g_in[n] = sum_m conj(H[m][n]) g[m]
        float2 acc = f2(0.f, 0.f);
        if (r < op.n_in) {
          if (op.kind == FSWEEP_OP_TABLE) {
            const float2* H = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i]) + r;
            for (int m = 0; m < op.n_out; ++m) {
              const float2 h = H[m * op.n_in], gm = gout[m];
              acc.x = fmaf(h.x, gm.x, fmaf(h.y, gm.y, acc.x));
              acc.y = fmaf(h.x, gm.y, fmaf(-h.y, gm.x, acc.y));
            }
          } else if (op.kind == FSWEEP_OP_PTABLE) {
            const float2 h = reinterpret_cast<const float2*>(tab + S.tab_off[i] + (size_t)bi * S.row_bytes[i])[r];
            const float2 gm = gout[r];
            acc = f2(h.x * gm.x + h.y * gm.y, h.x * gm.y - h.y * gm.x);
          } else if (op.kind == FSWEEP_OP_GAIN) {
            const float* W = reinterpret_cast<const float*>(op.coef) + r;
            for (int m = 0; m < op.n_out; ++m) {
              const float w = __ldg(W + m * op.n_in);
              acc.x = fmaf(w, gout[m].x, acc.x);
              acc.y = fmaf(w, gout[m].y, acc.y);
            }
          } else {
            const float w = __ldg(reinterpret_cast<const float*>(op.coef) + r);
            acc = f2(w * gout[r].x, w * gout[r].y);
          }
        }
        (void)dense;
        gin[r] = acc;  // rows >= n_in: zero
        buf ^= 1;
      }
      __syncthreads();
      // ---- table gradients of the tile: contiguous blocks, coalesced 8-byte stores
      {
        const long long b0 = tile * tb;
        const int nb = (int)min((long long)tb, A.n_bins - b0);
        for (int i = 0; i < S.n_ops; ++i) {
          const OpK& op = P.ops[i];
          if (op.acc_mode != ACC_TABLE || S.tab_off[i] < 0) continue;
          float2* dst = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(op.gtab) +
                                                  (size_t)(A.bin_begin + b0) * S.row_bytes[i]);
          const float2* src = reinterpret_cast<const float2*>(sGTab + S.tab_off[i]);
          const int n8 = nb * S.row_bytes[i] / 8;
          for (int e = tid; e < n8; e += blockDim.x) dst[e] = src[e];
        }
      }
      if (A.gx != nullptr && r < in_ch && live) {
        const float2 gi = (myGrad + (size_t)buf * tb * qc * SW)[r];
        st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)b * A.gxbs + ((size_t)bl * in_ch + r) * A.cols + cc,
              mk<float>(gi.x, gi.y));
      }
    }
    __syncthreads();  // everyone is done with this stage (and the staging buffers) before they are refilled
    stage = (stage + 1) % S_STAGES;
  }
  __pipeline_wait_prior(0);

  if constexpr (BWD) {
    // PGAIN accumulators -> partial[(row_off + 0) * G + row]
    __syncthreads();
    float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    for (int i = 0; i < S.n_ops; ++i) {
      if (S.pg_off[i] < 0) continue;
      const OpK& op = P.ops[i];
      for (int m = tid; m < op.n_out; m += blockDim.x) partial[op.row_off * G + m] = sAcc[S.pg_off[i] + m];
    }
  }
  (void)lacc;
}

cudaError_t launch_stream(bool bwd, int grid, size_t smem, cudaStream_t st, const ProgK& P, const StreamInfo& S,
                          const SweepArgs& A, int G);
cudaError_t occupancy_stream(bool bwd, int threads, size_t smem, int* blocks_per_sm);

}  // namespace fsweep
