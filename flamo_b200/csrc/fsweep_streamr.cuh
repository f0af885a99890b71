// fsweep_streamr.cuh — streaming sweep, FORWARD, with the signal state in REGISTERS.
//
// fsweep_stream_kernel keeps every thread's state vectors in (thread-private columns of) shared memory: 2 x 16 slots
// x 8 B per thread — for a four-column signal that is 1 KB per bin next to 936 B of tables, so half of the shared
// memory that decides how many bins an SM has in flight held state, not tables (r02g capture: 13 one-warp blocks per
// SM, issue slots 35 % busy, 13.6 warp instructions per complex multiply-add: 24 % of the HBM roofline).  Here
//   * the state of a (bin, column) thread is two register arrays v[W], o[W] (W = 4 | 8 | 16, the program's widest op):
//     all indexing is compile time — the op chain is runtime data, but every dense op body exists for each exact
//     input width and each output-width class of four, fully unrolled, and is reached by a switch (real, uniform
//     branches): 1 LDS.64 + 4 FFMA per complex multiply-add;
//   * shared memory holds ONLY table tiles, brought in by the TMA bulk-copy engine one tile ahead of the arithmetic
//     (cp.async.bulk + mbarrier; cp.async granules when a bin shard breaks the 16-byte alignment);
//   * blocks of 64 threads (16 bins of 4 columns), as many resident as registers allow.
// (Rows padded to an odd number of 8-byte units would keep the 8 bins of a warp on different banks — the 416-byte rows
// of a 13 x 4 table put bins 0 and 4 on the same ones — but the scatter loop that padding needs cost more issue slots
// than the conflicts: dropped.)
// Same supported programs as fsweep_stream_kernel (TABLE / PTABLE / GAIN / PGAIN, widths <= 16, batch * cols a power of
// two <= 16), forward only; the backward pass and the TMA variant stay in fsweep_stream.cuh.
#pragma once
#include "fsweep_stream.cuh"

namespace fsweep {

struct StreamRInfo {
  int pad_units[MAX_OPS];  // row length in 8-byte units; 0: no table
  int pad_off[MAX_OPS];    // unit offset of op i's block (tb rows) inside the tile
  int tile_units;          // units per tile
  int aligned16;           // every table block of the call starts on a 16-byte boundary: 16-byte cp.async granules
};

// Dense op body for EXACT input width NI and output-width class NO (a multiple of 4): o[m] = sum_n C[m][n] v[n] with
// C complex (a table row in shared memory) or real (a GAIN matrix in global memory).  The (NI, NO) pair is picked by a
// switch on kernel-uniform values, i.e. by real branches: nothing of the 16 x 16 worst case is issued for a 4 x 13
// table (a first version guarded an unrolled 16 x 16 body instead — ptxas predicated it and the kernel issued 2900
// warp instructions per tile for 650 of use).  Only the last four rows of a class carry a guard.
template <int W, int NI, int NO, bool REAL>
__device__ __forceinline__ void r_dense(const void* __restrict__ Cp, const float2 (&v)[W], float2 (&o)[W], int n_out) {
#pragma unroll
  for (int m = 0; m < NO; ++m) {
    if (m < NO - 4 || m < n_out) {
      if (REAL) {
        const float* w = reinterpret_cast<const float*>(Cp) + m * NI;
        float ax = 0.f, ay = 0.f;
#pragma unroll
        for (int n = 0; n < NI; ++n) {
          const float wv = __ldg(w + n);
          ax = fmaf(wv, v[n].x, ax);
          ay = fmaf(wv, v[n].y, ay);
        }
        o[m] = f2(ax, ay);
      } else {
        const float2* h = reinterpret_cast<const float2*>(Cp) + m * NI;
        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
#pragma unroll
        for (int n = 0; n < NI; ++n) {
          const float2 hh = h[n];
          ax = fmaf(hh.x, v[n].x, ax);
          bx = fmaf(-hh.y, v[n].y, bx);
          ay = fmaf(hh.x, v[n].y, ay);
          by = fmaf(hh.y, v[n].x, by);
        }
        o[m] = f2(ax + bx, ay + by);
      }
    }
  }
}
template <int W, int NI, bool REAL>
__device__ __forceinline__ void r_dense_no(const void* Cp, const float2 (&v)[W], float2 (&o)[W], int n_out) {
  if constexpr (NI <= W) {
    if (W == 4 || n_out <= 4) {
      r_dense<W, NI, 4, REAL>(Cp, v, o, n_out);
    } else if (W == 8 || n_out <= 8) {
      r_dense<W, NI, (W >= 8 ? 8 : W), REAL>(Cp, v, o, n_out);
    } else if (n_out <= 12) {
      r_dense<W, NI, (W >= 16 ? 12 : W), REAL>(Cp, v, o, n_out);
    } else {
      r_dense<W, NI, W, REAL>(Cp, v, o, n_out);
    }
  }
}
template <int W, bool REAL>
__device__ __forceinline__ void r_dense_op(const void* Cp, const float2 (&v)[W], float2 (&o)[W], int n_in, int n_out) {
  switch (n_in) {
    case 1: r_dense_no<W, 1, REAL>(Cp, v, o, n_out); break;
    case 2: r_dense_no<W, 2, REAL>(Cp, v, o, n_out); break;
    case 3: r_dense_no<W, 3, REAL>(Cp, v, o, n_out); break;
    case 4: r_dense_no<W, 4, REAL>(Cp, v, o, n_out); break;
    case 5: r_dense_no<W, 5, REAL>(Cp, v, o, n_out); break;
    case 6: r_dense_no<W, 6, REAL>(Cp, v, o, n_out); break;
    case 7: r_dense_no<W, 7, REAL>(Cp, v, o, n_out); break;
    case 8: r_dense_no<W, 8, REAL>(Cp, v, o, n_out); break;
    case 9: r_dense_no<W, 9, REAL>(Cp, v, o, n_out); break;
    case 10: r_dense_no<W, 10, REAL>(Cp, v, o, n_out); break;
    case 11: r_dense_no<W, 11, REAL>(Cp, v, o, n_out); break;
    case 12: r_dense_no<W, 12, REAL>(Cp, v, o, n_out); break;
    case 13: r_dense_no<W, 13, REAL>(Cp, v, o, n_out); break;
    case 14: r_dense_no<W, 14, REAL>(Cp, v, o, n_out); break;
    case 15: r_dense_no<W, 15, REAL>(Cp, v, o, n_out); break;
    default: r_dense_no<W, 16, REAL>(Cp, v, o, n_out); break;
  }
}

// TMA: every table block of the call starts on a 16-byte boundary (whole tables; not every bin shard): the tiles come
// in through the bulk-copy engine — ONE elected thread issues one cp.async.bulk per table op and tile against an
// mbarrier with the expected byte count, one tile AHEAD of the arithmetic (two tile buffers) — so that nobody spends an
// instruction on moving table bytes (the per-thread cp.async loop was 36 % of all issued instructions: r02t capture).
// Otherwise: every thread copies 8-byte granules of the tile with cp.async, one buffer.
template <int W, bool TMA>
__global__ void __launch_bounds__(256) fsweep_streamr_kernel(const __grid_constant__ ProgK P,
                                                            const __grid_constant__ StreamInfo S,
                                                            const __grid_constant__ StreamRInfo R, const SweepArgs A) {
  extern __shared__ __align__(128) unsigned char rsm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int qc = S.qc, tb = S.tb;
  const int c = tid & (qc - 1);
  const int bi = tid / qc;
  const long long n_tiles = (A.n_bins + tb - 1) / tb;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  const int in_ch = P.in_ch, out_ch = P.out_ch, n_ops = S.n_ops;
  const int b = (A.cols == 1) ? c : c / A.cols, cc = c - b * A.cols;
  const int tile_units = R.tile_units;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(rsm + (size_t)(TMA ? 2 : 1) * tile_units * 8);  // [2], TMA

  auto is_full = [&](long long tile) { return (tile + 1) * tb <= A.n_bins; };
  auto issue_tma = [&](long long tile, int stage) {  // one thread
    if (tile < n_tiles && is_full(tile)) {
      float2* dst = reinterpret_cast<float2*>(rsm) + (size_t)stage * tile_units;
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
    float2* sTab = reinterpret_cast<float2*>(rsm) + (size_t)stage * tile_units;
    bool plain = !TMA;
    if constexpr (TMA) {
      // the other buffer was released by the barrier that closed the previous iteration
      if (tid == 0) issue_tma(tile + gridDim.x, stage ^ 1);
      plain = !is_full(tile);
    }
    if (plain) {  // cp.async path, and the one partial tile at the end of a range on the TMA path
      const int nb = (int)min((long long)tb, A.n_bins - b0);
      for (int i = 0; i < n_ops; ++i) {
        if (R.pad_units[i] == 0) continue;
        const float2* src = reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(P.ops[i].coef) +
                                                            (size_t)(A.bin_begin + b0) * S.row_bytes[i]);
        float2* dst = sTab + R.pad_off[i];
        const int n8 = nb * R.pad_units[i];
        if (R.aligned16 && !(n8 & 1)) {
          const float4* s4 = reinterpret_cast<const float4*>(src);
          float4* d4 = reinterpret_cast<float4*>(dst);
          for (int e = tid; e < (n8 >> 1); e += T) __pipeline_memcpy_async(d4 + e, s4 + e, 16);
        } else {
          for (int e = tid; e < n8; e += T) cp_async8(dst + e, src + e);
        }
      }
      __pipeline_commit();
    }
    // the thread's input column is requested while the tile is in flight
    const long long bl = b0 + bi;
    const bool live = bl < A.n_bins;
    float2 v[W], o[W];
#pragma unroll
    for (int n = 0; n < W; ++n) {
      v[n] = f2(0.f, 0.f);
      if (n < in_ch && live) {
        const cx<float> t = ld_cx(x + (size_t)b * A.xbs + ((size_t)bl * in_ch + n) * A.cols + cc);
        v[n] = f2(t.x, t.y);
      }
    }
    if (plain) {
      __pipeline_wait_prior(0);
      __syncthreads();
    } else {
      mbar_wait(sBar + stage, (phases >> stage) & 1u);
      phases ^= 1u << stage;
    }

    for (int i = 0; i < n_ops; ++i) {
      const OpK& op = P.ops[i];
      const int n_in = op.n_in, n_out = op.n_out;
      if (op.kind == FSWEEP_OP_TABLE) {
        r_dense_op<W, false>(sTab + R.pad_off[i] + bi * R.pad_units[i], v, o, n_in, n_out);
      } else if (op.kind == FSWEEP_OP_PTABLE) {
        const float2* H = sTab + R.pad_off[i] + bi * R.pad_units[i];
#pragma unroll
        for (int m = 0; m < W; ++m)
          if (m < n_out) {
            const float2 h = H[m];
            o[m] = f2(h.x * v[m].x - h.y * v[m].y, h.x * v[m].y + h.y * v[m].x);
          }
      } else if (op.kind == FSWEEP_OP_GAIN) {
        r_dense_op<W, true>(op.coef, v, o, n_in, n_out);
      } else {  // PGAIN
        const float* Wg = reinterpret_cast<const float*>(op.coef);
#pragma unroll
        for (int m = 0; m < W; ++m)
          if (m < n_out) {
            const float w = __ldg(Wg + m);
            o[m] = f2(w * v[m].x, w * v[m].y);
          }
      }
#pragma unroll
      for (int m = 0; m < W; ++m) v[m] = o[m];
    }
    if (live) {
#pragma unroll
      for (int m = 0; m < W; ++m)
        if (m < out_ch) {
          const size_t off = (size_t)b * A.ybs + ((size_t)bl * out_ch + m) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS)
            reinterpret_cast<float*>(A.y)[off] = abs_t(v[m].x, v[m].y);
          else
            st_cx(reinterpret_cast<cx<float>*>(A.y) + off, mk<float>(v[m].x, v[m].y));
        }
    }
    __syncthreads();  // everyone is done with the tile before its buffer is refilled
    if constexpr (TMA) stage ^= 1;
  }
}

cudaError_t launch_streamr(int w, bool tma, int grid, int threads, size_t smem, cudaStream_t st, const ProgK& P,
                           const StreamInfo& S, const StreamRInfo& R, const SweepArgs& A);
cudaError_t occupancy_streamr(int w, bool tma, int threads, size_t smem, int* blocks_per_sm);

}  // namespace fsweep
