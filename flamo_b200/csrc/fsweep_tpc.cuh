// fsweep_tpc.cuh — COMPACT thread-per-bin sweep for small FDN loops (width <= 8, float32).
//
// Same pattern and math as fsweep_loop.cuh / fsweep_tpb.cuh
//     [GAIN N x 1]  ->  RECURSION( diagonal chain ; one real N x N matrix )  ->  [GAIN 1 x N]
// with ONE THREAD per frequency bin, but built around what the ncu captures of the unrolled thread-per-bin kernels
// showed (profiles/r01e_ncu_full_sweep_kernels.md): at M ~ 5e4 bins every thread owns ONE bin, so every instruction
// of the kernel is executed once per warp, cold, and the kernel runs at the speed of INSTRUCTION FETCH
// (131 - 217 KB of unrolled SASS, `stalled_no_instruction` = 70 % of the time).  Here
//   * the per-bin matrix A = I - D(w) W lives in thread-private SHARED-MEMORY columns (A[i][j] of thread t at
//     float2 index (i*NP + j)*BLOCK + t: conflict-free), so that the LU factorisation is a RUNTIME loop over the
//     elimination steps and no row ever has to be addressed through a select chain;
//   * 576 B of shared memory per thread and <= 168 registers let 6 blocks of 64 threads live on an SM:
//     56 832 resident threads cover the 48 001 bins of the headline config in ONE wave (the unrolled backward
//     kernel needed 255 registers -> 4 blocks -> a two-wave tail);
//   * gradient accumulators (W_fb: N x N, the two gain vectors, the diagonal op) stay in registers across bins and
//     are reduced once per block through the then-idle matrix storage into the `partial` layout shared with the
//     other kernel families (fsweep_finalize_kernel).
// LU convention: P A = L U with full row interchanges, 1/U_kk stored on the diagonal; the row map is one packed
// word, applied by the solves as a gather / scatter through the shared-memory vector slot.  The k loop of the
// factorisation is a runtime loop, the triangular solves are fully static (vector in registers, every matrix
// element one LDS at a constant offset: independent loads that issue back to back).
#pragma once
#include "fsweep_tma.cuh"
#include "fsweep_tpb.cuh"

namespace fsweep {

constexpr int TPC_BLOCK = 64;

__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) { return f2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc - a*b
__device__ __forceinline__ float2 cnma2(float2 acc, float2 a, float2 b) {
  acc.x = fmaf(-a.x, b.x, acc.x);
  acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(-a.x, b.y, acc.y);
  acc.y = fmaf(-a.y, b.x, acc.y);
  return acc;
}
// acc - conj(a)*b
__device__ __forceinline__ float2 cnmaj2(float2 acc, float2 a, float2 b) {
  acc.x = fmaf(-a.x, b.x, acc.x);
  acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(-a.x, b.y, acc.y);
  acc.y = fmaf(a.y, b.x, acc.y);
  return acc;
}

template <int NP>
struct TpcMat {
  float2* a;  // this thread's column base: element (i, j) at a[(i*NP + j) * TPC_BLOCK]
  float2* v;  // this thread's vector slot: element i at v[i * TPC_BLOCK]
  __device__ __forceinline__ float2& at(int i, int j) const { return a[(i * NP + j) * TPC_BLOCK]; }
  __device__ __forceinline__ float2& vec(int i) const { return v[i * TPC_BLOCK]; }

  // In-place P A = L U with partial pivoting and FULL row interchanges (multipliers move with their rows), unit-lower
  // L below the diagonal, U above, 1/U_kk on the diagonal.  Returns the row map packed 3 bits per row:
  // row i of P A is row ((pvec >> 3i) & 7) of A.  The k loop stays a runtime loop (compact code); the work inside is
  // written with static inner loops so that the loads of one step are independent and issue back to back.
  __device__ __forceinline__ unsigned factor() const {
    unsigned pvec = 0u;
#pragma unroll
    for (int i = 0; i < NP; ++i) pvec |= (unsigned)i << (3 * i);
#pragma unroll 1
    for (int k = 0; k < NP; ++k) {
      float2* colk = a + k * TPC_BLOCK;  // (r, k) at colk[r*NP*TPC_BLOCK]
      float best = -1.f;
      int pr = k;
#pragma unroll
      for (int r = 0; r < NP; ++r) {
        if (r >= k) {
          const float2 c = colk[r * NP * TPC_BLOCK];
          const float m = c.x * c.x + c.y * c.y;
          if (m > best) {
            best = m;
            pr = r;
          }
        }
      }
      float2* rowk = a + k * NP * TPC_BLOCK;
      if (pr != k) {
        float2* rowp = a + pr * NP * TPC_BLOCK;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float2 t = rowk[j * TPC_BLOCK];
          rowk[j * TPC_BLOCK] = rowp[j * TPC_BLOCK];
          rowp[j * TPC_BLOCK] = t;
        }
        const unsigned fx = ((pvec >> (3 * k)) ^ (pvec >> (3 * pr))) & 7u;
        pvec ^= (fx << (3 * k)) | (fx << (3 * pr));
      }
      float2 prow[NP];  // pivot row in registers (entries j < k are loaded but unused)
#pragma unroll
      for (int j = 0; j < NP; ++j) prow[j] = rowk[j * TPC_BLOCK];
      const float2 d = colk[k * NP * TPC_BLOCK];
      const float id = rcp_t(d.x * d.x + d.y * d.y);
      const float2 inv = f2(d.x * id, -d.y * id);
      colk[k * NP * TPC_BLOCK] = inv;
#pragma unroll 1
      for (int r = k + 1; r < NP; ++r) {
        float2* rowr = a + r * NP * TPC_BLOCK;
        const float2 l = cmul2(rowr[k * TPC_BLOCK], inv);
        rowr[k * TPC_BLOCK] = l;
#pragma unroll
        for (int j = 1; j < NP; ++j)
          if (j > k) rowr[j * TPC_BLOCK] = cnma2(rowr[j * TPC_BLOCK], l, prow[j]);
      }
    }
    return pvec;
  }

  // A x = b.  b is in the vector slot on entry; x is returned in registers.  Fully static: every matrix element is
  // one LDS at a constant offset, the row permutation is a gather through the vector slot.
  __device__ __forceinline__ void solve(unsigned pvec, float2 (&x)[NP]) const {
#pragma unroll
    for (int i = 0; i < NP; ++i) x[i] = vec((int)((pvec >> (3 * i)) & 7u));
#pragma unroll
    for (int i = 1; i < NP; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) x[i] = cnma2(x[i], at(i, j), x[j]);
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
#pragma unroll
      for (int j = i + 1; j < NP; ++j) x[i] = cnma2(x[i], at(i, j), x[j]);
      x[i] = cmul2(x[i], at(i, i));
    }
  }

  // A^H lam = g with A = P^T L U:  U^H w = g,  L^H v = w,  lam[pvec_i] = v_i.  g in registers (destroyed); lam is
  // left in the vector slot.
  __device__ __forceinline__ void solve_adj(unsigned pvec, float2 (&g)[NP]) const {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
#pragma unroll
      for (int j = 0; j < i; ++j) g[i] = cnmaj2(g[i], at(j, i), g[j]);
      const float2 di = at(i, i);
      g[i] = cmul2(g[i], f2(di.x, -di.y));
    }
#pragma unroll
    for (int i = NP - 2; i >= 0; --i)
#pragma unroll
      for (int j = i + 1; j < NP; ++j) g[i] = cnmaj2(g[i], at(j, i), g[j]);
#pragma unroll
    for (int i = 0; i < NP; ++i) vec((int)((pvec >> (3 * i)) & 7u)) = g[i];
  }
};

// The flagship shape only: pre = GAIN N x 1 and post = GAIN 1 x N both present (one input, one output channel;
// host-checked: plan->tpc_np).  Everything else runs on the row-distributed kernels.
template <int NP, bool BWD>
__global__ void __launch_bounds__(TPC_BLOCK, 6) fsweep_tpc_kernel(const __grid_constant__ ProgK P,
                                                                const __grid_constant__ LoopInfo L, const SweepArgs A,
                                                                int G) {
  extern __shared__ __align__(16) float2 tsm[];  // [NP*NP + NP][TPC_BLOCK]
  __shared__ __align__(16) float wfb[NP * NP];
  __shared__ float wpre[NP], wpost[NP];
  __shared__ __align__(8) uint64_t wbar;
  const int tid = threadIdx.x;
  const int N = P.rec_n;
  {
    // the block's feedback matrix W_fb enters shared memory through the TMA bulk-copy engine when it is a full,
    // 16-byte aligned NP x NP block (the headline config: 8 x 8 = 256 B; SASS UBLKCP + an mbarrier carrying the byte
    // count): one elected thread issues the copy, the gains are fetched meanwhile.  Smaller matrices are padded into
    // the NP x NP layout with plain loads.
    const OpK& fb = P.ops[L.fb];
    const bool bulk = fb.n_out == NP && fb.n_in == NP && (reinterpret_cast<uintptr_t>(fb.coef) & 15) == 0;
    if (bulk) {
      if (tid == 0) {
        mbar_init(&wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&wbar, (uint32_t)(NP * NP * sizeof(float)));
        tma_load_1d(wfb, fb.coef, (uint32_t)(NP * NP * sizeof(float)), &wbar);
      }
    } else {
      for (int e = tid; e < NP * NP; e += TPC_BLOCK) {
        const int r = e / NP, c = e - r * NP;
        wfb[e] = (r < fb.n_out && c < fb.n_in) ? __ldg(reinterpret_cast<const float*>(fb.coef) + r * fb.n_in + c) : 0.f;
      }
    }
    if (tid < NP) {
      wpre[tid] = tid < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.pre].coef) + tid) : 0.f;
      wpost[tid] = tid < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.post].coef) + tid) : 0.f;
    }
    __syncthreads();  // (also publishes the barrier's initialisation)
    if (bulk) mbar_wait(&wbar, 0);
  }
  TpcMat<NP> mat;
  mat.a = tsm + tid;
  mat.v = tsm + NP * NP * TPC_BLOCK + tid;

  const int ncols_total = A.batch * A.cols;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  double lacc = 0.0;

  // gradient accumulators (BWD): registers, across all bins of this thread
  float gwfb[BWD ? NP : 1][BWD ? NP : 1], gpre[NP], gpost[NP], gdiag[NP];
  if constexpr (BWD) {
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      gpre[m] = gpost[m] = gdiag[m] = 0.f;
#pragma unroll
      for (int j = 0; j < NP; ++j) gwfb[m][j] = 0.f;
    }
  }
  const OpK& ffop = P.ops[L.ff_begin];
  const bool want_ff = BWD && L.n_ff == 1 && ffop.acc_mode == ACC_SMEM;
  const bool ff_delay = ffop.kind == FSWEEP_OP_PDELAY;

  for (long long bl = (long long)blockIdx.x * TPC_BLOCK + tid; bl < A.n_bins; bl += (long long)gridDim.x * TPC_BLOCK) {
    const Ctx<float> ctx = make_ctx<float>(P, A.bin_begin + bl);
    // ---- diagonal chain D (runtime loop over the channels: one copy of the response code; unrolling by two was
    //      measured and bought nothing)
#pragma unroll 1
    for (int m = 0; m < NP; ++m) {
      cx<float> d = mk<float>(m < N ? 1.f : 0.f, 0.f);
      for (int i = 0; i < L.n_ff; ++i) {
        bool gd;
        d = cmul(d, op_diag<float>(P.ops[L.ff_begin + i], ctx, m, gd));
      }
      mat.vec(m) = f2(d.x, d.y);
    }
    float2 D[NP];
#pragma unroll
    for (int m = 0; m < NP; ++m) D[m] = mat.vec(m);
    // ---- A = I - D W  (rows >= N: identity)
#pragma unroll 1
    for (int m = 0; m < NP; ++m) {
      const float2 d = mat.vec(m);
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float w = wfb[m * NP + j];
        mat.at(m, j) = f2((m == j ? 1.f : 0.f) - d.x * w, -d.y * w);
      }
    }
    const unsigned pvec = mat.factor();

    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      // ---- y = A^-1 D (w_pre x)
      const cx<float> xv = ld_cx(x + (size_t)b * A.xbs + (size_t)bl * A.cols + cc);
#pragma unroll
      for (int m = 0; m < NP; ++m) mat.vec(m) = cmul2(D[m], f2(wpre[m] * xv.x, wpre[m] * xv.y));
      float2 y[NP];
      mat.solve(pvec, y);
      float ox = 0.f, oy = 0.f;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        ox = fmaf(wpost[m], y[m].x, ox);
        oy = fmaf(wpost[m], y[m].y, oy);
      }
      const size_t ooff = (size_t)bl * A.cols + cc;  // one output channel
      if constexpr (!BWD) {
        if (epi_fused(A.epilogue)) {
          const float e = abs_t(ox, oy) - __ldg(reinterpret_cast<const float*>(A.tgt) + (size_t)b * A.tbs + bl);
          lacc += (double)e * (double)e;
        } else if (A.epilogue == FSWEEP_EPI_ABS) {
          reinterpret_cast<float*>(A.y)[(size_t)b * A.ybs + ooff] = abs_t(ox, oy);
        } else {
          st_cx(reinterpret_cast<cx<float>*>(A.y) + (size_t)b * A.ybs + ooff, mk<float>(ox, oy));
        }
      } else {
        // ---- output gradient go
        float gox = 0.f, goy = 0.f;
        if (A.epilogue == FSWEEP_EPI_NONE) {
          const cx<float> g = ld_cx(reinterpret_cast<const cx<float>*>(A.gy) + (size_t)b * A.gybs + ooff);
          gox = g.x;
          goy = g.y;
        } else {
          const float mag = abs_t(ox, oy);
          float gabs;
          if (epi_fused(A.epilogue)) {
            const float e = mag - __ldg(reinterpret_cast<const float*>(A.tgt) + (size_t)b * A.tbs + bl);
            lacc += (double)e * (double)e;
            gabs = (float)(2.0 * A.crit_scale) * e;
          } else {
            gabs = __ldg(reinterpret_cast<const float*>(A.gy) + (size_t)b * A.gybs + ooff);
          }
          if (mag > 0.f) {
            const float t = gabs * rcp_t(mag);
            gox = t * ox;
            goy = t * oy;
          }
        }
        // ---- through the output gain: lam0 = w_post go, dw_post = Re(go conj(y))
        float2 g0[NP];
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          g0[m] = f2(wpost[m] * gox, wpost[m] * goy);
          gpost[m] = fmaf(gox, y[m].x, fmaf(goy, y[m].y, gpost[m]));
        }
        mat.solve_adj(pvec, g0);
        // ---- lam, g_u = conj(D) lam, u = s + W y; dW_fb = Re(g_u y^H); diagonal op; input gain
        float gxr = 0.f, gxi = 0.f;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          const float2 lam = mat.vec(m);
          const float2 gu = f2(D[m].x * lam.x + D[m].y * lam.y, D[m].x * lam.y - D[m].y * lam.x);
          float ux = wpre[m] * xv.x, uy = wpre[m] * xv.y;
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float w = wfb[m * NP + j];
            ux = fmaf(w, y[j].x, ux);
            uy = fmaf(w, y[j].y, uy);
            gwfb[m][j] = fmaf(gu.x, y[j].x, fmaf(gu.y, y[j].y, gwfb[m][j]));
          }
          if (want_ff) {
            // gh = lam conj(u);  PGAIN: Re gh;  PDELAY (fractional): Re(gh conj((ln g - j w) D))
            const float ghx = lam.x * ux + lam.y * uy, ghy = lam.y * ux - lam.x * uy;
            if (ff_delay) {
              const float tx = (float)ctx.lng * D[m].x + ctx.omega * D[m].y;
              const float ty = (float)ctx.lng * D[m].y - ctx.omega * D[m].x;
              gdiag[m] += ghx * tx + ghy * ty;
            } else {
              gdiag[m] += ghx;
            }
          }
          gpre[m] = fmaf(gu.x, xv.x, fmaf(gu.y, xv.y, gpre[m]));
          gxr = fmaf(wpre[m], gu.x, gxr);
          gxi = fmaf(wpre[m], gu.y, gxi);
        }
        if (A.gx != nullptr)
          st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)b * A.gxbs + (size_t)bl * A.cols + cc, mk<float>(gxr, gxi));
      }
    }
  }

  if constexpr (BWD) {
    // ---- block reduction of the register accumulators through the (now idle) matrix storage.  Staging layout is
    //      static (slot s at stage[s * TPC_BLOCK + tid]): s = m*NP + j for dW_fb[m][j], then NP slots each for
    //      dw_pre, dw_post and the diagonal op — every store has a constant offset.
    constexpr int S_PRE = NP * NP, S_POST = NP * NP + NP, S_DIAG = NP * NP + 2 * NP, S_TOTAL = NP * NP + 3 * NP;
    static_assert(S_TOTAL <= 2 * (NP * NP + NP), "staging does not fit the matrix storage");
    __syncthreads();
    float* stage = reinterpret_cast<float*>(tsm) + tid;
#pragma unroll
    for (int m = 0; m < NP; ++m) {
#pragma unroll
      for (int j = 0; j < NP; ++j) stage[(m * NP + j) * TPC_BLOCK] = gwfb[m][j];
      stage[(S_PRE + m) * TPC_BLOCK] = gpre[m];
      stage[(S_POST + m) * TPC_BLOCK] = gpost[m];
      stage[(S_DIAG + m) * TPC_BLOCK] = gdiag[m];
    }
    __syncthreads();
    const OpK& fbop = P.ops[L.fb];
    const OpK& preop = P.ops[L.pre];
    const OpK& postop = P.ops[L.post];
    float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    // partial[(op.row_off + i) * G + row] for entry i of row `row` (the layout fsweep_finalize_kernel reads; entries
    // outside the live shapes are never read, so they are not written either)
    for (int sidx = tid; sidx < S_TOTAL; sidx += TPC_BLOCK) {
      int dst = -1;
      if (sidx < S_PRE) {
        const int m = sidx / NP, j = sidx - m * NP;
        if (fbop.acc_mode == ACC_SMEM && m < N && j < N) dst = (fbop.row_off + j) * G + m;
      } else if (sidx < S_POST) {
        const int m = sidx - S_PRE;  // N x 1: row m, entry 0
        if (preop.acc_mode == ACC_SMEM && m < N) dst = preop.row_off * G + m;
      } else if (sidx < S_DIAG) {
        const int m = sidx - S_POST;  // 1 x N: row 0, entry m
        if (postop.acc_mode == ACC_SMEM && m < N) dst = (postop.row_off + m) * G;
      } else {
        const int m = sidx - S_DIAG;  // diagonal op: row m, entry 0
        if (want_ff && m < N) dst = ffop.row_off * G + m;
      }
      if (dst < 0) continue;
      const float* col = reinterpret_cast<const float*>(tsm) + (size_t)sidx * TPC_BLOCK;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int j = 0; j < TPC_BLOCK; j += 2) {  // rotated start: conflict-free
        s0 += col[(j + tid) & (TPC_BLOCK - 1)];
        s1 += col[(j + 1 + tid) & (TPC_BLOCK - 1)];
      }
      partial[dst] = s0 + s1;
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<float>(lacc, A.loss_partial);
}

constexpr size_t tpc_smem_bytes(int np) { return (size_t)(np * np + np) * TPC_BLOCK * sizeof(float2); }

cudaError_t launch_tpc(int np, bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                       int G);

}  // namespace fsweep
