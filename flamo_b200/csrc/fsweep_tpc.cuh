// fsweep_tpc.cuh — COMPACT thread-per-bin sweep for small FDN loops (width <= 8, float32).
//
// Same pattern and math as fsweep_loop.cuh / fsweep_tpb.cuh
//     [GAIN N x 1]  ->  RECURSION( diagonal chain ; one real N x N matrix )  ->  [GAIN 1 x N]
// with ONE THREAD per frequency bin, but built around what the ncu captures of the unrolled thread-per-bin kernels
// showed (profiles/r01e_ncu_full_sweep_kernels.md): at M ~ 5e4 bins every thread owns ONE bin, so every instruction
// of the kernel is executed once per warp, cold, and the kernel runs at the speed of INSTRUCTION FETCH
// (131 - 217 KB of unrolled SASS, `stalled_no_instruction` = 70 % of the time).  Here
//   * the per-bin matrix A = I - D(w) W lives in thread-private SHARED-MEMORY columns (A[i][j] of thread t at
//     float2 index (i*NP + j)*BLOCK + t: conflict-free), so that the LU factorisation and the four triangular
//     solves are RUNTIME loops — a few hundred SASS instructions that stay in the instruction cache;
//   * 576 B of shared memory per thread and <= 168 registers let 6 blocks of 64 threads live on an SM:
//     56 832 resident threads cover the 48 001 bins of the headline config in ONE wave (the unrolled backward
//     kernel needed 255 registers -> 4 blocks -> a two-wave tail);
//   * gradient accumulators (W_fb: N x N, the two gain vectors, the diagonal op) stay in registers across bins and
//     are reduced once per block through the then-idle matrix storage into the `partial` layout shared with the
//     other kernel families (fsweep_finalize_kernel).
// LU convention (LINPACK, as fsweep_tpb.cuh): at step k rows k and p_k are swapped in columns >= k only, the
// multipliers stay where they were produced, 1/U_kk is stored on the diagonal, and the solves replay the
// interchanges progressively.
#pragma once
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

  // in-place LU with partial pivoting; returns the packed interchange word (3 bits per step)
  __device__ __forceinline__ unsigned factor() const {
    unsigned perm = 0u;
#pragma unroll 1
    for (int k = 0; k < NP; ++k) {
      float2 d = at(k, k);
      float best = d.x * d.x + d.y * d.y;
      int pr = k;
#pragma unroll 1
      for (int r = k + 1; r < NP; ++r) {
        const float2 c = at(r, k);
        const float m = c.x * c.x + c.y * c.y;
        if (m > best) {
          best = m;
          pr = r;
        }
      }
      perm |= (unsigned)pr << (3 * k);
      if (pr != k) {
#pragma unroll 1
        for (int j = k; j < NP; ++j) {
          const float2 t = at(k, j);
          at(k, j) = at(pr, j);
          at(pr, j) = t;
        }
      }
      float2 prow[NP];  // pivot row in registers (entries j <= k are loaded but unused)
#pragma unroll
      for (int j = 0; j < NP; ++j) prow[j] = at(k, j);
      d = at(k, k);
      const float id = rcp_t(d.x * d.x + d.y * d.y);
      const float2 inv = f2(d.x * id, -d.y * id);
      at(k, k) = inv;
#pragma unroll 1
      for (int r = k + 1; r < NP; ++r) {
        const float2 l = cmul2(at(r, k), inv);
        at(r, k) = l;
#pragma unroll
        for (int j = 1; j < NP; ++j)
          if (j > k) at(r, j) = cnma2(at(r, j), l, prow[j]);
      }
    }
    return perm;
  }

  // A x = b, b and x in the vector slot
  __device__ __forceinline__ void solve(unsigned perm) const {
#pragma unroll 1
    for (int k = 0; k < NP; ++k) {
      const int pr = (int)((perm >> (3 * k)) & 7u);
      const float2 vk = vec(pr);
      vec(pr) = vec(k);
      vec(k) = vk;
#pragma unroll 1
      for (int r = k + 1; r < NP; ++r) vec(r) = cnma2(vec(r), at(r, k), vk);
    }
#pragma unroll 1
    for (int k = NP - 1; k >= 0; --k) {
      float2 acc = vec(k);
#pragma unroll 1
      for (int j = k + 1; j < NP; ++j) acc = cnma2(acc, at(k, j), vec(j));
      vec(k) = cmul2(acc, at(k, k));
    }
  }

  // A^H x = g:  w = U^-H g, then for k = n-1 .. 0:  w <- P_k (M_k^H w)
  __device__ __forceinline__ void solve_adj(unsigned perm) const {
#pragma unroll 1
    for (int i = 0; i < NP; ++i) {
      float2 acc = vec(i);
#pragma unroll 1
      for (int j = 0; j < i; ++j) acc = cnmaj2(acc, at(j, i), vec(j));
      const float2 di = at(i, i);
      vec(i) = cmul2(acc, f2(di.x, -di.y));
    }
#pragma unroll 1
    for (int k = NP - 1; k >= 0; --k) {
      float2 acc = vec(k);
#pragma unroll 1
      for (int r = k + 1; r < NP; ++r) acc = cnmaj2(acc, at(r, k), vec(r));
      const int pr = (int)((perm >> (3 * k)) & 7u);
      const float2 t = vec(pr);  // pr >= k; for pr == k this is the stale value and is overwritten below
      vec(pr) = acc;
      if (pr != k) vec(k) = t;
    }
  }
};

// The flagship shape only: pre = GAIN N x 1 and post = GAIN 1 x N both present (one input, one output channel;
// host-checked: plan->tpc_np).  Everything else runs on the row-distributed kernels.
template <int NP, bool BWD>
__global__ void __launch_bounds__(TPC_BLOCK, 6) fsweep_tpc_kernel(const __grid_constant__ ProgK P,
                                                                const __grid_constant__ LoopInfo L, const SweepArgs A,
                                                                int G) {
  extern __shared__ __align__(16) float2 tsm[];  // [NP*NP + NP][TPC_BLOCK]
  __shared__ float wfb[NP * NP], wpre[NP], wpost[NP];
  const int tid = threadIdx.x;
  const int N = P.rec_n;
  {
    const OpK& fb = P.ops[L.fb];
    for (int e = tid; e < NP * NP; e += TPC_BLOCK) {
      const int r = e / NP, c = e - r * NP;
      wfb[e] = (r < fb.n_out && c < fb.n_in) ? __ldg(reinterpret_cast<const float*>(fb.coef) + r * fb.n_in + c) : 0.f;
    }
    if (tid < NP) {
      wpre[tid] = tid < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.pre].coef) + tid) : 0.f;
      wpost[tid] = tid < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.post].coef) + tid) : 0.f;
    }
    __syncthreads();
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
    // ---- diagonal chain D (runtime loop over the channels: one copy of the response code)
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
    const unsigned perm = mat.factor();

    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      // ---- y = A^-1 D (w_pre x)
      const cx<float> xv = ld_cx(x + (size_t)b * A.xbs + (size_t)bl * A.cols + cc);
#pragma unroll
      for (int m = 0; m < NP; ++m) mat.vec(m) = cmul2(D[m], f2(wpre[m] * xv.x, wpre[m] * xv.y));
      mat.solve(perm);
      float2 y[NP];
      float ox = 0.f, oy = 0.f;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        y[m] = mat.vec(m);
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
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          mat.vec(m) = f2(wpost[m] * gox, wpost[m] * goy);
          gpost[m] = fmaf(gox, y[m].x, fmaf(goy, y[m].y, gpost[m]));
        }
        mat.solve_adj(perm);
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
    // ---- block reduction of the register accumulators through the (now idle) matrix storage:
    //      stage[slot * TPC_BLOCK + tid], slot = flat accumulator index op.acc_off + row*row_len + e
    __syncthreads();
    float* stage = reinterpret_cast<float*>(tsm);  // 2 * (NP*NP + NP) floats per thread >= NP*NP + 3*NP slots
    const OpK& fbop = P.ops[L.fb];
    const OpK& preop = P.ops[L.pre];
    const OpK& postop = P.ops[L.post];
    if (fbop.acc_mode == ACC_SMEM) {
#pragma unroll
      for (int m = 0; m < NP; ++m)
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (m < fbop.n_out && j < fbop.n_in) stage[(fbop.acc_off + m * fbop.row_len + j) * TPC_BLOCK + tid] = gwfb[m][j];
    }
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      if (m < N) {
        if (preop.acc_mode == ACC_SMEM) stage[(preop.acc_off + m) * TPC_BLOCK + tid] = gpre[m];     // N x 1: row m
        if (postop.acc_mode == ACC_SMEM) stage[(postop.acc_off + m) * TPC_BLOCK + tid] = gpost[m];  // 1 x N: entry m
        if (want_ff) stage[(ffop.acc_off + m) * TPC_BLOCK + tid] = gdiag[m];
      }
    }
    __syncthreads();
    float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    for (int e = tid; e < P.acc_per_lane * G; e += TPC_BLOCK) partial[e] = 0.f;
    __syncthreads();
    for (int opi = 0; opi < P.n_ops; ++opi) {
      const OpK& op = P.ops[opi];
      if (op.acc_mode != ACC_SMEM) continue;
      const int total = op.n_out * op.row_len;
      for (int e = tid; e < total; e += TPC_BLOCK) {
        const int row = e / op.row_len, i = e - row * op.row_len;
        const float* col = stage + (size_t)(op.acc_off + e) * TPC_BLOCK;
        float sum = 0.f;
        for (int j = 0; j < TPC_BLOCK; ++j) sum += col[(j + tid) & (TPC_BLOCK - 1)];  // rotated: conflict-free
        partial[(op.row_off + i) * G + row] = sum;
      }
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<float>(lacc, A.loss_partial);
}

constexpr size_t tpc_smem_bytes(int np) { return (size_t)(np * np + np) * TPC_BLOCK * sizeof(float2); }

cudaError_t launch_tpc(int np, bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                       int G);

}  // namespace fsweep
