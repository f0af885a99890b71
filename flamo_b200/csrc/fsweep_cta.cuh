// fsweep_cta.cuh — CTA-per-bin sweep for WIDE FDN loops (32 < N <= 64, float32; BASELINE config 5: 64 x 64, batch 32).
//
// Same pattern and math as fsweep_loop.cuh / fsweep_tpc.cuh
//     [GAIN N x 1]  ->  RECURSION( diagonal chain ; one real N x N matrix )  ->  [GAIN 1 x N]
// but one thread BLOCK (256 threads) owns one frequency bin, with the per-bin matrix A = I - D(w) W and up to 32
// right-hand sides resident in shared memory (33 KB + 17 KB), the way a batched small-matrix LAPACK would do it:
//   * P A = L U: right-looking, the (63-k)^2 trailing update of step k spread over all 256 threads;
//   * all batch items of the bin are solved TOGETHER (forward / backward substitution on a 64 x 32 block), where the
//     row-distributed path (two warps per bin, fsweep_kernels.cuh with G = 64) solved them four at a time with two
//     named barriers per substitution step;
//   * the output has ONE channel, so the adjoint solve is shared by the whole batch: lambda_b = g_b v with
//     v = A^-H w_post solved once per bin, and the feedback-matrix gradient collapses to ONE outer product per bin,
//         dW = Re( (conj(D) v) (sum_b g_b conj(y_b))^T ),
//     accumulated in 16 registers per thread across all bins of the block: no atomics at all (the row-distributed
//     path issued 4096 global atomics per bin and batch chunk — 2.5e10 per step of config 5);
//   * gradients leave the block once, in the `partial` layout fsweep_finalize_kernel sums in float64.
// Gradients of the diagonal chain (learnable delays / gains inside the loop) are not formed here: such plans stay on
// the row-distributed kernels (host-checked, plan->cta).
#pragma once
#include "fsweep_tpc.cuh"

namespace fsweep {

constexpr int CTA_T = 256;    // threads per block
constexpr int CN = 64;        // padded loop width
constexpr int CLD = CN + 1;   // row stride of A in float2 (odd: column walks are conflict-free)
constexpr int CQ = 32;        // right-hand sides per pass
constexpr int CYLD = CQ + 1;  // row stride of Y

__device__ __forceinline__ float2 cmulj2(float2 a, float2 b) {  // conj(a) * b
  return f2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}

constexpr size_t cta_smem_bytes() {
  return (size_t)(CN * CLD + CN * CYLD + 4 * CN + 2 * CQ + 8 * CQ) * sizeof(float2) + (size_t)(2 * CN) * sizeof(float) +
         (size_t)(CN + 8) * sizeof(int);
}

template <bool BWD>
__global__ void __launch_bounds__(CTA_T, 4) fsweep_cta_kernel(const __grid_constant__ ProgK P,
                                                            const __grid_constant__ LoopInfo L, const SweepArgs A, int G) {
  extern __shared__ __align__(16) float2 csm[];
  float2* sA = csm;                 // [CN][CLD]   L\U in place, 1/U_kk on the diagonal
  float2* sY = sA + CN * CLD;       // [2][CQ]     pivot-entry broadcast of the substitutions (rest: spare)
  float2* sD = sY + CN * CYLD;      // [CN]  diagonal chain response
  float2* sV = sD + CN;             // [CN]  adjoint vector v, then vd = conj(D) v
  float2* sT = sV + CN;             // [CN]  T[j] = sum_b g_b conj(y_b[j])
  float2* sW = sT + CN;             // [CN]  scratch of the adjoint solve
  float2* sGo = sW + CN;            // [CQ]  output gradients g_b of the pass
  float2* sX = sGo + CQ;            // [CQ]  inputs x_b of the pass
  float2* sRed = sX + CQ;           // [8][CQ] cross-warp partial sums
  float* sWpre = reinterpret_cast<float*>(sRed + 8 * CQ);  // [CN]
  float* sWpost = sWpre + CN;                               // [CN]
  int* sPiv = reinterpret_cast<int*>(sWpost + CN);          // [CN] row map of P A; [CN..]: scalars
  int* sScalar = sPiv + CN;

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int N = P.rec_n;
  const OpK& fbop = P.ops[L.fb];
  const float* Wfb = reinterpret_cast<const float*>(fbop.coef);
  if (t < CN) {
    sWpre[t] = t < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.pre].coef) + t) : 0.f;
    sWpost[t] = t < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.post].coef) + t) : 0.f;
  }
  const int Q = A.batch * A.cols;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  double lacc = 0.0;
  float gw[BWD ? 16 : 1];  // dW_fb entries e = t + 256 i  (m = e >> 6, j = e & 63)
  float gpre = 0.f, gpost = 0.f;
  if constexpr (BWD) {
#pragma unroll
    for (int i = 0; i < 16; ++i) gw[i] = 0.f;
  }
  __syncthreads();

  for (long long bl = blockIdx.x; bl < A.n_bins; bl += gridDim.x) {
    // ---- 1. diagonal chain, row map
    if (t < CN) {
      const Ctx<float> ctx = make_ctx<float>(P, A.bin_begin + bl);
      cx<float> d = mk<float>(t < N ? 1.f : 0.f, 0.f);
      for (int i = 0; i < L.n_ff; ++i) {
        bool gd;
        d = cmul(d, op_diag<float>(P.ops[L.ff_begin + i], ctx, t, gd));
      }
      sD[t] = f2(d.x, d.y);
      sPiv[t] = t;
    }
    __syncthreads();
    // ---- 2. A = I - D W (rows / columns >= N: identity)
    for (int e = t; e < CN * CN; e += CTA_T) {
      const int m = e >> 6, j = e & 63;
      const float w = (m < N && j < N) ? __ldg(Wfb + m * N + j) : 0.f;
      const float2 d = sD[m];
      sA[m * CLD + j] = f2((m == j ? 1.f : 0.f) - d.x * w, -d.y * w);
    }
    __syncthreads();
    // ---- 3. P A = L U
    for (int k = 0; k < CN; ++k) {
      if (warp == 0) {  // pivot search in column k, rows k .. 63
        float best = -1.f;
        int pr = k;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = k + lane + 32 * h;
          if (r < CN) {
            const float2 c = sA[r * CLD + k];
            const float mg = c.x * c.x + c.y * c.y;
            if (mg > best) {
              best = mg;
              pr = r;
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(FULL, best, o);
          const int op_ = __shfl_xor_sync(FULL, pr, o);
          if (ob > best || (ob == best && op_ < pr)) {
            best = ob;
            pr = op_;
          }
        }
        if (lane == 0) sScalar[0] = pr;
      }
      __syncthreads();
      const int pr = sScalar[0];
      if (pr != k) {  // uniform
        if (t < CN) {
          const float2 a = sA[k * CLD + t];
          sA[k * CLD + t] = sA[pr * CLD + t];
          sA[pr * CLD + t] = a;
        } else if (t == CN) {
          const int a = sPiv[k];
          sPiv[k] = sPiv[pr];
          sPiv[pr] = a;
        }
        __syncthreads();
      }
      const float2 d = sA[k * CLD + k];
      const float id = rcp_t(d.x * d.x + d.y * d.y);
      const float2 inv = f2(d.x * id, -d.y * id);
      if (t < CN - 1 - k) {
        const int r = k + 1 + t;
        sA[r * CLD + k] = cmul2(sA[r * CLD + k], inv);
      }
      __syncthreads();
      if (t == 0) sA[k * CLD + k] = inv;  // nobody reads the diagonal during the update
      {
        // trailing update: thread (ty, tx) owns rows k+1+ty+16i and columns k+1+tx+16jj; its (up to four) pivot-row
        // entries are loaded once per step
        const int ty = t >> 4, tx = t & 15;
        const int j0 = k + 1 + tx;
        float2 u[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) u[jj] = (j0 + 16 * jj < CN) ? sA[k * CLD + j0 + 16 * jj] : f2(0.f, 0.f);
        for (int r = k + 1 + ty; r < CN; r += 16) {
          const float2 l = sA[r * CLD + k];
          float2* row = sA + r * CLD + j0;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            if (j0 + 16 * jj < CN) row[16 * jj] = cnma2(row[16 * jj], l, u[jj]);
        }
      }
      __syncthreads();
    }

    if constexpr (BWD) {
      // ---- 4. v = A^-H w_post, once per bin (threads 0..63, named barrier 1): A^H = U^H L^H P
      if (t < CN) {
        float2 g = f2(sWpost[t], 0.f);
        // U^H w = g  (lower triangular, column-oriented: after w_i is final, g_j -= conj(U[i][j]) w_i for j > i)
        for (int i = 0; i < CN; ++i) {
          if (t == i) {
            const float2 di = sA[i * CLD + i];
            g = cmul2(g, f2(di.x, -di.y));
            sW[i] = g;
          }
          asm volatile("bar.sync 1, 64;" ::: "memory");
          if (t > i) {
            const float2 u = sA[i * CLD + t];
            const float2 wi = sW[i];
            g.x -= u.x * wi.x + u.y * wi.y;  // conj(u) * wi
            g.y -= u.x * wi.y - u.y * wi.x;
          }
        }
        // L^H z = w  (unit upper triangular): z_i final when all j > i are done; z_j -= conj(L[i][j]) z_i for j < i
        for (int i = CN - 1; i >= 0; --i) {
          if (t == i) sW[i] = g;
          asm volatile("bar.sync 1, 64;" ::: "memory");
          if (t < i) {
            const float2 l = sA[i * CLD + t];
            const float2 zi = sW[i];
            g.x -= l.x * zi.x + l.y * zi.y;
            g.y -= l.x * zi.y - l.y * zi.x;
          }
        }
        // v[piv[i]] = z_i ; vd = conj(D) v
        sV[sPiv[t]] = g;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        const float2 v = sV[t];
        const float2 dd = sD[t];
        asm volatile("bar.sync 1, 64;" ::: "memory");
        sV[t] = f2(dd.x * v.x + dd.y * v.y, dd.x * v.y - dd.y * v.x);
        sT[t] = f2(0.f, 0.f);
      }
      if (t == 0) {
        sRed[0] = f2(0.f, 0.f);  // unused here; keeps the scratch initialised
      }
      __syncthreads();
    }
    float2 xbar = f2(0.f, 0.f);  // sum_b g_b conj(x_b) (thread 0)

    // ---- passes of up to CQ right-hand sides
    for (int q0 = 0; q0 < Q; q0 += CQ) {
      const int nq = min(CQ, Q - q0);
      if (t < CQ) {
        float2 xv = f2(0.f, 0.f);
        if (t < nq) {
          const int q = q0 + t, b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
          const cx<float> v = ld_cx(x + (size_t)b * A.xbs + (size_t)bl * A.cols + cc);
          xv = f2(v.x, v.y);
        }
        sX[t] = xv;
      }
      __syncthreads();
      // A warp owns rows ty, ty + 8, ..., ty + 56 and a lane one right-hand side: the thread's eight entries of the
      // solution block stay in REGISTERS through both substitutions; only the pivot entry of each step crosses
      // shared memory (double buffered: one barrier per step).
      const int ty = warp, b = lane;
      float2 yr[8];
      {
        const float2 xb = sX[b];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // right-hand side, already row-permuted: Y'[r][b] = D[p_r] w_pre[p_r] x_b
          const int src = sPiv[ty + 8 * i];
          const float2 d = sD[src];
          const float w = sWpre[src];
          yr[i] = cmul2(f2(d.x * w, d.y * w), xb);
        }
      }
      float2* sBk = sY;  // [2][CQ] pivot-entry broadcast
      // forward substitution (unit lower): step k = 8p + kk is owned by warp kk, register p
#pragma unroll
      for (int p = 0; p < 8; ++p) {
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {
          const int k = 8 * p + kk;
          if (ty == kk) sBk[(k & 1) * CQ + b] = yr[p];
          __syncthreads();
          const float2 yk = sBk[(k & 1) * CQ + b];
          if (ty > kk) yr[p] = cnma2(yr[p], sA[(ty + 8 * p) * CLD + k], yk);
#pragma unroll
          for (int i = p + 1; i < 8; ++i) yr[i] = cnma2(yr[i], sA[(ty + 8 * i) * CLD + k], yk);
        }
      }
      // backward substitution (upper, reciprocal diagonal)
#pragma unroll
      for (int p = 7; p >= 0; --p) {
#pragma unroll 1
        for (int kk = 7; kk >= 0; --kk) {
          const int k = 8 * p + kk;
          if (ty == kk) {  // (buffer parity flipped w.r.t. the forward pass: its last step used buffer 1)
            yr[p] = cmul2(yr[p], sA[k * CLD + k]);
            sBk[((k + 1) & 1) * CQ + b] = yr[p];
          }
          __syncthreads();
          const float2 xk = sBk[((k + 1) & 1) * CQ + b];
          if (ty < kk) yr[p] = cnma2(yr[p], sA[(ty + 8 * p) * CLD + k], xk);
#pragma unroll
          for (int i = 0; i < p; ++i) yr[i] = cnma2(yr[i], sA[(ty + 8 * i) * CLD + k], xk);
        }
      }
      // ---- output o_b = w_post . y_b
      {
        float ox = 0.f, oy = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float w = sWpost[ty + 8 * i];
          ox = fmaf(w, yr[i].x, ox);
          oy = fmaf(w, yr[i].y, oy);
        }
        sRed[ty * CQ + b] = f2(ox, oy);
      }
      __syncthreads();
      if (t < CQ) {
        float ox = 0.f, oy = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          ox += sRed[w * CQ + t].x;
          oy += sRed[w * CQ + t].y;
        }
        float2 go = f2(0.f, 0.f);
        if (t < nq) {
          const int q = q0 + t, bb = (A.cols == 1) ? q : q / A.cols, cc = q - bb * A.cols;
          const size_t ooff = (size_t)bl * A.cols + cc;
          if constexpr (!BWD) {
            if (epi_fused(A.epilogue)) {
              const float e = abs_t(ox, oy) - __ldg(reinterpret_cast<const float*>(A.tgt) + (size_t)bb * A.tbs + bl);
              lacc += (double)e * (double)e;
            } else if (A.epilogue == FSWEEP_EPI_ABS) {
              reinterpret_cast<float*>(A.y)[(size_t)bb * A.ybs + ooff] = abs_t(ox, oy);
            } else {
              st_cx(reinterpret_cast<cx<float>*>(A.y) + (size_t)bb * A.ybs + ooff, mk<float>(ox, oy));
            }
          } else {
            if (A.epilogue == FSWEEP_EPI_NONE) {
              const cx<float> g = ld_cx(reinterpret_cast<const cx<float>*>(A.gy) + (size_t)bb * A.gybs + ooff);
              go = f2(g.x, g.y);
            } else {
              const float mag = abs_t(ox, oy);
              float gabs;
              if (epi_fused(A.epilogue)) {
                const float e = mag - __ldg(reinterpret_cast<const float*>(A.tgt) + (size_t)bb * A.tbs + bl);
                lacc += (double)e * (double)e;
                gabs = (float)(2.0 * A.crit_scale) * e;
              } else {
                gabs = __ldg(reinterpret_cast<const float*>(A.gy) + (size_t)bb * A.gybs + ooff);
              }
              if (mag > 0.f) {
                const float s = gabs * rcp_t(mag);
                go = f2(s * ox, s * oy);
              }
            }
          }
        }
        sGo[t] = go;
      }
      __syncthreads();
      if constexpr (BWD) {
        // T[j] += sum_b g_b conj(y_b[j]);  xbar += sum_b g_b conj(x_b);  g_x = g_b * sum_m w_pre[m] vd[m]
        const float2 go = sGo[b];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = ty + 8 * i;
          const float2 y = yr[i];
          float vx = go.x * y.x + go.y * y.y, vy = go.y * y.x - go.x * y.y;  // g conj(y)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            vx += __shfl_xor_sync(FULL, vx, o);
            vy += __shfl_xor_sync(FULL, vy, o);
          }
          if (lane == 0) {
            float2 acc = sT[j];
            acc.x += vx;
            acc.y += vy;
            sT[j] = acc;  // row j belongs to warp j & 7 only
          }
        }
        if (warp == 0) {
          const float2 xv = sX[b];
          float vx = go.x * xv.x + go.y * xv.y, vy = go.y * xv.x - go.x * xv.y;
          float sx = 0.f, sy = 0.f;  // S = sum_m w_pre[m] vd[m], two rows per lane
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float w = sWpre[lane + 32 * h];
            sx = fmaf(w, sV[lane + 32 * h].x, sx);
            sy = fmaf(w, sV[lane + 32 * h].y, sy);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            vx += __shfl_xor_sync(FULL, vx, o);
            vy += __shfl_xor_sync(FULL, vy, o);
            sx += __shfl_xor_sync(FULL, sx, o);
            sy += __shfl_xor_sync(FULL, sy, o);
          }
          xbar.x += vx;
          xbar.y += vy;
          if (A.gx != nullptr && lane < nq) {
            const int q = q0 + lane, bb = (A.cols == 1) ? q : q / A.cols, cc = q - bb * A.cols;
            st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)bb * A.gxbs + (size_t)bl * A.cols + cc,
                  mk<float>(go.x * sx - go.y * sy, go.x * sy + go.y * sx));
          }
        }
        __syncthreads();
      }
    }

    if constexpr (BWD) {
      // ---- 5. this bin's gradient contributions:  dW[m][j] += Re(vd[m] T[j]),  dw_post[m] += Re T[m],
      //         dw_pre[m] += Re(vd[m] xbar)
      if (t == 0) sRed[0] = xbar;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int e = t + CTA_T * i, m = e >> 6, j = e & 63;
        const float2 vd = sV[m], tj = sT[j];
        gw[i] = fmaf(vd.x, tj.x, fmaf(-vd.y, tj.y, gw[i]));
      }
      if (t < CN) {
        const float2 vd = sV[t], xb = sRed[0];
        gpost += sT[t].x;
        gpre = fmaf(vd.x, xb.x, fmaf(-vd.y, xb.y, gpre));
      }
      __syncthreads();
    }
  }

  if constexpr (BWD) {
    // partial[(op.row_off + i) * G + row] for entry i of row `row` (fsweep_finalize_kernel's layout)
    float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    const OpK& preop = P.ops[L.pre];
    const OpK& postop = P.ops[L.post];
    if (fbop.acc_mode == ACC_SMEM) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int e = t + CTA_T * i, m = e >> 6, j = e & 63;
        if (m < N && j < N) partial[(fbop.row_off + j) * G + m] = gw[i];
      }
    }
    if (t < N) {
      if (preop.acc_mode == ACC_SMEM) partial[preop.row_off * G + t] = gpre;       // N x 1: row t, entry 0
      if (postop.acc_mode == ACC_SMEM) partial[(postop.row_off + t) * G] = gpost;  // 1 x N: row 0, entry t
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<float>(lacc, A.loss_partial);
}

cudaError_t launch_cta(bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G);
cudaError_t occupancy_cta(bool bwd, int* blocks_per_sm);

}  // namespace fsweep
