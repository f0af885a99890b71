// fsweep_tpr.cuh — thread-per-bin sweep of the flagship FDN shape with the loop matrix in REGISTERS (width <= 8, float32).
//
// Same pattern, math and accumulator layout as fsweep_tpc.cuh
//     [GAIN N x 1]  ->  RECURSION( diagonal chain ; one real N x N matrix )  ->  [GAIN 1 x N]
// but built from what the ncu capture of that kernel showed (profiles/r02h_ncu_full_tpc_bwd.md): 6026 warp instructions
// per bin, 1210 of them shared-memory loads / stores of the per-bin matrix, ~560 predicated-off slots of its runtime
// elimination loops, issue slots 40 % busy at 2.5 warps per scheduler, short-scoreboard (LDS) the top stall.  Here
//   * A = I - D(w) W lives in 128 registers and the elimination is fully unrolled, which is only possible WITHOUT row
//     interchanges — so pivoting is THRESHOLD pivoting: step k keeps the diagonal pivot unless it is smaller than
//     TPR_TAU times the largest candidate below it.  For loop matrices I - D W with |D| < 1 and W orthogonal (every FDN)
//     the diagonal never gets that small (numpy on config 2, all 48 001 bins: no interchange at tau = 0.25, element growth
//     1.9), so the common path is straight-line code on registers.  A bin that does need an interchange finishes its
//     factorisation in a slow path with full partial pivoting on a local-memory copy (runtime loops, one copy of the
//     code); the solves apply the row map through shared memory only then;
//   * the gradient accumulators live in shared memory (88 per thread), written — not added — by a thread's first bin,
//     which at ~one bin per thread is its only bin; the forward solution is parked in shared memory during the adjoint
//     solve, so nothing but the matrix and one vector is live in registers at a time (184-register cap);
//   * ONE block of 352 threads per SM: 148 x 352 = 52 096 threads hold the 48 001 bins of the headline config in a
//     single wave, and the gradient finalize kernel reads 137 per-block partial rows instead of 751.
#pragma once
#include <type_traits>

#include "fsweep_tpc.cuh"

namespace fsweep {

constexpr int TPR_BLOCK = 352;      // 11 warps; 65536 / 352 = 186 -> 184 registers per thread
constexpr int TPR_SS = TPR_BLOCK + 1;  // accumulator row stride (odd: a thread summing ROW s walks banks s, s + 1, ... — no rotation needed)
constexpr float TPR_TAU2 = 0.0625f;  // interchange when |a_kk|^2 < tau^2 max_r |a_rk|^2, tau = 0.25

template <int NP>
struct TprSmem {
  static constexpr int S_PRE = NP * NP, S_POST = NP * NP + NP, S_DIAG = NP * NP + 2 * NP, S_TOTAL = NP * NP + 3 * NP;
  // [stage: S_TOTAL][TPR_SS] float | [dslot | yslot | vslot : NP][BLOCK] float2 | red [4][S_TOTAL] float
  static constexpr size_t stage_bytes = (size_t)S_TOTAL * TPR_SS * 4;
  static_assert(stage_bytes % 8 == 0, "float2 arrays follow the accumulators");
  static constexpr size_t bytes = stage_bytes + (size_t)3 * NP * TPR_BLOCK * 8 + (size_t)4 * S_TOTAL * 4;
};

// Finish P A = L U from step k0 on with full partial pivoting (same conventions as TpcMat::factor: full row interchanges,
// 1 / U_kk on the diagonal, row map packed 3 bits per row).  t is dynamically indexed: local memory.
template <int NP>
__device__ __noinline__ void tpr_factor_slow(float2* t, int k0, unsigned* pvec_io) {
  unsigned pvec = *pvec_io;
#pragma unroll 1
  for (int k = k0; k < NP; ++k) {
    float best = -1.f;
    int pr = k;
#pragma unroll 1
    for (int r = k; r < NP; ++r) {
      const float2 c = t[r * NP + k];
      const float m = c.x * c.x + c.y * c.y;
      if (m > best) {
        best = m;
        pr = r;
      }
    }
    if (pr != k) {
#pragma unroll 1
      for (int j = 0; j < NP; ++j) {
        const float2 s = t[k * NP + j];
        t[k * NP + j] = t[pr * NP + j];
        t[pr * NP + j] = s;
      }
      const unsigned fx = ((pvec >> (3 * k)) ^ (pvec >> (3 * pr))) & 7u;
      pvec ^= (fx << (3 * k)) | (fx << (3 * pr));
    }
    const float2 d = t[k * NP + k];
    const float id = rcp_t(d.x * d.x + d.y * d.y);
    const float2 inv = f2(d.x * id, -d.y * id);
    t[k * NP + k] = inv;
#pragma unroll 1
    for (int r = k + 1; r < NP; ++r) {
      const float2 l = cmul2(t[r * NP + k], inv);
      t[r * NP + k] = l;
#pragma unroll 1
      for (int j = k + 1; j < NP; ++j) t[r * NP + j] = cnma2(t[r * NP + j], l, t[k * NP + j]);
    }
  }
  *pvec_io = pvec;
}

template <int NP>
__device__ __forceinline__ constexpr unsigned tpr_ident() {
  unsigned p = 0u;
  for (int i = 0; i < NP; ++i) p |= (unsigned)i << (3 * i);
  return p;
}

template <int NP, bool BWD>
__global__ void __launch_bounds__(TPR_BLOCK, 1) fsweep_tpr_kernel(const __grid_constant__ ProgK P,
                                                                const __grid_constant__ LoopInfo L, const SweepArgs A,
                                                                int G) {
  using SM = TprSmem<NP>;
  constexpr int S_PRE = SM::S_PRE, S_POST = SM::S_POST, S_DIAG = SM::S_DIAG, S_TOTAL = SM::S_TOTAL;
  constexpr unsigned IDENT = tpr_ident<NP>();
  extern __shared__ __align__(16) unsigned char tpr_raw[];
  __shared__ __align__(16) float wfb[NP * NP];
  __shared__ float wpre[NP], wpost[NP];
  __shared__ __align__(8) uint64_t wbar;
  __shared__ int sdi[NP];     // integer-delay fast path: the delays in samples ...
  __shared__ float smag[NP];  // ... and gamma^delay (bin-invariant)
  const int tid = threadIdx.x;
  const int N = P.rec_n;
  const OpK& ffop = P.ops[L.ff_begin];
  bool fastd;
  pdl_sync();  // (fsweep_pdl.cuh) the blocks were scheduled while the preceding kernel of the step was still running
  float* stage = reinterpret_cast<float*>(tpr_raw) + tid;  // slot s at stage[s * TPR_SS]
  float2* dslot = reinterpret_cast<float2*>(tpr_raw + SM::stage_bytes) + tid;  // entry i at [i * TPR_BLOCK]
  float2* yslot = dslot + NP * TPR_BLOCK;
  float2* vslot = yslot + NP * TPR_BLOCK;
  {
    // W_fb through the TMA bulk-copy engine when it is a full, 16-byte aligned NP x NP block (as in fsweep_tpc.cuh)
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
      for (int e = tid; e < NP * NP; e += TPR_BLOCK) {
        const int r = e / NP, c = e - r * NP;
        wfb[e] = (r < fb.n_out && c < fb.n_in) ? __ldg(reinterpret_cast<const float*>(fb.coef) + r * fb.n_in + c) : 0.f;
      }
    }
    if (tid < NP) {
      wpre[tid] = tid < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.pre].coef) + tid) : 0.f;
      wpost[tid] = tid < N ? __ldg(reinterpret_cast<const float*>(P.ops[L.post].coef) + tid) : 0.f;
    }
    // The usual FDN chain is ONE parallelDelay with integer delays: gamma^d does not depend on the bin and the phase index
    // (k d) mod nfft fits 32-bit integer arithmetic — no float64 range reduction, no expf per bin and channel.  One thread
    // per channel prepares (d, gamma^d); the block agrees on the path with the barrier's vote.
    int my_ok = 1;
    if (tid >= 32 && tid < 32 + NP) {
      const int m = tid - 32;
      const bool base = L.n_ff == 1 && ffop.kind == FSWEEP_OP_PDELAY && (ffop.flags & FSWEEP_F_ISINT) != 0 &&
                        P.nfft >= 1024 && P.nfft < (1 << 24);
      double dd = 0.0;
      if (base && m < N) dd = rint(__ldg(reinterpret_cast<const double*>(ffop.coef) + m));
      const bool ok = base && dd >= 0.0 && dd * (double)(A.bin_begin + A.n_bins) < 2147483648.0;
      sdi[m] = ok ? (int)dd : 0;
      smag[m] = (ok && m < N) ? exp_t((float)(P.lng * dd)) : 0.f;
      my_ok = ok ? 1 : 0;
    }
    fastd = __syncthreads_and(my_ok) != 0;  // (also publishes the shared tables and the barrier's initialisation)
    if (bulk) mbar_wait(&wbar, 0);
  }

  const int ncols_total = A.batch * A.cols;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  double lacc = 0.0;
  const bool want_ff = BWD && L.n_ff == 1 && ffop.acc_mode == ACC_SMEM;
  const bool ff_delay = ffop.kind == FSWEEP_OP_PDELAY;
  bool first = true;  // this thread's accumulators have not been written yet

  for (long long bl = (long long)blockIdx.x * TPR_BLOCK + tid; bl < A.n_bins; bl += (long long)gridDim.x * TPR_BLOCK) {
    const Ctx<float> ctx = make_ctx<float>(P, A.bin_begin + bl);
    // the first column's input (and target) are requested now: their latency hides behind the chain and the elimination
    const cx<float> xv0 = ld_cx(x + (size_t)bl * A.cols);
    const float tg0 = epi_fused(A.epilogue) ? __ldg(reinterpret_cast<const float*>(A.tgt) + bl) : 0.f;
    // ---- diagonal chain D
    if (fastd) {
      const unsigned kk = (unsigned)(A.bin_begin + bl), nf = (unsigned)P.nfft;
      const float inv_f = (float)P.inv_nfft;
      const double two_inv = 2.0 * P.inv_nfft;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const unsigned t = kk * (unsigned)sdi[m];
        const unsigned q = (unsigned)((float)t * inv_f);  // floor(t / nfft) +- 1
        int r = (int)(t - q * nf);
        if (r < 0) r += (int)nf;
        if (r >= (int)nf) r -= (int)nf;
        float sn, cs;
        sincospif((float)(two_inv * (double)r), &sn, &cs);  // (the same float as delay_eval's: 2 r / nfft rounded once)
        const float mag = smag[m];
        dslot[m * TPR_BLOCK] = f2(mag * cs, -mag * sn);
      }
    } else
      // general chain: runtime loop over the channels, one copy of the response code
#pragma unroll 1
    for (int m = 0; m < NP; ++m) {
      cx<float> d = mk<float>(m < N ? 1.f : 0.f, 0.f);
      for (int i = 0; i < L.n_ff; ++i) {
        bool gd;
        d = cmul(d, op_diag<float>(P.ops[L.ff_begin + i], ctx, m, gd));
      }
      dslot[m * TPR_BLOCK] = f2(d.x, d.y);
    }
    // ---- A = I - D W in registers (rows >= N: identity)
    float2 a[NP][NP];
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      const float2 d = dslot[m * TPR_BLOCK];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float w = wfb[m * NP + j];
        a[m][j] = f2((m == j ? 1.f : 0.f) - d.x * w, -d.y * w);
      }
    }
    // ---- P A = L U, threshold pivoting: straight-line elimination until a step wants an interchange
    unsigned pvec = IDENT;
    int kslow = NP;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (kslow == NP) {
        const float2 d = a[k][k];
        const float pk = d.x * d.x + d.y * d.y;
        float mx = 0.f;
#pragma unroll
        for (int r = k + 1; r < NP; ++r) mx = fmaxf(mx, a[r][k].x * a[r][k].x + a[r][k].y * a[r][k].y);
        if (pk < TPR_TAU2 * mx) {
          kslow = k;
        } else {
          const float id = rcp_t(pk);
          const float2 inv = f2(d.x * id, -d.y * id);
          a[k][k] = inv;
#pragma unroll
          for (int r = k + 1; r < NP; ++r) {
            const float2 l = cmul2(a[r][k], inv);
            a[r][k] = l;
#pragma unroll
            for (int j = k + 1; j < NP; ++j) a[r][j] = cnma2(a[r][j], l, a[k][j]);
          }
        }
      }
    }
    if (kslow != NP) {  // rare: finish with partial pivoting on a local-memory copy
      float2 t[NP * NP];
#pragma unroll
      for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) t[i * NP + j] = a[i][j];
      tpr_factor_slow<NP>(t, kslow, &pvec);
#pragma unroll
      for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) a[i][j] = t[i * NP + j];
    }

    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      // ---- y = A^-1 D (w_pre x):  L U y = P b
      const cx<float> xv = q == 0 ? xv0 : ld_cx(x + (size_t)b * A.xbs + (size_t)bl * A.cols + cc);
      float2 y[NP];
#pragma unroll
      for (int m = 0; m < NP; ++m) y[m] = cmul2(dslot[m * TPR_BLOCK], f2(wpre[m] * xv.x, wpre[m] * xv.y));
      if (pvec != IDENT) {
#pragma unroll
        for (int m = 0; m < NP; ++m) vslot[m * TPR_BLOCK] = y[m];
#pragma unroll
        for (int i = 0; i < NP; ++i) y[i] = vslot[(int)((pvec >> (3 * i)) & 7u) * TPR_BLOCK];
      }
#pragma unroll
      for (int i = 1; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] = cnma2(y[i], a[i][j], y[j]);
#pragma unroll
      for (int i = NP - 1; i >= 0; --i) {
#pragma unroll
        for (int j = i + 1; j < NP; ++j) y[i] = cnma2(y[i], a[i][j], y[j]);
        y[i] = cmul2(y[i], a[i][i]);
      }
      float ox = 0.f, oy = 0.f;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        ox = fmaf(wpost[m], y[m].x, ox);
        oy = fmaf(wpost[m], y[m].y, oy);
      }
      const size_t ooff = (size_t)bl * A.cols + cc;  // one output channel
      if constexpr (!BWD) {
        if (epi_fused(A.epilogue)) {
          const float e = abs_t(ox, oy) - (q == 0 ? tg0 : __ldg(reinterpret_cast<const float*>(A.tgt) + (size_t)b * A.tbs + bl));
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
            const float e = mag - (q == 0 ? tg0 : __ldg(reinterpret_cast<const float*>(A.tgt) + (size_t)b * A.tbs + bl));
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
        // ---- through the output gain: dw_post = Re(go conj(y)); park y; lam = A^-H (w_post go)
        float2 g[NP];
        auto post_acc = [&](auto first_c) {  // (two copies: a thread's FIRST bin writes its accumulators, later ones add)
          constexpr bool F = decltype(first_c)::value;
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            const float v = gox * y[m].x + goy * y[m].y;
            float* p = stage + (S_POST + m) * TPR_SS;
            *p = F ? v : *p + v;
          }
        };
        if (first) post_acc(std::true_type()); else post_acc(std::false_type());
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          yslot[m * TPR_BLOCK] = y[m];
          g[m] = f2(wpost[m] * gox, wpost[m] * goy);
        }
        // A^H lam = g with A = P^T L U:  U^H w = g,  L^H v = w,  lam[perm_i] = v_i
#pragma unroll
        for (int i = 0; i < NP; ++i) {
#pragma unroll
          for (int j = 0; j < i; ++j) g[i] = cnmaj2(g[i], a[j][i], g[j]);
          g[i] = cmul2(g[i], f2(a[i][i].x, -a[i][i].y));
        }
#pragma unroll
        for (int i = NP - 2; i >= 0; --i)
#pragma unroll
          for (int j = i + 1; j < NP; ++j) g[i] = cnmaj2(g[i], a[j][i], g[j]);
        if (pvec != IDENT) {
#pragma unroll
          for (int i = 0; i < NP; ++i) vslot[(int)((pvec >> (3 * i)) & 7u) * TPR_BLOCK] = g[i];
#pragma unroll
          for (int m = 0; m < NP; ++m) g[m] = vslot[m * TPR_BLOCK];
        }
        // ---- lam = g; g_u = conj(D) lam; dW_fb = Re(g_u y^H); diagonal op (u = s + W y); input gain
#pragma unroll
        for (int m = 0; m < NP; ++m) y[m] = yslot[m * TPR_BLOCK];
        float gxr = 0.f, gxi = 0.f;
        auto loop_acc = [&](auto first_c) {
          constexpr bool F = decltype(first_c)::value;
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            const float2 lam = g[m];
            const float2 D = dslot[m * TPR_BLOCK];
            const float2 gu = f2(D.x * lam.x + D.y * lam.y, D.x * lam.y - D.y * lam.x);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              const float v = gu.x * y[j].x + gu.y * y[j].y;
              float* p = stage + (m * NP + j) * TPR_SS;
              *p = F ? v : *p + v;
            }
            if (want_ff) {
              float ux = wpre[m] * xv.x, uy = wpre[m] * xv.y;
#pragma unroll
              for (int j = 0; j < NP; ++j) {
                const float w = wfb[m * NP + j];
                ux = fmaf(w, y[j].x, ux);
                uy = fmaf(w, y[j].y, uy);
              }
              // gh = lam conj(u);  PGAIN: Re gh;  PDELAY (fractional): Re(gh conj((ln g - j w) D))
              const float ghx = lam.x * ux + lam.y * uy, ghy = lam.y * ux - lam.x * uy;
              float v = ghx;
              if (ff_delay) {
                const float tx = (float)ctx.lng * D.x + ctx.omega * D.y;
                const float ty = (float)ctx.lng * D.y - ctx.omega * D.x;
                v = ghx * tx + ghy * ty;
              }
              float* p = stage + (S_DIAG + m) * TPR_SS;
              *p = F ? v : *p + v;
            }
            {
              const float v = gu.x * xv.x + gu.y * xv.y;
              float* p = stage + (S_PRE + m) * TPR_SS;
              *p = F ? v : *p + v;
            }
            gxr = fmaf(wpre[m], gu.x, gxr);
            gxi = fmaf(wpre[m], gu.y, gxi);
          }
        };
        if (first) loop_acc(std::true_type()); else loop_acc(std::false_type());
        first = false;
        if (A.gx != nullptr)
          st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)b * A.gxbs + (size_t)bl * A.cols + cc, mk<float>(gxr, gxi));
      }
    }
  }

  if constexpr (BWD) {
    // ---- block reduction of the per-thread accumulator rows stage[s][0 .. BLOCK): thread (s, quarter) sums a quarter of
    //      row s (the odd row stride keeps the 32 rows of a warp on 32 banks), the four quarters meet in `red`.
    if (first) {  // a thread without a bin (the grid's tail) contributes zeros
#pragma unroll 4
      for (int s = 0; s < S_TOTAL; ++s) stage[s * TPR_SS] = 0.f;
    }
    __syncthreads();
    float* red = reinterpret_cast<float*>(tpr_raw + SM::stage_bytes + (size_t)3 * NP * TPR_BLOCK * 8);
    static_assert(TPR_BLOCK % 4 == 0, "quarter layout");
    constexpr int QN = TPR_BLOCK / 4;  // values per quarter
    static_assert(QN % 4 == 0, "four accumulation chains");
    for (int t = tid; t < 4 * S_TOTAL; t += TPR_BLOCK) {
      const int s = t % S_TOTAL, qd = t / S_TOTAL;
      const float* row = reinterpret_cast<const float*>(tpr_raw) + (size_t)s * TPR_SS + qd * QN;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
      for (int i = 0; i < QN; i += 4) {
        s0 += row[i];
        s1 += row[i + 1];
        s2 += row[i + 2];
        s3 += row[i + 3];
      }
      red[qd * S_TOTAL + s] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    const OpK& fbop = P.ops[L.fb];
    const OpK& preop = P.ops[L.pre];
    const OpK& postop = P.ops[L.post];
    float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    for (int sidx = tid; sidx < S_TOTAL; sidx += TPR_BLOCK) {
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
      if (dst >= 0) partial[dst] = (red[sidx] + red[S_TOTAL + sidx]) + (red[2 * S_TOTAL + sidx] + red[3 * S_TOTAL + sidx]);
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<float>(lacc, A.loss_partial);
}

cudaError_t launch_tpr(int np, bool bwd, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A,
                       int G);

}  // namespace fsweep
