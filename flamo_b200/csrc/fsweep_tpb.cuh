// fsweep_tpb.cuh — thread-per-bin sweep for SMALL FDN loops (width <= 8, float32).
//
// Same pattern as fsweep_loop.cuh ([GAIN] RECURSION(diagonal chain ; one real matrix) [GAIN]) and the same math,
// but ONE THREAD owns one frequency bin: the whole NP x NP complex matrix A = I - D(w) W lives in that thread's
// registers, LU-factored with partial pivoting (row swaps are compile-time-indexed conditional selects), and the
// forward / adjoint solves need no exchange at all.  At N = 8 the row-distributed mapping spends ~3/4 of its
// instructions on exchanges, role selects and pivot bookkeeping that this mapping simply does not have: ~85
// warp-instructions per bin instead of ~360 (profiles/r01_notes.md).
//
// The constant matrices (W_fb, W_pre, W_post) are block-shared in shared memory (every thread reads the same address:
// broadcast); coefficient gradients are accumulated in thread-private shared-memory columns laid out exactly like
// the flat accumulator of the plan ([op.acc_off + row*row_len + e]) and reduced per block into the same `partial`
// layout the row-distributed kernels produce, so fsweep_finalize_kernel is shared.
#pragma once
#include "fsweep_loop.cuh"

namespace fsweep {

constexpr int TPB_BLOCK = 64;

__device__ __forceinline__ void swap_if(bool p, cx<float>& a, cx<float>& b) {
  const cx<float> ta = a, tb = b;
  a.x = p ? tb.x : ta.x;
  a.y = p ? tb.y : ta.y;
  b.x = p ? ta.x : tb.x;
  b.y = p ? ta.y : tb.y;
}

// Thread-private NP x NP complex LU, partial pivoting, LINPACK convention: at step k rows k and p_k are swapped in
// columns >= k only, multipliers stay where they were produced, and the interchanges are replayed progressively
// by the solves:  M_{n-1} P_{n-1} ... M_0 P_0 A = U.
template <int NP>
struct TLU {
  cx<float> a[NP][NP];
  cx<float> dinv[NP];
  unsigned perm;  // 3 bits per step

  __device__ __forceinline__ void factor() {
    perm = 0u;
    static_for<0, NP>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      float best = a[k][k].x * a[k][k].x + a[k][k].y * a[k][k].y;
      int pr = k;
      static_for<k + 1, NP>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const float m = a[r][k].x * a[r][k].x + a[r][k].y * a[r][k].y;
        const bool gt = m > best;
        best = gt ? m : best;
        pr = gt ? r : pr;
      });
      perm |= (unsigned)pr << (3 * k);
      static_for<k + 1, NP>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const bool sw = pr == r;
        static_for<k, NP>([&](auto jc) { swap_if(sw, a[k][decltype(jc)::value], a[r][decltype(jc)::value]); });
      });
      const cx<float> inv = crcp(a[k][k]);
      dinv[k] = inv;
      static_for<k + 1, NP>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const cx<float> l = cmul(a[r][k], inv);
        a[r][k] = l;
        static_for<k + 1, NP>([&](auto jc) { cfnma(a[r][decltype(jc)::value], l, a[k][decltype(jc)::value]); });
      });
    });
  }

  __device__ __forceinline__ void solve(cx<float> (&b)[NP]) const {
    static_for<0, NP>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const int pr = (int)((perm >> (3 * k)) & 7u);
      static_for<k + 1, NP>([&](auto rc) { swap_if(pr == decltype(rc)::value, b[k], b[decltype(rc)::value]); });
      static_for<k + 1, NP>([&](auto rc) { cfnma(b[decltype(rc)::value], a[decltype(rc)::value][k], b[k]); });
    });
    static_rfor<0, NP>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      static_for<k + 1, NP>([&](auto jc) { cfnma(b[k], a[k][decltype(jc)::value], b[decltype(jc)::value]); });
      b[k] = cmul(b[k], dinv[k]);
    });
  }

  // A^H lam = g :  w = U^-H g, then for k = n-1 .. 0 :  w <- P_k (M_k^H w)
  __device__ __forceinline__ void solve_adj(cx<float> (&g)[NP]) const {
    static_for<0, NP>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      static_for<0, i>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const cx<float> t = cmul(mk<float>(a[j][i].x, -a[j][i].y), g[j]);
        g[i].x -= t.x;
        g[i].y -= t.y;
      });
      g[i] = cmul(g[i], mk<float>(dinv[i].x, -dinv[i].y));
    });
    static_rfor<0, NP>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      static_for<k + 1, NP>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const cx<float> t = cmul(mk<float>(a[r][k].x, -a[r][k].y), g[r]);
        g[k].x -= t.x;
        g[k].y -= t.y;
      });
      const int pr = (int)((perm >> (3 * k)) & 7u);
      static_for<k + 1, NP>([&](auto rc) { swap_if(pr == decltype(rc)::value, g[k], g[decltype(rc)::value]); });
    });
  }
};

// block-shared zero-padded copies of the three constant matrices (row-major NP x NP)
template <int NP>
__device__ __forceinline__ void tpb_load_weights(const ProgK& P, const LoopInfo& L, float* wfb, float* wpre,
                                                 float* wpost) {
  for (int e = threadIdx.x; e < NP * NP; e += blockDim.x) {
    const int r = e / NP, c = e - r * NP;
    const OpK& fb = P.ops[L.fb];
    wfb[e] = (r < fb.n_out && c < fb.n_in) ? __ldg(reinterpret_cast<const float*>(fb.coef) + r * fb.n_in + c) : 0.f;
    float vp = 0.f, vq = 0.f;
    if (L.pre >= 0) {
      const OpK& o = P.ops[L.pre];
      if (r < o.n_out && c < o.n_in) vp = __ldg(reinterpret_cast<const float*>(o.coef) + r * o.n_in + c);
    }
    if (L.post >= 0) {
      const OpK& o = P.ops[L.post];
      if (r < o.n_out && c < o.n_in) vq = __ldg(reinterpret_cast<const float*>(o.coef) + r * o.n_in + c);
    }
    wpre[e] = vp;
    wpost[e] = vq;
  }
  __syncthreads();
}

template <int NP>
__device__ __forceinline__ void tpb_chain(const ProgK& P, const LoopInfo& L, const Ctx<float>& ctx, cx<float> (&D)[NP]) {
  static_for<0, NP>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    cx<float> d = mk<float>(1.f, 0.f);
    for (int i = 0; i < L.n_ff; ++i) {
      bool gd;
      d = cmul(d, op_diag<float>(P.ops[L.ff_begin + i], ctx, m, gd));
    }
    D[m] = d;  // 0 for m >= loop width: the padded rows of A are identity
  });
}

template <int NP>
__device__ __forceinline__ void tpb_build(const float* wfb, const cx<float> (&D)[NP], TLU<NP>& lu) {
  static_for<0, NP>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    static_for<0, NP>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      const float w = wfb[m * NP + j];
      lu.a[m][j] = mk<float>((m == j ? 1.f : 0.f) - D[m].x * w, -D[m].y * w);
    });
  });
  lu.factor();
}

// recursion input s = W_pre x (or x), xin = the raw input channels (zero padded)
template <int NP>
__device__ __forceinline__ void tpb_input(const ProgK& P, const LoopInfo& L, const float* wpre, const cx<float>* x,
                                          long long xbs, long long bl, int b, int cc, int cols, cx<float> (&xin)[NP],
                                          cx<float> (&s)[NP]) {
  const int n_in = P.in_ch;
  static_for<0, NP>([&](auto nc) {
    constexpr int n = decltype(nc)::value;
    xin[n] = mk<float>(0.f, 0.f);
    if (n < n_in) xin[n] = ld_cx(x + (size_t)b * xbs + ((size_t)bl * n_in + n) * cols + cc);
  });
  if (L.pre < 0) {
    static_for<0, NP>([&](auto mc) { s[decltype(mc)::value] = xin[decltype(mc)::value]; });
    return;
  }
  static_for<0, NP>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    cx<float> acc = mk<float>(0.f, 0.f);
    static_for<0, NP>([&](auto nc) {
      constexpr int n = decltype(nc)::value;
      if (n < n_in) {
        const float w = wpre[m * NP + n];
        acc.x = fmaf(w, xin[n].x, acc.x);
        acc.y = fmaf(w, xin[n].y, acc.y);
      }
    });
    s[m] = acc;
  });
}

template <int NP>
__device__ __forceinline__ void tpb_output(const ProgK& P, const LoopInfo& L, const float* wpost,
                                           const cx<float> (&y)[NP], cx<float> (&o)[NP]) {
  if (L.post < 0) {
    static_for<0, NP>([&](auto mc) { o[decltype(mc)::value] = y[decltype(mc)::value]; });
    return;
  }
  const int n_out = P.out_ch;
  static_for<0, NP>([&](auto rc) {
    constexpr int r = decltype(rc)::value;
    cx<float> acc = mk<float>(0.f, 0.f);
    if (r < n_out) {
      static_for<0, NP>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        const float w = wpost[r * NP + m];
        acc.x = fmaf(w, y[m].x, acc.x);
        acc.y = fmaf(w, y[m].y, acc.y);
      });
    }
    o[r] = acc;
  });
}

// fused criterion, one thread per bin: e over this bin's output rows; returns dL/d|Y_r| per row in ga[]
template <int NP>
__device__ __forceinline__ void tpb_crit(const SweepArgs& A, const cx<float> (&o)[NP], int n_out, long long bl, int b,
                                         float (&mag)[NP], float (&ga)[NP], double& lacc) {
  const float* tg = reinterpret_cast<const float*>(A.tgt) + (size_t)b * A.tbs;
  const float sc = (float)(2.0 * A.crit_scale);
  float tot = 0.f;
  static_for<0, NP>([&](auto rc) {
    constexpr int r = decltype(rc)::value;
    mag[r] = (r < n_out) ? abs_t(o[r].x, o[r].y) : 0.f;
    tot += mag[r];
  });
  if (A.epilogue == EPI_ABSSUM_MSE) {
    const float e = tot - __ldg(tg + bl);
    lacc += (double)e * (double)e;
    static_for<0, NP>([&](auto rc) { ga[decltype(rc)::value] = sc * e; });
  } else {
    static_for<0, NP>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      ga[r] = 0.f;
      if (r < n_out) {
        const float e = mag[r] - __ldg(tg + (size_t)bl * n_out + r);
        lacc += (double)e * (double)e;
        ga[r] = sc * e;
      }
    });
  }
}

// ------------------------------------------------------------------------------------ forward
template <int NP>
__global__ void __launch_bounds__(TPB_BLOCK, 5) fsweep_tpb_fwd_kernel(const __grid_constant__ ProgK P,
                                                                    const __grid_constant__ LoopInfo L,
                                                                    const SweepArgs A) {
  __shared__ float wfb[NP * NP], wpre[NP * NP], wpost[NP * NP];
  tpb_load_weights<NP>(P, L, wfb, wpre, wpost);
  const int ncols_total = A.batch * A.cols;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  const int n_out = P.out_ch;
  double lacc = 0.0;

  for (long long bl = (long long)blockIdx.x * TPB_BLOCK + threadIdx.x; bl < A.n_bins;
       bl += (long long)gridDim.x * TPB_BLOCK) {
    const Ctx<float> ctx = make_ctx<float>(P, A.bin_begin + bl);
    cx<float> D[NP];
    tpb_chain<NP>(P, L, ctx, D);
    TLU<NP> lu;
    tpb_build<NP>(wfb, D, lu);
    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      cx<float> xin[NP], y[NP], o[NP];
      tpb_input<NP>(P, L, wpre, x, A.xbs, bl, b, cc, A.cols, xin, y);
      static_for<0, NP>([&](auto mc) { y[decltype(mc)::value] = cmul(D[decltype(mc)::value], y[decltype(mc)::value]); });
      lu.solve(y);
      tpb_output<NP>(P, L, wpost, y, o);
      if (epi_fused(A.epilogue)) {
        float mag[NP], ga[NP];
        tpb_crit<NP>(A, o, n_out, bl, b, mag, ga, lacc);
        continue;
      }
      static_for<0, NP>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        if (r < n_out) {
          const size_t off = (size_t)b * A.ybs + ((size_t)bl * n_out + r) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS)
            reinterpret_cast<float*>(A.y)[off] = abs_t(o[r].x, o[r].y);
          else
            st_cx(reinterpret_cast<cx<float>*>(A.y) + off, o[r]);
        }
      });
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<float>(lacc, A.loss_partial);
}

// ------------------------------------------------------------------------------------ backward
template <int NP>
__global__ void __launch_bounds__(TPB_BLOCK, 4) fsweep_tpb_bwd_kernel(const __grid_constant__ ProgK P,
                                                                    const __grid_constant__ LoopInfo L,
                                                                    const SweepArgs A, int G) {
  __shared__ float wfb[NP * NP], wpre[NP * NP], wpost[NP * NP];
  extern __shared__ __align__(16) float sacc[];  // [acc_total][TPB_BLOCK], thread-private columns
  tpb_load_weights<NP>(P, L, wfb, wpre, wpost);
  const int tid = threadIdx.x;
  for (int i = 0; i < P.acc_total; ++i) sacc[i * TPB_BLOCK + tid] = 0.f;

  const int ncols_total = A.batch * A.cols;
  const cx<float>* x = reinterpret_cast<const cx<float>*>(A.x);
  const int n_out = P.out_ch, n_in = P.in_ch, N = P.rec_n;
  const OpK& fbop = P.ops[L.fb];
  const bool want_fb = fbop.acc_mode == ACC_SMEM;
  const bool want_pre = L.pre >= 0 && P.ops[L.pre].acc_mode == ACC_SMEM;
  const bool want_post = L.post >= 0 && P.ops[L.post].acc_mode == ACC_SMEM;
  const OpK& ffop = P.ops[L.ff_begin];  // gradient supported for a single-op chain only (host-checked)
  const bool want_ff = L.n_ff == 1 && ffop.acc_mode == ACC_SMEM;
  auto accp = [&](const OpK& op, int row, int e, float v) { sacc[(op.acc_off + row * op.row_len + e) * TPB_BLOCK + tid] += v; };
  double lacc = 0.0;

  for (long long bl = (long long)blockIdx.x * TPB_BLOCK + tid; bl < A.n_bins; bl += (long long)gridDim.x * TPB_BLOCK) {
    const Ctx<float> ctx = make_ctx<float>(P, A.bin_begin + bl);
    cx<float> D[NP];
    tpb_chain<NP>(P, L, ctx, D);
    TLU<NP> lu;
    tpb_build<NP>(wfb, D, lu);
    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      cx<float> xin[NP], s[NP], y[NP], o[NP];
      tpb_input<NP>(P, L, wpre, x, A.xbs, bl, b, cc, A.cols, xin, s);
      static_for<0, NP>([&](auto mc) { y[decltype(mc)::value] = cmul(D[decltype(mc)::value], s[decltype(mc)::value]); });
      lu.solve(y);
      tpb_output<NP>(P, L, wpost, y, o);
      // ---- output gradient
      cx<float> go[NP];
      if (epi_fused(A.epilogue)) {
        float mag[NP], ga[NP];
        tpb_crit<NP>(A, o, n_out, bl, b, mag, ga, lacc);
        static_for<0, NP>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          go[r] = mk<float>(0.f, 0.f);
          if (r < n_out && mag[r] > 0.f) {
            const float t = ga[r] * rcp_t(mag[r]);
            go[r] = mk<float>(t * o[r].x, t * o[r].y);
          }
        });
      } else
      static_for<0, NP>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        go[r] = mk<float>(0.f, 0.f);
        if (r < n_out) {
          const size_t off = (size_t)b * A.gybs + ((size_t)bl * n_out + r) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS) {
            const float ga = __ldg(reinterpret_cast<const float*>(A.gy) + off);
            const float mag = abs_t(o[r].x, o[r].y);
            if (mag > 0.f) {
              const float t = ga * rcp_t(mag);
              go[r] = mk<float>(t * o[r].x, t * o[r].y);
            }
          } else {
            go[r] = ld_cx(reinterpret_cast<const cx<float>*>(A.gy) + off);
          }
        }
      });
      // ---- through the output gain
      cx<float> lam[NP];
      if (L.post < 0) {
        static_for<0, NP>([&](auto mc) { lam[decltype(mc)::value] = go[decltype(mc)::value]; });
      } else {
        static_for<0, NP>([&](auto mc) { lam[decltype(mc)::value] = mk<float>(0.f, 0.f); });
        static_for<0, NP>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          if (r < n_out) {
            static_for<0, NP>([&](auto mc) {
              constexpr int m = decltype(mc)::value;
              const float w = wpost[r * NP + m];
              lam[m].x = fmaf(w, go[r].x, lam[m].x);
              lam[m].y = fmaf(w, go[r].y, lam[m].y);
              if (want_post && m < N) accp(P.ops[L.post], r, m, go[r].x * y[m].x + go[r].y * y[m].y);
            });
          }
        });
      }
      // ---- adjoint solve, loop input, feedback-matrix gradient
      lu.solve_adj(lam);
      cx<float> gu[NP], u[NP];
      static_for<0, NP>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        gu[m] = mk<float>(D[m].x * lam[m].x + D[m].y * lam[m].y, D[m].x * lam[m].y - D[m].y * lam[m].x);
        cx<float> acc = s[m];
        static_for<0, NP>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          const float w = wfb[m * NP + j];
          acc.x = fmaf(w, y[j].x, acc.x);
          acc.y = fmaf(w, y[j].y, acc.y);
        });
        u[m] = acc;
      });
      if (want_fb) {
        static_for<0, NP>([&](auto mc) {
          constexpr int m = decltype(mc)::value;
          if (m < N) {
            static_for<0, NP>([&](auto jc) {
              constexpr int j = decltype(jc)::value;
              if (j < N) accp(fbop, m, j, gu[m].x * y[j].x + gu[m].y * y[j].y);
            });
          }
        });
      }
      // ---- diagonal chain (single op): gh = lam conj(u)
      if (want_ff) {
        static_for<0, NP>([&](auto mc) {
          constexpr int m = decltype(mc)::value;
          if (m < N) {
            const cx<float> gh = cmulc(lam[m], u[m]);
            if (ffop.kind == FSWEEP_OP_PGAIN) {
              accp(ffop, m, 0, gh.x);
            } else if (ffop.kind == FSWEEP_OP_PDELAY && !(ffop.flags & FSWEEP_F_ISINT)) {
              const cx<float> t = cmul(mk<float>((float)ctx.lng, -ctx.omega), D[m]);
              accp(ffop, m, 0, gh.x * t.x + gh.y * t.y);
            }
          }
        });
      }
      // ---- through the input gain
      if (L.pre < 0) {
        if (A.gx != nullptr) {
          static_for<0, NP>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            if (m < n_in)
              st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)b * A.gxbs + ((size_t)bl * n_in + m) * A.cols + cc, gu[m]);
          });
        }
      } else {
        static_for<0, NP>([&](auto nc) {
          constexpr int n = decltype(nc)::value;
          if (n < n_in) {
            cx<float> gxn = mk<float>(0.f, 0.f);
            static_for<0, NP>([&](auto mc) {
              constexpr int m = decltype(mc)::value;
              const float w = wpre[m * NP + n];
              gxn.x = fmaf(w, gu[m].x, gxn.x);
              gxn.y = fmaf(w, gu[m].y, gxn.y);
              if (want_pre && m < N) accp(P.ops[L.pre], m, n, gu[m].x * xin[n].x + gu[m].y * xin[n].y);
            });
            if (A.gx != nullptr)
              st_cx(reinterpret_cast<cx<float>*>(A.gx) + (size_t)b * A.gxbs + ((size_t)bl * n_in + n) * A.cols + cc, gxn);
          }
        });
      }
    }
  }

  // ---- block reduction into the row-distributed `partial` layout: partial[block][(row_off + i) * G + row]
  __syncthreads();
  float* partial = reinterpret_cast<float*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
  for (int e = tid; e < P.acc_per_lane * G; e += TPB_BLOCK) partial[e] = 0.f;
  __syncthreads();
  for (int opi = 0; opi < P.n_ops; ++opi) {
    const OpK& op = P.ops[opi];
    if (op.acc_mode != ACC_SMEM) continue;
    const int total = op.n_out * op.row_len;
    for (int e = tid; e < total; e += TPB_BLOCK) {
      const int row = e / op.row_len, i = e - row * op.row_len;
      const float* col = sacc + (size_t)(op.acc_off + e) * TPB_BLOCK;
      float sum = 0.f;
      for (int j = 0; j < TPB_BLOCK; ++j) sum += col[(j + tid) & (TPB_BLOCK - 1)];  // rotated: conflict-free
      partial[(op.row_off + i) * G + row] = sum;
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<float>(lacc, A.loss_partial);
}

cudaError_t launch_tpb_fwd(int np, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A);
cudaError_t launch_tpb_bwd(int np, int grid, size_t smem, cudaStream_t st, const ProgK& P, const LoopInfo& L,
                           const SweepArgs& A, int G);

}  // namespace fsweep
