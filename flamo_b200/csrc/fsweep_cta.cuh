// fsweep_cta.cuh — CTA-per-bin sweep for WIDE FDN loops (32 < N <= 64, float32; BASELINE config 5: 64 x 64, batch 32).
//
// Same pattern and math as fsweep_loop.cuh / fsweep_tpc.cuh
//     [GAIN N x 1]  ->  RECURSION( diagonal chain ; one real N x N matrix )  ->  [GAIN 1 x N]
// but one thread BLOCK (256 threads) owns one frequency bin, with the per-bin matrix A = I - D(w) W and up to 32
// right-hand sides resident in shared memory (33 KB + 17 KB), the way a batched small-matrix LAPACK would do it:
//   * P A = L U: right-looking, the (63-k)^2 trailing update of step k spread over all 256 threads;
//   * the loop has ONE input channel, so every right-hand side of a bin is a multiple of one vector: y_b = x_b z with
//     z = A^-1 (D w_pre) solved once per bin whatever the batch (the reference, system.py:417-425, likewise solves for
//     the closed-loop response once and applies it to the batch with an einsum); o_b = (w_post . z) x_b;
//   * the output has ONE channel, so the adjoint solve is shared by the whole batch too: lambda_b = g_b v with
//     v = A^-H w_post, and the feedback-matrix gradient collapses to ONE outer product per bin,
//         dW = Re( (conj(D) v) (xbar conj(z))^T ),   xbar = sum_b g_b conj(x_b),
//     accumulated in 16 registers per thread across all bins of the block: no atomics at all (the row-distributed
//     path issued 4096 global atomics per bin and batch chunk — 2.5e10 per step of config 5);
//   * gradients leave the block once, in the `partial` layout fsweep_finalize_kernel sums in float64.
// Gradients of the diagonal chain (learnable delays / gains inside the loop) are not formed here: such plans stay on
// the row-distributed kernels (host-checked, plan->cta).
#pragma once
#include <type_traits>

#include "fsweep_tpc.cuh"
#include "fsweep_tc.cuh"

namespace fsweep {

constexpr int CTA_T = 256;    // threads per block
constexpr int CN = 64;        // padded loop width
constexpr int CLD = CN + 1;   // row stride of A in float2 (odd: column walks are conflict-free)

__device__ __forceinline__ float2 cmulj2(float2 a, float2 b) {  // conj(a) * b
  return f2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
// the same helpers on double2: the SIMT elimination is also instantiated in float64 (float64 models with a 33..64-wide
// FDN loop: the reference's examples default to float64, examples/e8_colorless_fdn.py:197)
__device__ __forceinline__ double2 f2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ double2 cmul2(double2 a, double2 b) { return f2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulj2(double2 a, double2 b) { return f2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double2 cnma2(double2 acc, double2 a, double2 b) {
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}
template <typename T>
struct V2of;
template <>
struct V2of<float> {
  using type = float2;
};
template <>
struct V2of<double> {
  using type = double2;
};
// |c|^2 as an unsigned that orders like the value (pivot search key); float64: the high word of the bit pattern
__device__ __forceinline__ unsigned mag_bits(float m) { return __float_as_uint(m); }
__device__ __forceinline__ unsigned mag_bits(double m) { return (unsigned)((unsigned long long)__double_as_longlong(m) >> 32); }

constexpr size_t cta_smem_bytes(bool bwd, bool tc, size_t real_size = sizeof(float)) {
  return (size_t)(CN * CLD + (tc ? 4 : 6) * CN + 16) * 2 * real_size + (size_t)(2 * CN) * real_size +
         (size_t)(2 * CN + 8) * sizeof(int) +
         (tc ? (size_t)CN * sizeof(int) + 16 : (bwd ? (size_t)(CN * CN) * real_size : 0));
}
static_assert(tc::scratch_bytes() <= (size_t)CN * CLD * sizeof(float2), "tensor-core scratch must fit in the L\\U area");

template <typename T, bool BWD, bool TC>
__global__ void __launch_bounds__(TC ? tc::T : CTA_T, TC ? tc::BLOCKS_PER_SM : (sizeof(T) == 8 ? 1 : 4)) fsweep_cta_kernel(const __grid_constant__ ProgK P,
                                                            const __grid_constant__ LoopInfo L, const SweepArgs A, int G) {
  static_assert(!TC || sizeof(T) == 4, "the tensor-core elimination is float32 (3 x TF32)");
  using V2 = typename V2of<T>::type;
  extern __shared__ __align__(16) unsigned char csm_raw[];
  V2* csm = reinterpret_cast<V2*>(csm_raw);
  V2* sA = csm;                 // [CN][CLD]   L\U in place, 1/U_kk on the diagonal
  V2* sD = sA + CN * CLD;       // [CN]  diagonal chain response
  V2* sV = sD + CN;             // [CN]  adjoint vector v, then vd = conj(D) v
  V2* sZ = sV + CN;             // [CN]  z = A^-1 (D w_pre): every y_b = x_b z
  V2* sRed = sZ + CN;           // [16] h, S, per-warp partial sums of xbar
  V2* sL = sRed + 16;           // [2][CN] multipliers of the current elimination step (SIMT elimination only)
  V2* sInv = sL + (TC ? 0 : 2 * CN);  // [CN] reciprocal pivots
  T* sWpre = reinterpret_cast<T*>(sInv + CN);  // [CN]
  T* sWpost = sWpre + CN;                               // [CN]
  int* sPiv = reinterpret_cast<int*>(sWpost + CN);          // [CN] row map of P A; [CN..]: scalars
  int* sPos = sPiv + CN;                                    // [CN] inverse row map
  int* sScalar = sPos + CN;

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int N = P.rec_n;
  const OpK& fbop = P.ops[L.fb];
  const T* Wfb = reinterpret_cast<const T*>(fbop.coef);
  if (t < CN) {
    sWpre[t] = t < N ? __ldg(reinterpret_cast<const T*>(P.ops[L.pre].coef) + t) : T(0);
    sWpost[t] = t < N ? __ldg(reinterpret_cast<const T*>(P.ops[L.post].coef) + t) : T(0);
  }
  const int Q = A.batch * A.cols;
  const cx<T>* x = reinterpret_cast<const cx<T>*>(A.x);
  double lacc = 0.0;
  constexpr int NT = TC ? tc::T : CTA_T;  // threads per block
  constexpr int NGW = CN * CN / NT;       // dW_fb entries per thread: e = t + NT i (m = e >> 6, j = e & 63)
  // SIMT elimination: the dW_fb accumulators live in shared memory across the bins of this block (the elimination
  // wants every register for the matrix tile); tensor-core elimination: the matrix is in TMEM, so they are registers
  T* sGw = reinterpret_cast<T*>(sScalar + 8);  // [CN * CN], BWD && !TC only
  int* sFin = sScalar + 8;                              // [CN], TC only
  uint64_t* sBar = reinterpret_cast<uint64_t*>(sFin + CN);  // TC only (8-byte aligned: every array above is)
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 1);
  T gwr[(BWD && TC) ? NGW : 1];
  tc::State TS;
  if constexpr (TC) tc::setup(TS, sTmem, sBar);
  T gpre = T(0), gpost = T(0);
  if constexpr (BWD) {
#pragma unroll
    for (int i = 0; i < NGW; ++i) {
      if constexpr (TC)
        gwr[i] = T(0);
      else
        sGw[t + NT * i] = T(0);
    }
  }
  __syncthreads();

  for (long long bl = blockIdx.x; bl < A.n_bins; bl += gridDim.x) {
    // ---- 1. diagonal chain, row map
    if (t < CN) {
      const Ctx<T> ctx = make_ctx<T>(P, A.bin_begin + bl);
      cx<T> d = mk<T>(t < N ? T(1) : T(0), T(0));
      for (int i = 0; i < L.n_ff; ++i) {
        bool gd;
        d = cmul(d, op_diag<T>(P.ops[L.ff_begin + i], ctx, t, gd));
      }
      sD[t] = f2(d.x, d.y);
    }
    __syncthreads();
    if constexpr (TC) {
      // ---- 2 + 3 on the tensor cores: the matrix lives in TMEM, rank-8 updates as 3 x TF32 tcgen05.mma (fsweep_tc.cuh)
      if constexpr (sizeof(T) == 4) tc::lu<CLD>(TS, sA, sD, Wfb, N, sPiv, sPos, sFin, sInv);
    } else {
    // ---- 2. A = I - D W in REGISTERS (rows / columns >= N: identity).  Warp w owns the columns w + 8 j (j < 8), lane l
    //         the rows l and l + 32: a 2 x 8 tile per thread.
    V2 a[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = lane + 32 * i;
      const V2 d = sD[m];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = warp + 8 * j;
        const T w = (m < N && c < N) ? __ldg(Wfb + m * N + c) : T(0);
        a[i][j] = f2((m == c ? T(1) : T(0)) - d.x * w, -d.y * w);
      }
    }
    // ---- 3. P A = L U with IMPLICIT partial pivoting, as a producer / consumer pipeline over the warps.  Column k
    //         belongs to warp k & 7 (slot k >> 3).  Its owner finds the pivot among the rows that have not been pivots
    //         yet (one REDUX over a packed |.|^2 / row key), scales the column and publishes {pivot row, multipliers
    //         l[0..63]} (l = 0 for finished rows) with bar.arrive; the other warps bar.sync on it, fetch the pivot
    //         row's entries of their own columns with a shuffle (that row lives in lane pr & 31 of EVERY warp) and
    //         update their tiles.  Look-ahead: the owner of column k + 1 updates that column first, factors it and
    //         publishes before touching its other columns, so the next step's multipliers are normally waiting when
    //         the consumers arrive.  Nothing but the multipliers crosses shared memory; no CTA-wide barrier.
    unsigned act = 3u;  // bit i: row lane + 32 i has not been a pivot row yet
    int pr = 0;
    V2 l0 = f2(T(0), T(0)), l1 = f2(T(0), T(0));
    auto upd = [&](int pr_, V2 la, V2 lb, V2& r0, V2& r1) {
      V2 u = (pr_ & 32) ? r1 : r0;
      u.x = __shfl_sync(FULL, u.x, pr_ & 31);
      u.y = __shfl_sync(FULL, u.y, pr_ & 31);
      r0 = cnma2(r0, la, u);
      r1 = cnma2(r1, lb, u);
    };
    auto factor = [&](int k, V2& c0, V2& c1, int& pr_, V2& la, V2& lb) {
      const T m0 = c0.x * c0.x + c0.y * c0.y, m1 = c1.x * c1.x + c1.y * c1.y;
      // |c|^2 >= 0: its bit pattern orders like an unsigned; the low 6 bits carry the row (ties: lowest row)
      const unsigned k0 = (act & 1u) ? ((mag_bits(m0) & ~63u) | (unsigned)(63 - lane)) + 64u : 0u;
      const unsigned k1 = (act & 2u) ? ((mag_bits(m1) & ~63u) | (unsigned)(31 - lane)) + 64u : 0u;
      const unsigned key = __reduce_max_sync(FULL, max(k0, k1));
      pr_ = 63 - (int)(key & 63u);
      V2 pv = (pr_ & 32) ? c1 : c0;
      pv.x = __shfl_sync(FULL, pv.x, pr_ & 31);
      pv.y = __shfl_sync(FULL, pv.y, pr_ & 31);
      const T id = rcp_t(pv.x * pv.x + pv.y * pv.y);
      const V2 inv = f2(pv.x * id, -pv.y * id);
      la = f2(T(0), T(0));
      lb = f2(T(0), T(0));
      if ((act & 1u) && lane != pr_) c0 = la = cmul2(c0, inv);  // L is kept in place
      if ((act & 2u) && lane + 32 != pr_) c1 = lb = cmul2(c1, inv);
      V2* bufL = sL + (k & 1) * CN;
      bufL[lane] = la;
      bufL[lane + 32] = lb;
      if (lane == 0) {
        sScalar[k & 1] = pr_;
        sPiv[k] = pr_;  // position k of P A holds original row pr
        sPos[pr_] = k;
        sInv[k] = inv;
      }
      asm volatile("bar.arrive %0, 256;" ::"r"(3 + (k & 1)) : "memory");
    };
    if (warp == 0) factor(0, a[0][0], a[1][0], pr, l0, l1);
    // One elimination step for this warp.  JK = k >> 3 is static (unrolled), kk = k & 7 a runtime loop index; LAST
    // says kk == 7 (the next column then lives in slot JK + 1 of warp 0).  Slots > JK are always right of column k;
    // slot JK only for the warps > kk.
    auto step = [&](auto JKc, auto LASTc, int kk) {
      constexpr int JK = decltype(JKc)::value;
      constexpr bool LAST = decltype(LASTc)::value;
      const int k = 8 * JK + kk;
      if (warp != kk) {  // consumer of l_k (its owner kept them in registers)
        asm volatile("bar.sync %0, 256;" ::"r"(3 + (k & 1)) : "memory");
        const V2* bufL = sL + (k & 1) * CN;
        pr = sScalar[k & 1];
        l0 = bufL[lane];
        l1 = bufL[lane + 32];
      }
      if ((pr & 31) == lane) act &= ~(1u << (pr >> 5));
      const int opr = pr;
      const V2 ol0 = l0, ol1 = l1;
      if constexpr (!LAST) {
        if (warp > kk) {
          upd(opr, ol0, ol1, a[0][JK], a[1][JK]);
          if (warp == kk + 1) factor(k + 1, a[0][JK], a[1][JK], pr, l0, l1);  // look-ahead
        }
#pragma unroll
        for (int j = JK + 1; j < 8; ++j) upd(opr, ol0, ol1, a[0][j], a[1][j]);
      } else if constexpr (JK < 7) {
        upd(opr, ol0, ol1, a[0][JK + 1], a[1][JK + 1]);
        if (warp == 0) factor(k + 1, a[0][JK + 1], a[1][JK + 1], pr, l0, l1);  // look-ahead
#pragma unroll
        for (int j = JK + 2; j < 8; ++j) upd(opr, ol0, ol1, a[0][j], a[1][j]);
      }
    };
    auto phase = [&](auto JKc) {
#pragma unroll 1
      for (int kk = 0; kk < 7; ++kk) step(JKc, std::false_type{}, kk);
      step(JKc, std::true_type{}, 7);
    };
    phase(std::integral_constant<int, 0>{});
    phase(std::integral_constant<int, 1>{});
    phase(std::integral_constant<int, 2>{});
    phase(std::integral_constant<int, 3>{});
    phase(std::integral_constant<int, 4>{});
    phase(std::integral_constant<int, 5>{});
    phase(std::integral_constant<int, 6>{});
    phase(std::integral_constant<int, 7>{});
    // L\U -> shared memory in pivot order (row pos[r] of P A is original row r), reciprocal pivots on the diagonal
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      V2* row = sA + sPos[lane + 32 * i] * CLD + warp;
#pragma unroll
      for (int j = 0; j < 8; ++j) row[8 * j] = a[i][j];
    }
    __syncthreads();
    if (t < CN) sA[t * CLD + t] = sInv[t];
    __syncthreads();
    }

    // ---- 4. The loop has ONE input and ONE output channel, so every right-hand side of the bin is a multiple of the
    //         same vector: y_b = x_b z with z = A^-1 (D w_pre), o_b = h x_b with h = w_post . z — one forward and (BWD)
    //         one adjoint substitution per bin, whatever the batch.  They run side by side, each inside ONE warp (lane l
    //         owns entries l and l + 32; the entry that becomes final in a step is broadcast with a shuffle: no barrier,
    //         no shared-memory round trip on the dependent chain): warp 0 solves for z, warp 1 for v = A^-H w_post.
    if (warp == 0) {
      // L U z = P r,  r = D w_pre:  row i of P r is r[piv[i]]
      V2 g0, g1;
      {
        const int s0 = sPiv[lane], s1 = sPiv[lane + 32];
        const V2 d0 = sD[s0], d1 = sD[s1];
        const T w0 = sWpre[s0], w1 = sWpre[s1];
        g0 = f2(d0.x * w0, d0.y * w0);
        g1 = f2(d1.x * w1, d1.y * w1);
      }
      const V2* row0 = sA + lane * CLD;
      const V2* row1 = sA + (lane + 32) * CLD;
      // L c = P r (unit lower, column oriented)
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const V2 ci = f2(__shfl_sync(FULL, g0.x, i), __shfl_sync(FULL, g0.y, i));
        if (lane > i) g0 = cnma2(g0, row0[i], ci);
        g1 = cnma2(g1, row1[i], ci);
      }
#pragma unroll 4
      for (int i = 32; i < CN - 1; ++i) {
        const V2 ci = f2(__shfl_sync(FULL, g1.x, i - 32), __shfl_sync(FULL, g1.y, i - 32));
        if (lane + 32 > i) g1 = cnma2(g1, row1[i], ci);
      }
      // U z = c (reciprocal diagonal stored)
#pragma unroll 4
      for (int i = CN - 1; i >= 32; --i) {
        if (lane + 32 == i) g1 = cmul2(g1, row1[i]);
        const V2 zi = f2(__shfl_sync(FULL, g1.x, i - 32), __shfl_sync(FULL, g1.y, i - 32));
        if (lane + 32 < i) g1 = cnma2(g1, row1[i], zi);
        g0 = cnma2(g0, row0[i], zi);
      }
#pragma unroll 4
      for (int i = 31; i >= 0; --i) {
        if (lane == i) g0 = cmul2(g0, row0[i]);
        const V2 zi = f2(__shfl_sync(FULL, g0.x, i), __shfl_sync(FULL, g0.y, i));
        if (lane < i) g0 = cnma2(g0, row0[i], zi);
      }
      sZ[lane] = g0;
      sZ[lane + 32] = g1;
    } else if (BWD && warp == 1) {
      // v = A^-H w_post:  A^H = U^H L^H P
      V2 g0 = f2(sWpost[lane], T(0)), g1 = f2(sWpost[lane + 32], T(0));
      const V2* col0 = sA + lane;       // entry [i][lane]
      const V2* col1 = sA + lane + 32;  // entry [i][lane + 32]
      auto cjnma = [](V2 acc, V2 a, V2 b) {  // acc - conj(a) b
        acc.x -= a.x * b.x + a.y * b.y;
        acc.y -= a.x * b.y - a.y * b.x;
        return acc;
      };
      // U^H w = g  (lower triangular, column oriented: once w_i is final, g_j -= conj(U[i][j]) w_i for j > i)
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        if (lane == i) {
          const V2 di = col0[i * CLD];
          g0 = cmul2(g0, f2(di.x, -di.y));
        }
        const V2 wi = f2(__shfl_sync(FULL, g0.x, i), __shfl_sync(FULL, g0.y, i));
        if (lane > i) g0 = cjnma(g0, col0[i * CLD], wi);
        g1 = cjnma(g1, col1[i * CLD], wi);
      }
#pragma unroll 4
      for (int i = 32; i < CN; ++i) {
        if (lane + 32 == i) {
          const V2 di = col1[i * CLD];
          g1 = cmul2(g1, f2(di.x, -di.y));
        }
        const V2 wi = f2(__shfl_sync(FULL, g1.x, i - 32), __shfl_sync(FULL, g1.y, i - 32));
        if (lane + 32 > i) g1 = cjnma(g1, col1[i * CLD], wi);
      }
      // L^H z = w  (unit upper triangular): z_i is final when all j > i are done; z_j -= conj(L[i][j]) z_i for j < i
#pragma unroll 4
      for (int i = CN - 1; i >= 32; --i) {
        const V2 zi = f2(__shfl_sync(FULL, g1.x, i - 32), __shfl_sync(FULL, g1.y, i - 32));
        if (lane + 32 < i) g1 = cjnma(g1, col1[i * CLD], zi);
        g0 = cjnma(g0, col0[i * CLD], zi);
      }
#pragma unroll 4
      for (int i = 31; i > 0; --i) {
        const V2 zi = f2(__shfl_sync(FULL, g0.x, i), __shfl_sync(FULL, g0.y, i));
        if (lane < i) g0 = cjnma(g0, col0[i * CLD], zi);
      }
      // v[piv[i]] = z_i ; vd = conj(D) v
      {
        const int p0 = sPiv[lane], p1 = sPiv[lane + 32];
        const V2 d0 = sD[p0], d1 = sD[p1];
        sV[p0] = f2(d0.x * g0.x + d0.y * g0.y, d0.x * g0.y - d0.y * g0.x);
        sV[p1] = f2(d1.x * g1.x + d1.y * g1.y, d1.x * g1.y - d1.y * g1.x);
      }
    }
    __syncthreads();
    // ---- 5. h = w_post . z (warp 0);  S = sum_m w_pre[m] vd[m] (warp 1, BWD: g_x = g_b S)
    if (warp < (BWD ? 2 : 1)) {
      const T* wv = warp == 0 ? sWpost : sWpre;
      const V2* vec = warp == 0 ? sZ : sV;
      T sx = T(0), sy = T(0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const T w = wv[lane + 32 * h];
        sx = fma(w, vec[lane + 32 * h].x, sx);
        sy = fma(w, vec[lane + 32 * h].y, sy);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(FULL, sx, o);
        sy += __shfl_xor_sync(FULL, sy, o);
      }
      if (lane == 0) sRed[warp] = f2(sx, sy);
    }
    __syncthreads();
    // ---- 6. the batch: o_b = h x_b, criterion / output gradient g_b, xbar = sum_b g_b conj(x_b)
    {
      const V2 hh = sRed[0];
      const V2 S = BWD ? sRed[1] : f2(T(0), T(0));
      T xbx = T(0), xby = T(0);
      for (int q = t; q < Q; q += NT) {
        const int bb = (A.cols == 1) ? q : q / A.cols, cc = q - bb * A.cols;
        const size_t ooff = (size_t)bl * A.cols + cc;
        const cx<T> xv = ld_cx(x + (size_t)bb * A.xbs + ooff);
        const T ox = hh.x * xv.x - hh.y * xv.y, oy = hh.x * xv.y + hh.y * xv.x;
        if constexpr (!BWD) {
          if (epi_fused(A.epilogue)) {
            const T e = abs_t(ox, oy) - __ldg(reinterpret_cast<const T*>(A.tgt) + (size_t)bb * A.tbs + bl);
            lacc += (double)e * (double)e;
          } else if (A.epilogue == FSWEEP_EPI_ABS) {
            reinterpret_cast<T*>(A.y)[(size_t)bb * A.ybs + ooff] = abs_t(ox, oy);
          } else {
            st_cx(reinterpret_cast<cx<T>*>(A.y) + (size_t)bb * A.ybs + ooff, mk<T>(ox, oy));
          }
        } else {
          V2 go = f2(T(0), T(0));
          if (A.epilogue == FSWEEP_EPI_NONE) {
            const cx<T> g = ld_cx(reinterpret_cast<const cx<T>*>(A.gy) + (size_t)bb * A.gybs + ooff);
            go = f2(g.x, g.y);
          } else {
            const T mag = abs_t(ox, oy);
            T gabs;
            if (epi_fused(A.epilogue)) {
              const T e = mag - __ldg(reinterpret_cast<const T*>(A.tgt) + (size_t)bb * A.tbs + bl);
              lacc += (double)e * (double)e;
              gabs = (T)(2.0 * A.crit_scale) * e;
            } else {
              gabs = __ldg(reinterpret_cast<const T*>(A.gy) + (size_t)bb * A.gybs + ooff);
            }
            if (mag > T(0)) {
              const T s = gabs * rcp_t(mag);
              go = f2(s * ox, s * oy);
            }
          }
          xbx += go.x * xv.x + go.y * xv.y;  // g conj(x)
          xby += go.y * xv.x - go.x * xv.y;
          if (A.gx != nullptr)
            st_cx(reinterpret_cast<cx<T>*>(A.gx) + (size_t)bb * A.gxbs + ooff,
                  mk<T>(go.x * S.x - go.y * S.y, go.x * S.y + go.y * S.x));
        }
      }
      if constexpr (BWD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          xbx += __shfl_xor_sync(FULL, xbx, o);
          xby += __shfl_xor_sync(FULL, xby, o);
        }
        if (lane == 0) sRed[2 + warp] = f2(xbx, xby);
      }
    }
    if constexpr (BWD) {
      __syncthreads();
      // ---- 7. this bin's gradient contributions, with T[j] = sum_b g_b conj(y_b[j]) = conj(z_j) xbar:
      //         dW[m][j] += Re(vd[m] T[j]),  dw_post[m] += Re T[m],  dw_pre[m] += Re(vd[m] xbar)
      V2 xb = f2(T(0), T(0));
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) {
        xb.x += sRed[2 + w].x;
        xb.y += sRed[2 + w].y;
      }
#pragma unroll
      for (int i = 0; i < NGW; ++i) {
        const int e = t + NT * i, m = e >> 6, j = e & 63;
        const V2 vd = sV[m], tj = cmulj2(sZ[j], xb);
        if constexpr (TC)
          gwr[i] = fma(vd.x, tj.x, fma(-vd.y, tj.y, gwr[i]));
        else
          sGw[e] = fma(vd.x, tj.x, fma(-vd.y, tj.y, sGw[e]));
      }
      if (t < CN) {
        const V2 vd = sV[t], tj = cmulj2(sZ[t], xb);
        gpost += tj.x;
        gpre = fma(vd.x, xb.x, fma(-vd.y, xb.y, gpre));
      }
    }
    __syncthreads();
  }

  if constexpr (BWD) {
    // partial[(op.row_off + i) * G + row] for entry i of row `row` (fsweep_finalize_kernel's layout)
    T* partial = reinterpret_cast<T*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
    const OpK& preop = P.ops[L.pre];
    const OpK& postop = P.ops[L.post];
    if (fbop.acc_mode == ACC_SMEM) {
#pragma unroll
      for (int i = 0; i < NGW; ++i) {
        const int e = t + NT * i, m = e >> 6, j = e & 63;
        if (m < N && j < N) partial[(fbop.row_off + j) * G + m] = TC ? gwr[i] : sGw[e];
      }
    }
    if (t < N) {
      if (preop.acc_mode == ACC_SMEM) partial[preop.row_off * G + t] = gpre;       // N x 1: row t, entry 0
      if (postop.acc_mode == ACC_SMEM) partial[(postop.row_off + t) * G] = gpost;  // 1 x N: row 0, entry t
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<T>(lacc, A.loss_partial);
  if constexpr (TC) tc::teardown(TS);
}

cudaError_t launch_cta(int dtype, bool bwd, bool tc, int grid, cudaStream_t st, const ProgK& P, const LoopInfo& L, const SweepArgs& A, int G);
cudaError_t occupancy_cta(int dtype, bool bwd, bool tc, int* blocks_per_sm);

}  // namespace fsweep
