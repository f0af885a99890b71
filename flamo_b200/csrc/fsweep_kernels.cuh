// fsweep_kernels.cuh — the bin-sweep kernels (sm_100a), register-resident path for loop widths <= 32.
//
// Mapping: one GROUP of G lanes (G = 1,2,4,...,32, a power of two >= the widest channel count of
// the program) owns one frequency bin; lane r of the group owns ROW r of every per-bin matrix and
// element r of every per-bin vector.  Dense "H x" products broadcast x_n from lane n with
// __shfl_sync(width=G); adjoint products "H^H g" are butterfly reductions.  The closed loop
// (I - F*Fb) is built row-distributed in registers, LU-factored with implicit partial pivoting
// (pivot lane chosen by an arg-max butterfly; rows never move), and reused for every batch item
// and trailing column of that bin.  Nothing per-bin is ever written to HBM except y (or g_x).
//
// Backward = recompute-in-backward: the forward states are re-derived per bin, parked in
// thread-private shared-memory slots, and walked in reverse.  The recursion adjoint uses
//   lambda = A^-H g_y,  u = x + Fb*y,  g_F = lambda u^H,  g_Fb = (F^H lambda) y^H,  g_x = F^H lambda
// (SURVEY.md appendix A), i.e. two vector-valued chain back-propagations and one adjoint solve with
// the same LU.  Coefficient gradients are accumulated in thread-private shared-memory columns
// (no atomics, fixed order -> deterministic), reduced per block, and summed over blocks in float64
// by fsweep_finalize_kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fsweep.h"
#include "fsweep_pdl.cuh"

#include <type_traits>

namespace fsweep {

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E) — guarantees static register
// indices where `#pragma unroll` of nested loops is only a hint
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}
// descending: I = E-1 ... B
template <int B, int E, typename F>
__device__ __forceinline__ void static_rfor(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, E - 1>{});
    static_rfor<B, E - 1>(f);
  }
}

constexpr int MAX_OPS = 24;
constexpr int BLOCK = 128;
constexpr unsigned FULL = 0xffffffffu;

// accumulator placement for one op's coefficient gradient
constexpr int ACC_NONE = 0;    // no gradient wanted
constexpr int ACC_SMEM = 1;    // thread-private shared-memory column, reduced at kernel end
constexpr int ACC_GLOBAL = 2;  // too large for shared memory: atomicAdd straight into the global accumulator
constexpr int ACC_TABLE = 3;   // TABLE kinds: per-bin gradient written directly

struct OpK {
  int kind, n_out, n_in, K;
  unsigned flags;
  int acc_mode;
  int row_off;  // ACC_SMEM: offset of this op's per-lane accumulator row
  int row_len;  // accumulators per row (per lane)
  int acc_off;  // offset of this op in the flat accumulator / partial buffers (row-major [row][row_len])
  int h_off;    // offset (complex elements per thread) of this op's response row in the per-thread cache
  int def_off;  // >= 0: DEFERRED gradient — the backward kernel only records this op's input state and output
                // gradient per bin (offset, in complex elements, inside one record of the deferral buffer) and
                // fsweep_sos_defer_kernel forms the coefficient gradient afterwards; -1 otherwise
  const void* coef;
  void* gtab;  // ACC_TABLE: gradient table
  // per-item coefficient sets (fsweep_op_t::per_item, the generic kernels only): bytes between the sets of consecutive
  // batch items in `coef` / `gtab`; 0 when every item shares one set
  long long coef_is, gtab_is;
};

// The host lowers the op list into small step tables so that every kernel has exactly ONE call
// site of apply_op / backprop_op (keeps code size and compile time bounded).
struct Step {
  unsigned char op;     // index into ProgK::ops
  unsigned char flags;  // ST_* (forward tables) or RS_* (reverse table)
};
// forward-direction step flags
constexpr unsigned ST_SAVE = 1;    // park the op input in the next saved-state slot (backward only)
constexpr unsigned ST_SAVE_X = 2;  // before the op: park the state in slot X (recursion input x)
constexpr unsigned ST_SOLVE = 4;   // after the op: state <- A^-1 state
constexpr unsigned ST_ADD_X = 8;   // after the op: state += slot X            (u = x + Fb y)
constexpr unsigned ST_IDENT = 16;  // matrix build: the state is still the identity
// reverse-direction step flags
constexpr unsigned RS_NEED_GIN = 1;   // propagate the gradient to the op input
constexpr unsigned RS_ADJ = 2;        // before the op: g <- A^-H g
constexpr unsigned RS_SAVE_G = 4;     // after the op: park g in slot X          (g_u)
constexpr unsigned RS_RESTORE_G = 8;  // after the op: g <- slot X

constexpr int MAX_STEPS = 2 * MAX_OPS;

struct ProgK {
  int n_ops;
  int rec_n;  // loop width (0: no recursion)
  int in_ch, out_ch;
  int n_slots;       // backward: saved-state slots (the last one is slot X)
  int acc_per_lane;  // backward: shared-memory accumulators per lane
  int acc_total;     // backward: flat accumulator count (sum over ops of rows*row_len)
  int n_fsteps, n_msteps, n_bsteps, n_rsteps;
  int h_total;     // complex elements per thread in the response cache
  int needs_ctx;   // some op is a section cascade: the Taylor context (v, v^2) is needed
  int def_stride;  // complex elements per (column, bin) record of the deferral buffer (0: nothing deferred)
  long long nfft;
  double lng;   // ln(gamma)
  double inv_nfft;
  double gm1;   // gamma - 1
  double g2m1;  // gamma^2 - 1
  OpK ops[MAX_OPS];
  Step fsteps[MAX_OPS];    // forward kernel: signal path
  Step msteps[MAX_OPS];    // loop-matrix build: Fb chain then F chain, applied to the identity
  Step bsteps[MAX_STEPS];  // backward kernel: forward recompute with saves
  Step rsteps[MAX_STEPS];  // backward kernel: reverse sweep
};

// ------------------------------------------------------------------------------------------ complex
template <typename T>
struct cx {
  T x, y;
};
template <typename T>
__device__ __forceinline__ cx<T> mk(T a, T b) {
  cx<T> r;
  r.x = a;
  r.y = b;
  return r;
}
template <typename T>
__device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T>
__device__ __forceinline__ cx<T> cmulc(cx<T> a, cx<T> b) {  // a * conj(b)
  return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T>
__device__ __forceinline__ void cfma(cx<T>& acc, cx<T> a, cx<T> b) {  // acc += a*b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
template <typename T>
__device__ __forceinline__ void cfnma(cx<T>& acc, cx<T> a, cx<T> b) {  // acc -= a*b
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
template <typename T>
__device__ __forceinline__ void cfmac(cx<T>& acc, cx<T> a, cx<T> b) {  // acc += a*conj(b)
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  acc.y = fma(-a.x, b.y, acc.y);
}
template <typename T>
__device__ __forceinline__ void cfmacj(cx<T>& acc, cx<T> a, cx<T> b) {  // acc += conj(a)*b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ float rcp_t(float d) { return __fdividef(1.0f, d); }  // MUFU.RCP, <= 2 ulp
__device__ __forceinline__ double rcp_t(double d) { return 1.0 / d; }
template <typename T>
__device__ __forceinline__ cx<T> crcp(cx<T> a) {
  T d = rcp_t(a.x * a.x + a.y * a.y);
  return mk<T>(a.x * d, -a.y * d);
}
template <typename T>
__device__ __forceinline__ cx<T> crcp_exact(cx<T> a) {  // correctly rounded reciprocal: section cascades multiply 30 of these
  T d = T(1) / (a.x * a.x + a.y * a.y);
  return mk<T>(a.x * d, -a.y * d);
}
template <typename T>
__device__ __forceinline__ bool czero(cx<T> a) {
  return a.x == T(0) && a.y == T(0);
}

__device__ __forceinline__ cx<float> ld_cx(const cx<float>* p) {
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return mk<float>(v.x, v.y);
}
__device__ __forceinline__ cx<double> ld_cx(const cx<double>* p) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return mk<double>(v.x, v.y);
}
__device__ __forceinline__ void st_cx(cx<float>* p, cx<float> v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }
__device__ __forceinline__ void st_cx(cx<double>* p, cx<double> v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
}

// ---- group exchange.  G <= 32: warp shuffles.  G == 64: the group is two warps of the same CTA (threads
// 64g .. 64g+63); values cross through a static shared-memory buffer, ordered by a 64-thread named barrier
// (id 1 + g).  All callers keep control flow uniform inside a group, so the barriers always match up.
constexpr int XV = 4;  // widest vector exchanged at once
__device__ __forceinline__ void gbar64() { asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory"); }
template <typename T>
__device__ __forceinline__ cx<T>* xbuf64() {
  __shared__ double2 buf[BLOCK / 64][XV][64];
  return reinterpret_cast<cx<T>*>(&buf[threadIdx.x >> 6][0][0]);
}

template <int G, typename T>
__device__ __forceinline__ cx<T> shfl(cx<T> v, int src) {
  if constexpr (G == 64) {
    cx<T>* b = xbuf64<T>();
    b[threadIdx.x & 63] = v;
    gbar64();
    cx<T> r = b[src & 63];
    gbar64();
    return r;
  } else {
    return mk<T>(__shfl_sync(FULL, v.x, src, G), __shfl_sync(FULL, v.y, src, G));
  }
}
// NC values from lane `src` in one exchange
template <int G, typename T, int NC>
__device__ __forceinline__ void shflv(const cx<T> (&v)[NC], int src, cx<T> (&out)[NC]) {
  if constexpr (G == 64) {
    static_assert(NC <= XV, "vector exchange too wide");
    cx<T>* b = xbuf64<T>();
    static_for<0, NC>([&](auto cc) { b[decltype(cc)::value * 64 + (threadIdx.x & 63)] = v[decltype(cc)::value]; });
    gbar64();
    static_for<0, NC>([&](auto cc) { out[decltype(cc)::value] = b[decltype(cc)::value * 64 + (src & 63)]; });
    gbar64();
  } else {
    static_for<0, NC>([&](auto cc) { out[decltype(cc)::value] = shfl<G>(v[decltype(cc)::value], src); });
  }
}
template <int G, typename T>
__device__ __forceinline__ cx<T> group_sum(cx<T> v) {
  if constexpr (G == 64) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(FULL, v.x, o);
      v.y += __shfl_xor_sync(FULL, v.y, o);
    }
    cx<T>* b = xbuf64<T>();
    if ((threadIdx.x & 31) == 0) b[(threadIdx.x >> 5) & 1] = v;
    gbar64();
    cx<T> r = mk<T>(b[0].x + b[1].x, b[0].y + b[1].y);
    gbar64();
    return r;
  } else {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(FULL, v.x, o, G);
      v.y += __shfl_xor_sync(FULL, v.y, o, G);
    }
    return v;
  }
}
template <int G, typename T, int NC>
__device__ __forceinline__ void group_sumv(cx<T> (&v)[NC]) {
  if constexpr (G == 64) {
    static_assert(NC <= XV, "vector reduction too wide");
    static_for<0, NC>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v[c].x += __shfl_xor_sync(FULL, v[c].x, o);
        v[c].y += __shfl_xor_sync(FULL, v[c].y, o);
      }
    });
    cx<T>* b = xbuf64<T>();
    if ((threadIdx.x & 31) == 0) static_for<0, NC>([&](auto cc) { b[decltype(cc)::value * 64 + ((threadIdx.x >> 5) & 1)] = v[decltype(cc)::value]; });
    gbar64();
    static_for<0, NC>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      v[c] = mk<T>(b[c * 64].x + b[c * 64 + 1].x, b[c * 64].y + b[c * 64 + 1].y);
    });
    gbar64();
  } else {
    static_for<0, NC>([&](auto cc) { v[decltype(cc)::value] = group_sum<G>(v[decltype(cc)::value]); });
  }
}
template <int G>
__device__ __forceinline__ unsigned group_max(unsigned key) {
  if constexpr (G == 64) {
    key = __reduce_max_sync(FULL, key);
    __shared__ unsigned kb[BLOCK / 64][2];
    if ((threadIdx.x & 31) == 0) kb[threadIdx.x >> 6][(threadIdx.x >> 5) & 1] = key;
    gbar64();
    key = max(kb[threadIdx.x >> 6][0], kb[threadIdx.x >> 6][1]);
    gbar64();
    return key;
  } else if constexpr (G == 32) {
    return __reduce_max_sync(FULL, key);  // one REDUX when the group is the whole warp
  } else {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(FULL, key, o, G));
    return key;
  }
}

__device__ __forceinline__ void sincospi_t(float a, float* s, float* c) { sincospif(a, s, c); }
__device__ __forceinline__ void sincospi_t(double a, double* s, double* c) { sincospi(a, s, c); }
template <typename T>
__device__ __forceinline__ T eps_of();
template <>
__device__ __forceinline__ float eps_of<float>() {
  return 1.1920928955078125e-07f;
}
template <>
__device__ __forceinline__ double eps_of<double>() {
  return 2.220446049250313e-16;
}

// ------------------------------------------------------------------------------- per-bin context
template <typename T>
struct Ctx {
  long long k;   // absolute bin index
  T omega;       // 2*pi*k/nfft
  cx<T> u1;      // v  = gamma*z^-1 -/+ 1   (expansion point w = +1 if plus else -1)
  cx<T> u2;      // v^2
  bool plus;     // low half of the spectrum: expand section quadratics around w = +1
  long long nfft;
  double lng;
  double inv_nfft;
  long long item;  // batch item whose coefficient set this block reads (per-item launches: blockIdx.y; else 0)
};

template <typename T>
__device__ __forceinline__ Ctx<T> make_ctx(const ProgK& P, long long k, int item = 0) {
  Ctx<T> c;
  c.k = k;
  c.item = item;
  c.nfft = P.nfft;
  c.lng = P.lng;
  c.inv_nfft = P.inv_nfft;
  double fr = (double)(2 * k) * c.inv_nfft;  // omega/pi in [0,1]
  c.omega = (T)(fr * 3.141592653589793238462643383279502884);
  c.plus = true;
  c.u1 = mk<T>(0, 0);
  c.u2 = mk<T>(0, 0);
  if (P.needs_ctx) {
    T f = (T)fr;
    T s, co, sh, ch;
    sincospi_t(f, &s, &co);
    sincospi_t(f * T(0.5), &sh, &ch);
    T g = (T)(P.gm1 + 1.0);
    c.plus = co >= T(0);
    // w - 1 = (g-1) - 2 g sin^2(w/2) - j g sin w ;  w + 1 = 2 g cos^2(w/2) - (g-1) - j g sin w
    if (c.plus)
      c.u1 = mk<T>((T)P.gm1 - T(2) * g * sh * sh, -g * s);
    else
      c.u1 = mk<T>(T(2) * g * ch * ch - (T)P.gm1, -g * s);
    c.u2 = cmul(c.u1, c.u1);
  }
  return c;
}

// ---------------------------------------------------------------------------------- op responses
__device__ __forceinline__ bool is_dense(int kind) {
  return kind == FSWEEP_OP_GAIN || kind == FSWEEP_OP_SOS || kind == FSWEEP_OP_DELAY || kind == FSWEEP_OP_TABLE;
}

// coefficient set of the block's batch item (ctx.item is the constant 0 outside the per-item launches)
template <typename U, typename T>
__device__ __forceinline__ const U* coef_of(const OpK& op, const Ctx<T>& ctx) {
  return reinterpret_cast<const U*>(reinterpret_cast<const char*>(op.coef) + ctx.item * op.coef_is);
}
template <typename T>
__device__ __forceinline__ T* gtab_of(const OpK& op, const Ctx<T>& ctx) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(op.gtab) + ctx.item * op.gtab_is);
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, T (&c)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&c)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w;
  c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<double>(const double* p, double (&c)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p) + i);
    c[2 * i] = a.x;
    c[2 * i + 1] = a.y;
  }
}

// eight coefficients stored as TS, used as T (mixed precision: float64 coefficients, float32 gradient arithmetic)
template <typename T, typename TS>
__device__ __forceinline__ void load8_as(const TS* p, T (&c)[8]) {
  TS t[8];
  load8<TS>(p, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = (T)t[i];
}

// one packed section block {B(w0), B'(w0), b2, -, A(w0), A'(w0), a2, -} -> B(w), A(w) with
// B(w) = B(w0) + B'(w0) v + b2 v^2, v = w - w0, w0 = +1 | -1: no cancellation between large terms
template <typename T>
__device__ __forceinline__ void section_eval(const T (&c)[8], const Ctx<T>& ctx, cx<T>& Bv, cx<T>& Av) {
  Bv = mk<T>(fma(c[2], ctx.u2.x, fma(c[1], ctx.u1.x, c[0])), fma(c[2], ctx.u2.y, c[1] * ctx.u1.y));
  Av = mk<T>(fma(c[6], ctx.u2.x, fma(c[5], ctx.u1.x, c[4])), fma(c[6], ctx.u2.y, c[5] * ctx.u1.y));
}

// H = prod B / prod A with the reference's zero guard (dsp.py:1524-1525).  Accumulated as a product
// of per-section ratios B_s/A_s (each O(1)) — the plain products under/overflow float32 for long
// cascades (30-band GEQ: prod A ~ 1e-60 at DC).  prod A == 0 exactly iff some A_s == 0.
template <typename T>
__device__ __forceinline__ cx<T> sos_eval(const T* p, int K, long stride, const Ctx<T>& ctx, bool& guarded) {
  cx<T> H = mk<T>(1, 0);
  guarded = false;
  p += ctx.plus ? 0 : 8;
  for (int s = 0; s < K; ++s, p += stride) {
    T c[8];
    load8<T>(p, c);
    cx<T> Bv, Av;
    section_eval<T>(c, ctx, Bv, Av);
    guarded |= czero(Av);
    H = cmul(H, cmul(Bv, crcp_exact(Av)));
  }
  if (guarded) return mk<T>(eps_of<T>(), T(0));
  return H;
}

// float64 only, K <= 64: numerator and denominator products separately and ONE division (the range of float64 holds
// 64 sections; float32 needs the product of ratios above).  Half the FP64 instructions of sos_eval: the response-table
// kernel of a float32 model swept in float64 arithmetic (sweep._wants_f64) is FP64-pipe bound.
__device__ __forceinline__ cx<double> sos_eval_numden(const double* p, int K, long stride, const Ctx<double>& ctx,
                                                       bool& guarded) {
  cx<double> num = mk<double>(1, 0), den = mk<double>(1, 0);
  guarded = false;
  p += ctx.plus ? 0 : 8;
  for (int s = 0; s < K; ++s, p += stride) {
    double c[8];
    load8<double>(p, c);
    cx<double> Bv, Av;
    section_eval<double>(c, ctx, Bv, Av);
    guarded |= czero(Av);
    num = cmul(num, Bv);
    den = cmul(den, Av);
  }
  if (guarded) return mk<double>(eps_of<double>(), 0.0);
  return cmul(num, crcp_exact(den));
}

// cascade with section `skip` left out (rare path: a numerator section is exactly zero at this bin)
template <typename T>
__device__ __forceinline__ cx<T> sos_eval_without(const T* p, int K, long stride, const Ctx<T>& ctx, int skip) {
  cx<T> H = mk<T>(1, 0);
  p += ctx.plus ? 0 : 8;
  for (int s = 0; s < K; ++s, p += stride) {
    if (s == skip) continue;
    T c[8];
    load8<T>(p, c);
    cx<T> Bv, Av;
    section_eval<T>(c, ctx, Bv, Av);
    H = cmul(H, cmul(Bv, crcp_exact(Av)));
  }
  return H;
}

__device__ __forceinline__ float exp_t(float a) { return expf(a); }
__device__ __forceinline__ double exp_t(double a) { return exp(a); }
__device__ __forceinline__ float abs_t(float a, float b) { return sqrtf(fmaf(a, a, b * b)); }
__device__ __forceinline__ double abs_t(double a, double b) { return hypot(a, b); }

// H = gamma^d * exp(-j omega_k d).  Integer delays: phase index (k*d mod nfft) formed exactly
// (in float64 while k*d < 2^53, else in 64-bit integers); fractional: k*d/nfft range-reduced in float64.
template <typename T>
__device__ __forceinline__ cx<T> delay_eval(double d, unsigned flags, const Ctx<T>& ctx) {
  double fr;
  double dd = d;
  if (flags & FSWEEP_F_ISINT) {
    dd = rint(d);
    double t = (double)ctx.k * dd;
    if (fabs(t) < 4.0e15) {
      double q = floor(t * ctx.inv_nfft);
      double r = fma(-q, (double)ctx.nfft, t);  // exact: t and q*nfft are integers < 2^53
      if (r < 0.0) r += (double)ctx.nfft;
      if (r >= (double)ctx.nfft) r -= (double)ctx.nfft;
      fr = 2.0 * r * ctx.inv_nfft;
    } else {
      long long di = llrint(d);
      unsigned long long a = (unsigned long long)(di < 0 ? -di : di);
      long long idx = (long long)(((unsigned long long)ctx.k * a) % (unsigned long long)ctx.nfft);
      fr = (di < 0 ? -2.0 : 2.0) * (double)idx * ctx.inv_nfft;
    }
  } else {
    double t = (double)ctx.k * d * ctx.inv_nfft;
    t -= floor(t);
    fr = 2.0 * t;
  }
  T s, c;
  sincospi_t((T)fr, &s, &c);
  T mag = exp_t((T)(ctx.lng * dd));
  return mk<T>(mag * c, -mag * s);
}

// accumulator sink handed to the gradient routines
template <typename T>
struct Acc {
  T* sacc;   // shared: [acc_per_lane][BLOCK]
  T* gacc;   // global flat accumulator (ACC_GLOBAL ops)
  int tid;
  bool valid;  // this group holds a real bin
  // deferred gradients: record (q, bin) at defer[((q * n_bins + bl) * def_stride) ...]
  cx<T>* defer;
  long long bl, n_bins;
  int q0, ncols_total, def_stride;
  __device__ __forceinline__ void add(const OpK& op, int row, int e, T v) const {
    if (!valid) return;
    if (op.acc_mode == ACC_SMEM) {
      T* p = sacc + (size_t)(op.row_off + e) * BLOCK + tid;
      *p += v;
    } else if (op.acc_mode == ACC_GLOBAL) {
      atomicAdd(gacc + op.acc_off + (size_t)row * op.row_len + e, v);
    }
  }
};

// gradient of one (pair) cascade: gh = dL/dH (complex), H the cascade value
template <typename T>
__device__ __forceinline__ void sos_grad(const OpK& op, const T* p, long stride, const Ctx<T>& ctx, cx<T> H,
                                         cx<T> gh, const Acc<T>& acc, int row, int e0) {
  const int K = op.K;
  const T* p0 = p;
  p += ctx.plus ? 0 : 8;
  const int blk = ctx.plus ? 0 : 8;
  for (int s = 0; s < K; ++s, p += stride) {
    T c[8];
    load8<T>(p, c);
    cx<T> Bv, Av;
    section_eval<T>(c, ctx, Bv, Av);
    cx<T> qb;
    if (czero(Bv)) {
      // dH/dB_s = prod_{t != s} (B_t/A_t) / A_s  (H itself is 0 at this bin)
      qb = cmul(sos_eval_without<T>(p0, K, stride, ctx, s), crcp_exact(Av));
    } else {
      qb = cmul(H, crcp_exact(Bv));
    }
    cx<T> qa = cmul(H, crcp_exact(Av));
    qa.x = -qa.x;
    qa.y = -qa.y;
    cx<T> rb = cmulc(gh, qb), ra = cmulc(gh, qa);
    int e = e0 + s * (op.row_len / K) + blk;  // row_len = K * per-section stride of this row
    acc.add(op, row, e + 0, rb.x);
    acc.add(op, row, e + 1, rb.x * ctx.u1.x + rb.y * ctx.u1.y);
    acc.add(op, row, e + 2, rb.x * ctx.u2.x + rb.y * ctx.u2.y);
    acc.add(op, row, e + 4, ra.x);
    acc.add(op, row, e + 5, ra.x * ctx.u1.x + ra.y * ctx.u1.y);
    acc.add(op, row, e + 6, ra.x * ctx.u2.x + ra.y * ctx.u2.y);
  }
}

// One entry H[m][n] of a dense op's response (single switch; called from runtime loops over n).
template <typename T>
__device__ __forceinline__ cx<T> op_entry(const OpK& op, const Ctx<T>& ctx, int m, int n, bool& guard) {
  guard = false;
  switch (op.kind) {
    case FSWEEP_OP_GAIN:
      return mk<T>(__ldg(coef_of<T>(op, ctx) + m * op.n_in + n), T(0));
    case FSWEEP_OP_DELAY:
      return delay_eval<T>(__ldg(coef_of<double>(op, ctx) + m * op.n_in + n), op.flags, ctx);
    case FSWEEP_OP_SOS:
      return sos_eval<T>(coef_of<T>(op, ctx) + ((size_t)n * op.n_out + m) * 16, op.K,
                         (long)op.n_in * op.n_out * 16, ctx, guard);
    case FSWEEP_OP_TABLE: {
      const T* t = coef_of<T>(op, ctx) + 2 * (((size_t)ctx.k * op.n_out + m) * op.n_in + n);
      return mk<T>(__ldg(t), __ldg(t + 1));
    }
    default:
      return mk<T>(0, 0);
  }
}

template <typename T>
__device__ __forceinline__ cx<T> op_diag(const OpK& op, const Ctx<T>& ctx, int m, bool& guard) {
  guard = false;
  if (m >= op.n_out) return mk<T>(0, 0);
  switch (op.kind) {
    case FSWEEP_OP_PGAIN:
      return mk<T>(__ldg(coef_of<T>(op, ctx) + m), T(0));
    case FSWEEP_OP_PDELAY:
      return delay_eval<T>(__ldg(coef_of<double>(op, ctx) + m), op.flags, ctx);
    case FSWEEP_OP_PSOS:
      return sos_eval<T>(coef_of<T>(op, ctx) + (size_t)m * 16, op.K, (long)op.n_out * 16, ctx, guard);
    case FSWEEP_OP_PTABLE: {
      const T* t = coef_of<T>(op, ctx) + 2 * ((size_t)ctx.k * op.n_out + m);
      return mk<T>(__ldg(t), __ldg(t + 1));
    }
    default:
      return mk<T>(0, 0);
  }
}

__device__ __forceinline__ bool bin_invariant(int kind) { return kind == FSWEEP_OP_GAIN || kind == FSWEEP_OP_PGAIN; }

// Per-thread response cache in shared memory: hc[(op.h_off + n) * BLOCK + tid] = H[lane][n] of the
// current bin (diagonal kinds: one entry).  Bin-invariant ops (GAIN, PGAIN) are staged once per
// kernel, everything else once per bin; apply / backprop then only read the cache.
template <typename T>
__device__ __forceinline__ void stage_op(const OpK& op, const Ctx<T>& ctx, int lane, cx<T>* hc, unsigned* gmask,
                                         int opi, int tid) {
  unsigned gm = 0;
  const bool live = lane < op.n_out;
  if (is_dense(op.kind)) {
    for (int n = 0; n < op.n_in; ++n) {
      bool gd = false;
      cx<T> h = mk<T>(0, 0);
      if (live) h = op_entry<T>(op, ctx, lane, n, gd);
      if (gd) gm |= 1u << n;
      hc[(size_t)(op.h_off + n) * BLOCK + tid] = h;
    }
  } else {
    bool gd = false;
    hc[(size_t)op.h_off * BLOCK + tid] = op_diag<T>(op, ctx, lane, gd);
    gm = gd ? 1u : 0u;
  }
  gmask[opi * BLOCK + tid] = gm;
}

template <typename T>
__device__ __forceinline__ void stage_ops(const ProgK& P, const Ctx<T>& ctx, int lane, cx<T>* hc, unsigned* gmask,
                                          int tid, bool invariant_pass, int skip = -1) {
  for (int i = 0; i < P.n_ops; ++i)
    if (i != skip && bin_invariant(P.ops[i].kind) == invariant_pass) stage_op<T>(P.ops[i], ctx, lane, hc, gmask, i, tid);
}

// S <- H S for NC columns (row-distributed).  `ident`: S is the identity, so H S = H (dense only).
template <typename T, int G, int NC>
__device__ __forceinline__ void apply_op(const OpK& op, int lane, cx<T> (&S)[NC], int ncols, bool ident,
                                         const cx<T>* hc, int tid) {
  const cx<T>* hrow = hc + (size_t)op.h_off * BLOCK + tid;
  if (is_dense(op.kind)) {
    if (ident) {
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        S[c] = (c < op.n_in) ? hrow[(size_t)c * BLOCK] : mk<T>(0, 0);
      });
      return;
    }
    // column c of H S depends on column c of S only, so the columns are processed in place, CB at a time
    // (keeps the accumulators at CB registers even for the NC == G matrix build)
    constexpr int CB = NC < XV ? NC : XV;
    static_for<0, NC / CB>([&](auto bc) {
      constexpr int c0 = decltype(bc)::value * CB;
      cx<T> acc[CB], blk[CB], v[CB];
      static_for<0, CB>([&](auto cc) {
        acc[decltype(cc)::value] = mk<T>(0, 0);
        blk[decltype(cc)::value] = S[c0 + decltype(cc)::value];
      });
      for (int n = 0; n < op.n_in; ++n) {
        const cx<T> h = hrow[(size_t)n * BLOCK];
        shflv<G>(blk, n, v);
        static_for<0, CB>([&](auto cc) { cfma(acc[decltype(cc)::value], h, v[decltype(cc)::value]); });
      }
      static_for<0, CB>([&](auto cc) { S[c0 + decltype(cc)::value] = acc[decltype(cc)::value]; });
    });
  } else {
    const cx<T> h = hrow[0];
    static_for<0, NC>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      if (c < ncols) S[c] = cmul(h, S[c]);
    });
  }
}

// Back-propagate through one op: accumulate its coefficient gradient from (Sin, g) and replace g by
// H^H g (if need_gin).  Sin = the op's forward input, g = dL/d(output), both row-distributed.
template <typename T, int G, int NC>
__device__ __forceinline__ void backprop_op(const OpK& op, const Ctx<T>& ctx, int lane, const cx<T> (&Sin)[NC],
                                            cx<T> (&g)[NC], bool need_gin, const Acc<T>& acc, bool first_chunk,
                                            const cx<T>* hc, unsigned gmask, int tid) {
  bool want = op.acc_mode != ACC_NONE;
  const cx<T>* hrow = hc + (size_t)op.h_off * BLOCK + tid;
  if (want && op.def_off >= 0) {
    // deferred: lane r parks S_in[r] and g_out[r] of every live column; the per-section gradient work happens
    // in fsweep_sos_defer_kernel, parallel over (channel pair, bin) instead of serial per lane
    if (acc.valid) {
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const int q = acc.q0 + c;
        if (q < acc.ncols_total) {
          cx<T>* rec = acc.defer + ((size_t)q * acc.n_bins + acc.bl) * acc.def_stride + op.def_off;
          if (lane < op.n_in) st_cx(rec + lane, Sin[c]);
          if (lane < op.n_out) st_cx(rec + op.n_in + lane, g[c]);
        }
      });
    }
    want = false;
  }
  if (is_dense(op.kind)) {
    if (want) {
      for (int n = 0; n < op.n_in; ++n) {
        cx<T> gh = mk<T>(0, 0);
        {
          cx<T> v[NC];
          shflv<G>(Sin, n, v);
          static_for<0, NC>([&](auto cc) { cfmac(gh, g[decltype(cc)::value], v[decltype(cc)::value]); });
        }
        if (lane < op.n_out) {
          switch (op.kind) {
            case FSWEEP_OP_GAIN:
              acc.add(op, lane, n, gh.x);
              break;
            case FSWEEP_OP_DELAY:
              if (!(op.flags & FSWEEP_F_ISINT)) {
                // dH/dd = (ln g - j w) H
                cx<T> t = cmul(mk<T>((T)ctx.lng, -ctx.omega), hrow[(size_t)n * BLOCK]);
                acc.add(op, lane, n, gh.x * t.x + gh.y * t.y);
              }
              break;
            case FSWEEP_OP_SOS:
              if (!((gmask >> n) & 1u))
                sos_grad<T>(op, coef_of<T>(op, ctx) + ((size_t)n * op.n_out + lane) * 16,
                            (long)op.n_in * op.n_out * 16, ctx, hrow[(size_t)n * BLOCK], gh, acc, lane, n * 16);
              break;
            case FSWEEP_OP_TABLE:
              if (acc.valid && op.gtab) {
                T* t = gtab_of<T>(op, ctx) + 2 * (((size_t)ctx.k * op.n_out + lane) * op.n_in + n);
                if (first_chunk) {
                  t[0] = gh.x;
                  t[1] = gh.y;
                } else {
                  t[0] += gh.x;
                  t[1] += gh.y;
                }
              }
              break;
            default:
              break;
          }
        }
      }
    }
    if (need_gin) {
      cx<T> gin[NC];
      static_for<0, NC>([&](auto cc) { gin[decltype(cc)::value] = mk<T>(0, 0); });
      for (int n = 0; n < op.n_in; ++n) {
        const cx<T> h = hrow[(size_t)n * BLOCK];
        cx<T> t[NC];
        static_for<0, NC>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          t[c] = mk<T>(0, 0);
          cfmacj(t[c], h, g[c]);  // conj(h) * g
        });
        group_sumv<G>(t);
        static_for<0, NC>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          gin[c].x = (lane == n) ? t[c].x : gin[c].x;
          gin[c].y = (lane == n) ? t[c].y : gin[c].y;
        });
      }
      static_for<0, NC>([&](auto cc) { g[decltype(cc)::value] = gin[decltype(cc)::value]; });
    }
  } else {
    const cx<T> h = hrow[0];
    if (want && lane < op.n_out) {
      cx<T> gh = mk<T>(0, 0);
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        cfmac(gh, g[c], Sin[c]);
      });
      switch (op.kind) {
        case FSWEEP_OP_PGAIN:
          acc.add(op, lane, 0, gh.x);
          break;
        case FSWEEP_OP_PDELAY:
          if (!(op.flags & FSWEEP_F_ISINT)) {
            cx<T> t = cmul(mk<T>((T)ctx.lng, -ctx.omega), h);
            acc.add(op, lane, 0, gh.x * t.x + gh.y * t.y);
          }
          break;
        case FSWEEP_OP_PSOS:
          if (!(gmask & 1u))
            sos_grad<T>(op, coef_of<T>(op, ctx) + (size_t)lane * 16, (long)op.n_out * 16, ctx, h, gh,
                        acc, lane, 0);
          break;
        case FSWEEP_OP_PTABLE:
          if (acc.valid && op.gtab) {
            T* t = gtab_of<T>(op, ctx) + 2 * ((size_t)ctx.k * op.n_out + lane);
            if (first_chunk) {
              t[0] = gh.x;
              t[1] = gh.y;
            } else {
              t[0] += gh.x;
              t[1] += gh.y;
            }
          }
          break;
        default:
          break;
      }
    }
    if (need_gin) {
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        cx<T> t = mk<T>(0, 0);
        cfmacj(t, h, g[c]);
        g[c] = t;
      });
    }
  }
}

// ------------------------------------------------------------------------------------------- LU
__device__ __forceinline__ unsigned mag_key(float m) { return __float_as_uint(m); }
__device__ __forceinline__ unsigned mag_key(double m) { return __float_as_uint((float)fmin(m, 3.0e38)); }

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (G == 32)
    return FULL;
  else
    return ((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}

template <typename T>
__device__ __forceinline__ cx<T> csel(bool p, cx<T> a) {  // p ? a : 0, branch-free
  return mk<T>(p ? a.x : T(0), p ? a.y : T(0));
}

// G x G complex LU, row `lane` per lane.  Rows / columns beyond the live width are identity (the
// caller pads A that way), so all G steps always run: no width guards, no divergent branches.
template <typename T, int G>
struct LU {
  cx<T> a[G];   // row `lane` of A, overwritten by L multipliers (cols < mystep) and U (cols >= mystep)
  cx<T> dinv;   // 1 / U[mystep][mystep]
  int mystep;   // elimination step at which this row was the pivot row
  unsigned piv[(G + 3) / 4];  // pivot lane of every step, 8 bits each

  template <int K>
  __device__ __forceinline__ int pivot_lane() const {
    return (int)((piv[K / 4] >> (8 * (K % 4))) & 0xffu);
  }

  // Gaussian elimination with implicit partial pivoting (rows never move).  The arg-max over the
  // candidate rows is one REDUX: key = |a|^2 bits with the low log2(G) bits replaced by G-1-lane.
  __device__ __forceinline__ void factor(int lane) {
    mystep = -1;
    dinv = mk<T>(1, 0);
    static_for<0, (G + 3) / 4>([&](auto i) { piv[decltype(i)::value] = 0u; });
    __shared__ double2 rowbuf_raw[(G == 64) ? BLOCK / 64 : 1][(G == 64) ? 64 : 1];  // G == 64 pivot-row broadcast
    cx<T>* rowbuf = reinterpret_cast<cx<T>*>(&rowbuf_raw[(G == 64) ? (threadIdx.x >> 6) : 0][0]);
    (void)rowbuf;
    static_for<0, G>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      int who = 0;
      if constexpr (G > 1) {
        unsigned key = (mag_key(a[k].x * a[k].x + a[k].y * a[k].y) & ~(unsigned)(G - 1)) | (unsigned)(G - 1 - lane);
        key = (mystep < 0) ? key : 0u;
        key = group_max<G>(key);
        who = G - 1 - (int)(key & (unsigned)(G - 1));
      }
      piv[k / 4] |= (unsigned)who << (8 * (k % 4));
      if constexpr (G == 64) {
        // the pivot row crosses through shared memory once per step instead of one exchange per entry
        if (lane == who) static_for<k, G>([&](auto jc) { rowbuf[decltype(jc)::value] = a[decltype(jc)::value]; });
        gbar64();
        const cx<T> inv = crcp(rowbuf[k]);
        const bool act = (mystep < 0) && (lane != who);
        const cx<T> l = csel(act, cmul(a[k], inv));
        if (lane == who) {
          mystep = k;
          dinv = inv;
        }
        static_for<k + 1, G>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          cfnma(a[j], l, rowbuf[j]);
        });
        gbar64();
        a[k].x = act ? l.x : a[k].x;
        a[k].y = act ? l.y : a[k].y;
      } else {
      const cx<T> inv = crcp(shfl<G>(a[k], who));
      const bool act = (mystep < 0) && (lane != who);
      const cx<T> l = csel(act, cmul(a[k], inv));  // multiplier; 0 for the pivot row and finished rows
      if (lane == who) {
        mystep = k;
        dinv = inv;
      }
      static_for<k + 1, G>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        cfnma(a[j], l, shfl<G>(a[j], who));
      });
      a[k].x = act ? l.x : a[k].x;
      a[k].y = act ? l.y : a[k].y;
      }
    });
  }

  // A x = b for NC right-hand sides; b in natural row order on entry, x in natural order on exit.
  template <int NC>
  __device__ __forceinline__ void solve(int lane, cx<T> (&b)[NC]) const {
    static_for<0, G>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const int p = pivot_lane<k>();
      const cx<T> lk = csel(mystep > k, a[k]);
      cx<T> bk[NC];
      shflv<G>(b, p, bk);
      static_for<0, NC>([&](auto cc) { cfnma(b[decltype(cc)::value], lk, bk[decltype(cc)::value]); });
    });
    cx<T> x[NC];
    static_for<0, NC>([&](auto cc) { x[decltype(cc)::value] = mk<T>(0, 0); });
    static_rfor<0, G>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const int p = pivot_lane<k>();
      const cx<T> uk = csel(mystep < k, a[k]);
      cx<T> t[NC], xk[NC];
      static_for<0, NC>([&](auto cc) { t[decltype(cc)::value] = cmul(b[decltype(cc)::value], dinv); });
      shflv<G>(t, p, xk);
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        cfnma(b[c], uk, xk[c]);
        x[c].x = (lane == k) ? xk[c].x : x[c].x;
        x[c].y = (lane == k) ? xk[c].y : x[c].y;
      });
    });
    static_for<0, NC>([&](auto cc) { b[decltype(cc)::value] = x[decltype(cc)::value]; });
  }

  // A^H lam = g for NC right-hand sides (natural order in and out), reusing the same factors:
  // A = P^T L U  =>  U^H w = g ,  L^H v = w ,  lam = P^T v.
  template <int NC>
  __device__ __forceinline__ void solve_adj(int lane, cx<T> (&g)[NC]) const {
    cx<T> w[NC];
    shflv<G>(g, mystep, w);
    const cx<T> dinvc = mk<T>(dinv.x, -dinv.y);
    // forward substitution with U^H (lower triangular): w_i = (g_i - sum_{j<i} conj(U[j][i]) w_j) / conj(U[i][i])
    static_for<0, G>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      const cx<T> ui = csel(mystep < i, mk<T>(a[i].x, -a[i].y));
      cx<T> t[NC];
      static_for<0, NC>([&](auto cc) { t[decltype(cc)::value] = cmul(ui, w[decltype(cc)::value]); });
      group_sumv<G>(t);
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const cx<T> d = cmul(mk<T>(w[c].x - t[c].x, w[c].y - t[c].y), dinvc);
        w[c].x = (mystep == i) ? d.x : w[c].x;
        w[c].y = (mystep == i) ? d.y : w[c].y;
      });
    });
    // back substitution with L^H (unit upper triangular): v_j = w_j - sum_{i>j} conj(L[i][j]) v_i
    static_rfor<0, G>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      const cx<T> lj = csel(mystep > j, mk<T>(a[j].x, -a[j].y));
      cx<T> t[NC];
      static_for<0, NC>([&](auto cc) { t[decltype(cc)::value] = cmul(lj, w[decltype(cc)::value]); });
      group_sumv<G>(t);
      static_for<0, NC>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        w[c].x = (mystep == j) ? w[c].x - t[c].x : w[c].x;
        w[c].y = (mystep == j) ? w[c].y - t[c].y : w[c].y;
      });
    });
    static_for<0, NC>([&](auto cc) { g[decltype(cc)::value] = w[decltype(cc)::value]; });
  }
};

// Build A = I - F*Fb for this bin (row-distributed) and factor it.
template <typename T, int G>
__device__ __forceinline__ void build_loop(const ProgK& P, int lane, LU<T, G>& lu, const cx<T>* hc, int tid) {
  const int N = P.rec_n;
  static_for<0, G>([&](auto cc) {
    constexpr int c = decltype(cc)::value;
    lu.a[c] = mk<T>((c == lane && lane < N) ? T(1) : T(0), T(0));  // identity over the live width only
  });
  for (int i = 0; i < P.n_msteps; ++i) {
    const Step st = P.msteps[i];
    apply_op<T, G, G>(P.ops[st.op], lane, lu.a, G, (st.flags & ST_IDENT) != 0, hc, tid);
  }
  static_for<0, G>([&](auto cc) {
    constexpr int c = decltype(cc)::value;
    lu.a[c].x = (c == lane ? T(1) : T(0)) - lu.a[c].x;
    lu.a[c].y = -lu.a[c].y;
  });
  lu.factor(lane);  // rows / columns >= N of A are identity
}

// ------------------------------------------------------------------------------- kernel arguments
struct SweepArgs {
  const void* x;
  long long xbs;
  void* y;  // forward: output; backward: unused
  long long ybs;
  const void* gy;  // backward
  long long gybs;
  void* gx;  // backward, may be null
  long long gxbs;
  int batch, cols;
  long long bin_begin, n_bins;
  int epilogue;
  void* partial;  // backward: [grid][acc_per_lane * G]
  void* gacc;     // backward: [acc_total] (ACC_GLOBAL ops), zeroed by the host wrapper
  void* defer;    // backward: deferral buffer [batch*cols][n_bins][def_stride] cplx (ops with def_off >= 0)
  // fused criterion (epilogue EPI_ABS_MSE / EPI_ABSSUM_MSE, cols == 1): y / gy are unused, dL/d|Y| is formed in
  // the kernel from the target and the squared errors are summed per block
  const void* tgt;       // real (batch, n_bins, out_ch) or (batch, n_bins); first processed bin
  long long tbs;         // target batch stride in elements
  double crit_scale;     // loss = crit_scale * sum e^2
  double* loss_partial;  // [grid]
};

// internal epilogue codes of the fused criteria (the ABI passes an fsweep_criterion_t instead)
constexpr int EPI_ABS_MSE = 2;     // e = |Y_r| - t_r          (nn.MSELoss on the magnitudes)
constexpr int EPI_ABSSUM_MSE = 3;  // e = sum_r |Y_r| - t      (optimize/loss.py mse_loss)
__device__ __forceinline__ bool epi_fused(int e) { return e >= EPI_ABS_MSE; }

// Row-distributed fused criterion: `mag` = |Y_row| on lanes holding an output row (`live`), anything elsewhere.
// Returns dL/d|Y_row| (upstream gradient 1) and adds this lane's share of the squared error to lacc.
// Called with uniform control flow inside the group (the channel sum is a group reduction).
template <typename T, int G>
__device__ __forceinline__ T crit_rowdist(const SweepArgs& A, T mag, bool live, int row, int out_rows, long long bl,
                                          int b, int lane, bool count, double& lacc) {
  const T* tg = reinterpret_cast<const T*>(A.tgt) + (size_t)b * A.tbs;
  T e;
  if (A.epilogue == EPI_ABSSUM_MSE) {
    const T tot = group_sum<G>(mk<T>(live ? mag : T(0), T(0))).x;
    e = tot - __ldg(tg + bl);
    if (count && lane == 0) lacc += (double)e * (double)e;
  } else {
    e = live ? mag - __ldg(tg + (size_t)bl * out_rows + row) : T(0);
    if (count && live) lacc += (double)e * (double)e;
  }
  return (T)(2.0 * A.crit_scale) * e;
}

template <typename T>
__device__ __forceinline__ void block_loss_store(double lacc, double* loss_partial) {
  __shared__ double red[32];  // up to 1024 threads per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lacc += __shfl_xor_sync(FULL, lacc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) s += red[w];
    loss_partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

template <typename T, int G, int CC>
__device__ __forceinline__ void load_cols(const cx<T>* x, long long xbs, long long bl, int nch, int cols, int q0,
                                          int ncols_total, int lane, cx<T> (&S)[CC]) {
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    int q = q0 + c;
    S[c] = mk<T>(0, 0);
    if (q < ncols_total && lane < nch) {
      int b = (CC == 1 && cols == 1) ? q : q / cols;
      int cc = q - b * cols;
      S[c] = ld_cx(x + (size_t)b * xbs + ((size_t)bl * nch + lane) * cols + cc);
    }
  }
}

// shared memory: [hc: h_total*BLOCK cx] [gmask: n_ops*BLOCK u32] [save: n_slots*CC*BLOCK cx] [sacc: acc_per_lane*BLOCK T]
// ------------------------------------------------------------------------------------ forward
template <typename T, int G, int CC>
__global__ void __launch_bounds__(BLOCK) fsweep_fwd_kernel(const __grid_constant__ ProgK P, const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* hc = reinterpret_cast<cx<T>*>(smem_raw);
  unsigned* gmask = reinterpret_cast<unsigned*>(hc + (size_t)P.h_total * BLOCK);
  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const long long groups_total = (long long)gridDim.x * (BLOCK / G);
  const long long gg = (long long)blockIdx.x * (BLOCK / G) + tid / G;
  const long long n_iter = (A.n_bins + groups_total - 1) / groups_total;
  // per-item launches (fsweep_op_t::per_item): blockIdx.y is the batch item, its columns are the block's whole batch
  const int item = blockIdx.y;
  const int ncols_total = (gridDim.y > 1 ? 1 : A.batch) * A.cols;
  const cx<T>* x = reinterpret_cast<const cx<T>*>(A.x) + (size_t)item * A.xbs;
  double lacc = 0.0;

  {
    const Ctx<T> c0 = make_ctx<T>(P, A.bin_begin, item);
    stage_ops<T>(P, c0, lane, hc, gmask, tid, true);  // bin-invariant rows, once per kernel
  }
  for (long long it = 0; it < n_iter; ++it) {
    long long bl = it * groups_total + gg;
    const bool valid = bl < A.n_bins;
    if (!valid) bl = A.n_bins - 1;
    const Ctx<T> ctx = make_ctx<T>(P, A.bin_begin + bl, item);
    stage_ops<T>(P, ctx, lane, hc, gmask, tid, false);
    LU<T, G> lu;
    if (P.rec_n > 0) build_loop<T, G>(P, lane, lu, hc, tid);

    for (int q0 = 0; q0 < ncols_total; q0 += CC) {
      cx<T> S[CC];
      load_cols<T, G, CC>(x, A.xbs, bl, P.in_ch, A.cols, q0, ncols_total, lane, S);
      for (int i = 0; i < P.n_fsteps; ++i) {
        const Step st = P.fsteps[i];
        apply_op<T, G, CC>(P.ops[st.op], lane, S, CC, false, hc, tid);
        if (st.flags & ST_SOLVE) lu.template solve<CC>(lane, S);
      }
      if (epi_fused(A.epilogue)) {  // cols == 1: q is the batch item
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          const int q = q0 + c < ncols_total ? q0 + c : ncols_total - 1;
          crit_rowdist<T, G>(A, abs_t(S[c].x, S[c].y), lane < P.out_ch, lane, P.out_ch, bl, item + q, lane,
                             valid && q0 + c < ncols_total, lacc);
        }
      } else if (valid && lane < P.out_ch) {
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          int q = q0 + c;
          if (q < ncols_total) {
            int b = q / A.cols, cc = q - b * A.cols;
            size_t off = (size_t)(item + b) * A.ybs + ((size_t)bl * P.out_ch + lane) * A.cols + cc;
            if (A.epilogue == FSWEEP_EPI_ABS)
              reinterpret_cast<T*>(A.y)[off] = abs_t(S[c].x, S[c].y);
            else
              st_cx(reinterpret_cast<cx<T>*>(A.y) + off, S[c]);
          }
        }
      }
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<T>(lacc, A.loss_partial);
}

// ------------------------------------------------------------------------------------ backward
template <typename T, int G, int CC>
__global__ void __launch_bounds__(BLOCK) fsweep_bwd_kernel(const __grid_constant__ ProgK P, const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* hc = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* save = hc + (size_t)P.h_total * BLOCK;                           // [n_slots][CC][BLOCK]
  T* sacc = reinterpret_cast<T*>(save + (size_t)P.n_slots * CC * BLOCK);  // [acc_per_lane][BLOCK]
  unsigned* gmask = reinterpret_cast<unsigned*>(sacc + (size_t)P.acc_per_lane * BLOCK);

  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const long long groups_total = (long long)gridDim.x * (BLOCK / G);
  const long long gg = (long long)blockIdx.x * (BLOCK / G) + tid / G;
  const long long n_iter = (A.n_bins + groups_total - 1) / groups_total;
  // per-item launches (fsweep_op_t::per_item): blockIdx.y is the batch item, its columns are the block's whole batch
  const int item = blockIdx.y;
  const int ncols_total = (gridDim.y > 1 ? 1 : A.batch) * A.cols;
  const cx<T>* x = reinterpret_cast<const cx<T>*>(A.x) + (size_t)item * A.xbs;
  const int slot_x = P.n_slots - 1;
  double lacc = 0.0;

  for (int i = 0; i < P.acc_per_lane; ++i) sacc[(size_t)i * BLOCK + tid] = T(0);

  Acc<T> acc;
  acc.sacc = sacc;
  acc.gacc = reinterpret_cast<T*>(A.gacc) + (size_t)item * P.acc_total;
  acc.tid = tid;
  acc.defer = reinterpret_cast<cx<T>*>(A.defer);
  acc.n_bins = A.n_bins;
  acc.ncols_total = ncols_total;
  acc.def_stride = P.def_stride;

  auto put = [&](int slot, const cx<T>(&S)[CC]) {
#pragma unroll
    for (int c = 0; c < CC; ++c) save[((size_t)slot * CC + c) * BLOCK + tid] = S[c];
  };
  auto get = [&](int slot, cx<T>(&S)[CC]) {
#pragma unroll
    for (int c = 0; c < CC; ++c) S[c] = save[((size_t)slot * CC + c) * BLOCK + tid];
  };

  {
    const Ctx<T> c0 = make_ctx<T>(P, A.bin_begin, item);
    stage_ops<T>(P, c0, lane, hc, gmask, tid, true);
  }
  for (long long it = 0; it < n_iter; ++it) {
    long long bl = it * groups_total + gg;
    const bool valid = bl < A.n_bins;
    if (!valid) bl = A.n_bins - 1;
    acc.valid = valid;
    acc.bl = bl;
    const Ctx<T> ctx = make_ctx<T>(P, A.bin_begin + bl, item);
    stage_ops<T>(P, ctx, lane, hc, gmask, tid, false);
    LU<T, G> lu;
    if (P.rec_n > 0) build_loop<T, G>(P, lane, lu, hc, tid);

    for (int q0 = 0; q0 < ncols_total; q0 += CC) {
      const bool first_chunk = q0 == 0;
      acc.q0 = q0;
      cx<T> S[CC];
      load_cols<T, G, CC>(x, A.xbs, bl, P.in_ch, A.cols, q0, ncols_total, lane, S);
      int slot = 0;
      // ---- forward recompute, parking every op input
      for (int i = 0; i < P.n_bsteps; ++i) {
        const Step st = P.bsteps[i];
        if (st.flags & ST_SAVE_X) put(slot_x, S);
        if (st.flags & ST_SAVE) put(slot++, S);
        apply_op<T, G, CC>(P.ops[st.op], lane, S, CC, false, hc, tid);
        if (st.flags & ST_SOLVE) lu.template solve<CC>(lane, S);
        if (st.flags & ST_ADD_X) {
          cx<T> X[CC];
          get(slot_x, X);
#pragma unroll
          for (int c = 0; c < CC; ++c) {
            S[c].x += X[c].x;
            S[c].y += X[c].y;
          }
        }
      }
      // ---- output gradient
      cx<T> g[CC];
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        int q = q0 + c;
        g[c] = mk<T>(0, 0);
        if (epi_fused(A.epilogue)) {
          const bool inr = q < ncols_total;
          const T mag = abs_t(S[c].x, S[c].y);
          const T ga = crit_rowdist<T, G>(A, mag, lane < P.out_ch, lane, P.out_ch, bl, item + (inr ? q : ncols_total - 1), lane,
                                          valid && inr, lacc);
          if (inr && lane < P.out_ch && mag > T(0)) {
            const T r = ga * rcp_t(mag);
            g[c] = mk<T>(r * S[c].x, r * S[c].y);
          }
        } else if (q < ncols_total && lane < P.out_ch) {
          int b = q / A.cols, cc = q - b * A.cols;
          size_t off = (size_t)(item + b) * A.gybs + ((size_t)bl * P.out_ch + lane) * A.cols + cc;
          if (A.epilogue == FSWEEP_EPI_ABS) {
            T ga = __ldg(reinterpret_cast<const T*>(A.gy) + off);
            T mag = abs_t(S[c].x, S[c].y);
            if (mag > T(0)) {
              T r = ga * rcp_t(mag);
              g[c] = mk<T>(r * S[c].x, r * S[c].y);
            }
          } else {
            g[c] = ld_cx(reinterpret_cast<const cx<T>*>(A.gy) + off);
          }
        }
      }
      // ---- reverse sweep
      for (int i = 0; i < P.n_rsteps; ++i) {
        const Step st = P.rsteps[i];
        if (st.flags & RS_ADJ) lu.template solve_adj<CC>(lane, g);
        cx<T> Sin[CC];
        get(--slot, Sin);
        backprop_op<T, G, CC>(P.ops[st.op], ctx, lane, Sin, g, (st.flags & RS_NEED_GIN) != 0, acc, first_chunk, hc,
                              gmask[st.op * BLOCK + tid], tid);
        if (st.flags & RS_SAVE_G) put(slot_x, g);
        if (st.flags & RS_RESTORE_G) get(slot_x, g);
      }
      if (A.gx != nullptr && valid && lane < P.in_ch) {
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          int q = q0 + c;
          if (q < ncols_total) {
            int b = q / A.cols, cc = q - b * A.cols;
            st_cx(reinterpret_cast<cx<T>*>(A.gx) + (size_t)(item + b) * A.gxbs + ((size_t)bl * P.in_ch + lane) * A.cols + cc,
                  g[c]);
          }
        }
      }
    }
  }

  // ---- block reduction of the thread-private accumulator columns: partial[block][i][row]
  __syncthreads();
  T* partial = reinterpret_cast<T*>(A.partial) + ((size_t)item * gridDim.x + blockIdx.x) * P.acc_per_lane * G;
  for (int e = tid; e < P.acc_per_lane * G; e += BLOCK) {
    int i = e / G, row = e - i * G;
    T s = T(0);
    for (int j = 0; j < BLOCK / G; ++j) s += sacc[(size_t)i * BLOCK + j * G + row];
    partial[e] = s;
  }
  if (epi_fused(A.epilogue)) block_loss_store<T>(lacc, A.loss_partial);
}

// Sum the per-block partials (float64) and scatter into the caller's gradient buffers.
struct FinalizeOp {
  int kind, n_out, n_in, K;
  int acc_mode, row_off, row_len, acc_off;
  void* grad;  // caller buffer or null
  long long grad_is;  // per-item launches: elements between the gradients of consecutive items; 0: one set, summed over the items
};
struct FinalizeArgs {
  int n_ops, G, acc_per_lane, n_blocks;
  int n_items, acc_total;  // per-item launches: partial rows are [item][block], gacc is [item][acc_total]; else n_items = 1
  const void* partial;
  const void* gacc;
  // fused criterion: blockIdx.y == n_ops sums the per-block squared-error sums into *loss (real T)
  const double* loss_partial;
  void* loss;
  double crit_scale;
  FinalizeOp ops[MAX_OPS];
};

// One WARP per output element: lanes stride over the per-block partials, shuffle-reduce in float64.
template <typename T>
__global__ void __launch_bounds__(128) fsweep_finalize_kernel(const __grid_constant__ FinalizeArgs F) {
  const int opi = blockIdx.y;
  if (opi == F.n_ops) {
    if (F.loss == nullptr || blockIdx.x != 0 || blockIdx.z != 0 || threadIdx.x >= 32) return;
    double s = 0.0;
    for (int b = threadIdx.x; b < F.n_blocks * F.n_items; b += 32) s += F.loss_partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *reinterpret_cast<T*>(F.loss) = (T)(F.crit_scale * s);
    return;
  }
  const FinalizeOp& op = F.ops[opi];
  if (op.grad == nullptr || (op.acc_mode != ACC_SMEM && op.acc_mode != ACC_GLOBAL)) return;
  // per-item op: this z-slice sums its own item's rows; shared op: slice 0 sums the rows of every item
  const bool own = op.grad_is != 0;
  if (!own && blockIdx.z != 0) return;
  const int row0 = own ? (int)blockIdx.z * F.n_blocks : 0, rows = own ? F.n_blocks : F.n_blocks * F.n_items;
  const int it0 = own ? (int)blockIdx.z : 0, its = own ? 1 : F.n_items;
  const bool diag = !(op.kind == FSWEEP_OP_GAIN || op.kind == FSWEEP_OP_SOS || op.kind == FSWEEP_OP_DELAY);
  const int total = op.n_out * op.row_len;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int e = blockIdx.x * warps_per_block + (threadIdx.x >> 5); e < total; e += gridDim.x * warps_per_block) {
    int row = e / op.row_len, i = e - row * op.row_len;
    double s = 0.0;
    if (op.acc_mode == ACC_SMEM) {
      const T* p = reinterpret_cast<const T*>(F.partial) + (size_t)(op.row_off + i) * F.G + row;
      const size_t stride = (size_t)F.acc_per_lane * F.G;
      for (int b = lane; b < rows; b += 32) s += (double)p[(size_t)(row0 + b) * stride];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    } else {
      for (int t = 0; t < its; ++t) s += (double)reinterpret_cast<const T*>(F.gacc)[(size_t)(it0 + t) * F.acc_total + op.acc_off + e];
    }
    if (lane != 0) continue;
    // map (row, i) -> index in the caller's layout
    size_t o;
    if (op.kind == FSWEEP_OP_SOS) {
      // i = (s * n_in + n) * 16 + slot  ->  ((s * n_in + n) * n_out + row) * 16 + slot
      int slot = i & 15, sn = i >> 4;
      o = ((size_t)sn * op.n_out + row) * 16 + slot;
    } else if (op.kind == FSWEEP_OP_PSOS) {
      int slot = i & 15, sidx = i >> 4;
      o = ((size_t)sidx * op.n_out + row) * 16 + slot;
    } else if (diag) {
      o = row;
    } else {
      o = (size_t)row * op.n_in + i;
    }
    o += (size_t)it0 * op.grad_is;
    if (op.kind == FSWEEP_OP_DELAY || op.kind == FSWEEP_OP_PDELAY)
      reinterpret_cast<double*>(op.grad)[o] = s;
    else
      reinterpret_cast<T*>(op.grad)[o] = (T)s;
  }
}

// The default gradient finalize since round 2 (FSWEEP_FINALIZE_V2=0 selects fsweep_finalize_kernel above; parity of the
// two on every case: tests/test_gpu_random_trees.py).  Same sums as fsweep_finalize_kernel with coalesced reads: the 32 lanes of a warp own 32 CONSECUTIVE accumulators of
// one partial row (row index fastest, which is how the rows are laid out), the block's warps split the per-block rows,
// and the cross-warp sum runs in a fixed order through shared memory (deterministic).  The first version gives every
// lane of a warp a different block's row: n_blocks scattered 4-byte loads per gradient element.
constexpr int FIN2_WARPS = 32;  // 24 independent loads per lane for 751 partial rows: three batches of 8 in flight

template <typename T>
__device__ __forceinline__ void finalize_store(const FinalizeOp& op, int row, int i, double s, int item) {
  const bool diag = !(op.kind == FSWEEP_OP_GAIN || op.kind == FSWEEP_OP_SOS || op.kind == FSWEEP_OP_DELAY);
  size_t o;
  if (op.kind == FSWEEP_OP_SOS || op.kind == FSWEEP_OP_PSOS) {
    const int slot = i & 15, sn = i >> 4;  // i = (section [* n_in + n]) * 16 + slot
    o = ((size_t)sn * op.n_out + row) * 16 + slot;
  } else if (diag) {
    o = row;
  } else {
    o = (size_t)row * op.n_in + i;
  }
  o += (size_t)item * op.grad_is;
  if (op.kind == FSWEEP_OP_DELAY || op.kind == FSWEEP_OP_PDELAY)
    reinterpret_cast<double*>(op.grad)[o] = s;
  else
    reinterpret_cast<T*>(op.grad)[o] = (T)s;
}

template <typename T>
__global__ void __launch_bounds__(32 * FIN2_WARPS) fsweep_finalize_v2_kernel(const __grid_constant__ FinalizeArgs F) {
  __shared__ double red[FIN2_WARPS][33];
  const int opi = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_sync();
  if (opi == F.n_ops) {  // fused criterion: sum of the per-block squared-error sums
    if (F.loss == nullptr || blockIdx.x != 0 || blockIdx.z != 0) return;
    double s = 0.0;
    for (int b = threadIdx.x; b < F.n_blocks * F.n_items; b += 32 * FIN2_WARPS) s += F.loss_partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp][0] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < FIN2_WARPS; ++w) t += red[w][0];
      *reinterpret_cast<T*>(F.loss) = (T)(F.crit_scale * t);
    }
    return;
  }
  const FinalizeOp& op = F.ops[opi];
  if (op.grad == nullptr || (op.acc_mode != ACC_SMEM && op.acc_mode != ACC_GLOBAL)) return;  // uniform in the block
  // per-item op: this z-slice sums its own item's rows; shared op: slice 0 sums the rows of every item
  const bool own = op.grad_is != 0;
  if (!own && blockIdx.z != 0) return;
  const int row0 = own ? (int)blockIdx.z * F.n_blocks : 0, rows = own ? F.n_blocks : F.n_blocks * F.n_items;
  const int it0 = own ? (int)blockIdx.z : 0, its = own ? 1 : F.n_items;
  const int total = op.n_out * op.row_len;
  const size_t stride = (size_t)F.acc_per_lane * F.G;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {  // uniform trip count
    const int e = base + lane;  // e = i * n_out + row: consecutive lanes, consecutive rows of accumulator i
    const bool live = e < total;
    const int i = live ? e / op.n_out : 0, row = live ? e - i * op.n_out : 0;
    double s = 0.0;
    if (op.acc_mode == ACC_SMEM) {
      if (live) {
        const T* p = reinterpret_cast<const T*>(F.partial) + (size_t)(op.row_off + i) * F.G + row;
#pragma unroll 8
        for (int b = warp; b < rows; b += FIN2_WARPS) s += (double)p[(size_t)(row0 + b) * stride];
      }
      red[warp][lane] = s;
      __syncthreads();
      if (warp == 0) {
        s = 0.0;
#pragma unroll
        for (int w = 0; w < FIN2_WARPS; ++w) s += red[w][lane];
      }
      __syncthreads();  // red is reused by the next trip
    } else if (live) {
      for (int t = 0; t < its; ++t)
        s += (double)reinterpret_cast<const T*>(F.gacc)[(size_t)(it0 + t) * F.acc_total + op.acc_off + (size_t)row * op.row_len + i];
    }
    if (warp == 0 && live) finalize_store<T>(op, row, i, s, it0);
  }
}

// ------------------------------------------------------------------------------------ deferred SOS gradients
// Coefficient gradient of a section-cascade op whose accumulators do not fit shared memory (e.g. the 16 x 16 x 30
// sections of BASELINE config 3: 122 880 accumulators).  The backward kernel parked S_in and g_out per (column,
// bin); here ONE BLOCK OWNS ONE CHANNEL PAIR (m, n) and a range of bins, and works tile by tile (128 bins):
//   phase 1  thread t evaluates the whole cascade H of bin t and gh = sum_q g_out[m] conj(S_in[n]), and leaves
//            (H, gh, v, v^2) of its bin in shared memory;
//   phase 2  thread t = (section s, bin lane l) keeps the six coefficients of ITS section in registers, walks the
//            tile's bins l, l + TS, ... and adds dL/d{B(w0), B'(w0), b2, A(w0), A'(w0), a2} into six registers.
// The sums stay in registers across all tiles of the block and leave it as one atomicAdd per coefficient (the
// in-kernel path did six global atomics per section, pair and bin: 4.4e9 for config 3).  Small loop bodies: the
// first version of this kernel unrolled bins x sections into 0.8 MB of SASS and ran at instruction-fetch speed.
// Bins of one block lie in one half of the spectrum (one Taylor block); a bin whose rounded cos(w) disagrees with
// the geometric split (at most the boundary bin) takes a slow atomic path.
constexpr int DEF_BLOCK = 128;
constexpr int DEF_TILES = 8;       // tiles of DEF_BLOCK bins per block
constexpr int DEF_MAX_K = DEF_BLOCK;  // one thread per section at least

struct DeferArgs {
  const void* defer;
  void* gacc;
  long long n_bins, bin_begin, n_plus;  // n_plus: processed bins [0, n_plus) have 4k <= nfft
  int ncols_total, chunks_plus, opi;
};

// Response table of a deferred section-cascade op for the processed bins: tab[k][m][n] (PSOS: tab[k][n]), indexed by
// ABSOLUTE bin like a TABLE op (the pointer handed in is already offset by -bin_begin rows).  One block per channel
// pair, a bin per thread: the pair's coefficients are uniform loads that stay in L1, where the row-distributed
// backward kernel streamed the whole coefficient set (491 KB for config 3) from L2 once per bin.  The backward
// kernel then reads the op as a TABLE, and the gradient kernel below takes H from here instead of re-evaluating it.
template <typename T>
__global__ void __launch_bounds__(DEF_BLOCK) fsweep_sos_table_kernel(const __grid_constant__ ProgK P, const DeferArgs D) {
  const OpK& op = P.ops[D.opi];
  const bool par = op.kind == FSWEEP_OP_PSOS;
  const int pair = blockIdx.x;
  const int m = par ? pair : pair / op.n_in, n = par ? pair : pair - m * op.n_in;
  const int K = op.K;
  const T* coef = reinterpret_cast<const T*>(op.coef) + (par ? (size_t)n * 16 : ((size_t)n * op.n_out + m) * 16);
  const long stride = par ? (long)op.n_out * 16 : (long)op.n_in * op.n_out * 16;
  const size_t row = par ? (size_t)op.n_out : (size_t)op.n_out * op.n_in;
  cx<T>* tab = reinterpret_cast<cx<T>*>(op.gtab);
  const long long base = (long long)blockIdx.y * DEF_BLOCK * DEF_TILES;
  for (int tile = 0; tile < DEF_TILES; ++tile) {
    const long long bl = base + (long long)tile * DEF_BLOCK + threadIdx.x;
    if (bl >= D.n_bins) break;
    const Ctx<T> ctx = make_ctx<T>(P, D.bin_begin + bl);
    bool guarded;
    cx<T> H;
    if constexpr (sizeof(T) == 8) {
      H = K <= 64 ? sos_eval_numden(coef, K, stride, ctx, guarded) : sos_eval<T>(coef, K, stride, ctx, guarded);
    } else {
      H = sos_eval<T>(coef, K, stride, ctx, guarded);
    }
    st_cx(tab + (size_t)ctx.k * row + pair, H);
  }
}

// T: type of the buffers (coefficients, response table, parked vectors, accumulators); TC: arithmetic type.  TC = float
// with T = double serves float32 models swept in float64 arithmetic (plan flag FSWEEP_DT_GRAD32): their gradients are
// held to 1e-3 and the float64 version of this kernel is FP64-pipe bound (4.3 ms of config 3's 6.8 ms step).
template <typename T, typename TC = T>
__global__ void __launch_bounds__(DEF_BLOCK) fsweep_sos_defer_kernel(const __grid_constant__ ProgK P, const DeferArgs D) {
  constexpr int CH = DEF_BLOCK * DEF_TILES;
  __shared__ TC sH[2][DEF_BLOCK], sG[2][DEF_BLOCK], sU1[2][DEF_BLOCK], sU2[2][DEF_BLOCK];
  auto cvt = [](cx<T> v) { return mk<TC>((TC)v.x, (TC)v.y); };
  __shared__ int sState[DEF_BLOCK];
  const OpK& op = P.ops[D.opi];
  const bool par = op.kind == FSWEEP_OP_PSOS;
  const int pair = blockIdx.x;
  const int m = par ? pair : pair / op.n_in, n = par ? pair : pair - m * op.n_in;
  const bool plus = (int)blockIdx.y < D.chunks_plus;
  const long long base = plus ? (long long)blockIdx.y * CH : D.n_plus + (long long)((int)blockIdx.y - D.chunks_plus) * CH;
  const long long lim = plus ? D.n_plus : D.n_bins;
  const int tid = threadIdx.x;
  const int K = op.K;
  const T* coef = reinterpret_cast<const T*>(op.coef) + (par ? (size_t)n * 16 : ((size_t)n * op.n_out + m) * 16);
  const long stride = par ? (long)op.n_out * 16 : (long)op.n_in * op.n_out * 16;
  const int sec_stride = op.row_len / K;  // accumulator slots per section in this op's row
  T* gdst = reinterpret_cast<T*>(D.gacc) + op.acc_off + (size_t)m * op.row_len + (par ? 0 : n * 16);
  const cx<T>* defer = reinterpret_cast<const cx<T>*>(D.defer);

  // phase-2 role: section s, bin lane l of TS
  const int TS = DEF_BLOCK / K;
  const int s = tid / TS, l = tid - s * TS;
  const bool worker = s < K;
  TC c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = TC(0);
  if (worker) load8_as<TC, T>(coef + (size_t)s * stride + (plus ? 0 : 8), c);
  TC a0 = TC(0), a1 = TC(0), a2 = TC(0), a4 = TC(0), a5 = TC(0), a6 = TC(0);

  for (int tile = 0; tile < DEF_TILES; ++tile) {
    const long long bl = base + (long long)tile * DEF_BLOCK + tid;
    if (base + (long long)tile * DEF_BLOCK >= lim) break;  // uniform
    // ---- phase 1: this thread's bin
    int st = 0;
    cx<TC> H = mk<TC>(0, 0), gh = mk<TC>(0, 0), u1 = mk<TC>(0, 0), u2 = mk<TC>(0, 0);
    if (bl < lim) {
      const Ctx<T> ctx = make_ctx<T>(P, D.bin_begin + bl);
      bool guarded = false;
      cx<T> Hs;
      if (op.gtab != nullptr) {
        // the response table built for the backward kernel; (eps, 0) may be the reference's zero guard: re-check
        Hs = ld_cx(reinterpret_cast<const cx<T>*>(op.gtab) + (size_t)ctx.k * (par ? (size_t)op.n_out : (size_t)op.n_out * op.n_in) + pair);
        if (Hs.x == eps_of<T>() && Hs.y == T(0)) Hs = sos_eval<T>(coef, K, stride, ctx, guarded);
      } else {
        Hs = sos_eval<T>(coef, K, stride, ctx, guarded);
      }
      H = cvt(Hs);
      u1 = cvt(ctx.u1);
      u2 = cvt(ctx.u2);
      for (int q = 0; q < D.ncols_total; ++q) {
        const cx<T>* rec = defer + ((size_t)q * D.n_bins + bl) * P.def_stride + op.def_off;
        cfmac(gh, cvt(ld_cx(rec + op.n_in + m)), cvt(ld_cx(rec + n)));  // g_out[m] conj(S_in[n])
      }
      st = guarded ? 0 : (ctx.plus == plus ? 1 : 2);
    }
    if (st == 2) {
      // the boundary bin (rounded cos(w) on the other side of the split): its own thread, straight atomics
      const Ctx<T> ctx = make_ctx<T>(P, D.bin_begin + bl);
      const T* p = coef + (ctx.plus ? 0 : 8);
      const cx<T> Hs = mk<T>((T)H.x, (T)H.y), ghs = mk<T>((T)gh.x, (T)gh.y);
      for (int ss = 0; ss < K; ++ss) {
        T cc[8];
        load8<T>(p + (size_t)ss * stride, cc);
        cx<T> Bv, Av;
        section_eval<T>(cc, ctx, Bv, Av);
        const cx<T> qb = czero(Bv) ? cmul(sos_eval_without<T>(coef, K, stride, ctx, ss), crcp_exact(Av)) : cmul(Hs, crcp_exact(Bv));
        cx<T> qa = cmul(Hs, crcp_exact(Av));
        const cx<T> rb = cmulc(ghs, qb), ra = cmulc(ghs, mk<T>(-qa.x, -qa.y));
        T* g = gdst + (size_t)ss * sec_stride + (ctx.plus ? 0 : 8);
        atomicAdd(g + 0, rb.x);
        atomicAdd(g + 1, rb.x * ctx.u1.x + rb.y * ctx.u1.y);
        atomicAdd(g + 2, rb.x * ctx.u2.x + rb.y * ctx.u2.y);
        atomicAdd(g + 4, ra.x);
        atomicAdd(g + 5, ra.x * ctx.u1.x + ra.y * ctx.u1.y);
        atomicAdd(g + 6, ra.x * ctx.u2.x + ra.y * ctx.u2.y);
      }
      st = 0;
    }
    sH[0][tid] = H.x;
    sH[1][tid] = H.y;
    sG[0][tid] = gh.x;
    sG[1][tid] = gh.y;
    sU1[0][tid] = u1.x;
    sU1[1][tid] = u1.y;
    sU2[0][tid] = u2.x;
    sU2[1][tid] = u2.y;
    sState[tid] = st;
    __syncthreads();
    // ---- phase 2: my section over the tile's bins
    if (worker) {
      for (int b = l; b < DEF_BLOCK; b += TS) {
        if (sState[b] == 0) continue;
        Ctx<TC> cx_;
        cx_.u1 = mk<TC>(sU1[0][b], sU1[1][b]);
        cx_.u2 = mk<TC>(sU2[0][b], sU2[1][b]);
        const cx<TC> Hb = mk<TC>(sH[0][b], sH[1][b]), gb = mk<TC>(sG[0][b], sG[1][b]);
        cx<TC> Bv, Av;
        section_eval<TC>(c, cx_, Bv, Av);
        cx<TC> qb;
        if (czero(Bv)) {  // rare: this numerator section vanishes at this bin (H itself is 0 there)
          const Ctx<T> full = make_ctx<T>(P, D.bin_begin + base + (long long)tile * DEF_BLOCK + b);
          qb = cmul(cvt(sos_eval_without<T>(coef, K, stride, full, s)), crcp_exact(Av));
        } else {
          qb = cmul(Hb, crcp(Bv));
        }
        const cx<TC> qa = cmul(Hb, crcp(Av));
        const cx<TC> rb = cmulc(gb, qb), ra = cmulc(gb, mk<TC>(-qa.x, -qa.y));
        a0 += rb.x;
        a1 = fma(rb.x, cx_.u1.x, fma(rb.y, cx_.u1.y, a1));
        a2 = fma(rb.x, cx_.u2.x, fma(rb.y, cx_.u2.y, a2));
        a4 += ra.x;
        a5 = fma(ra.x, cx_.u1.x, fma(ra.y, cx_.u1.y, a5));
        a6 = fma(ra.x, cx_.u2.x, fma(ra.y, cx_.u2.y, a6));
      }
    }
    __syncthreads();
  }
  if (worker) {
    T* g = gdst + (size_t)s * sec_stride + (plus ? 0 : 8);
    atomicAdd(g + 0, (T)a0);
    atomicAdd(g + 1, (T)a1);
    atomicAdd(g + 2, (T)a2);
    atomicAdd(g + 4, (T)a4);
    atomicAdd(g + 5, (T)a5);
    atomicAdd(g + 6, (T)a6);
  }
}

// launch thunks, one translation unit per G (fsweep_inst.cu compiled with -DFSWEEP_G=...)
struct LaunchCfg {
  int grid;
  size_t smem;
  cudaStream_t stream;
  int items = 1;  // per-item launches of the generic kernels: grid.y
};
template <int G>
cudaError_t launch_fwd(int dtype, int cc, const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A);
template <int G>
cudaError_t launch_bwd(int dtype, int cc, const LaunchCfg& cfg, const ProgK& P, const SweepArgs& A);
template <int G>
cudaError_t occupancy(int dtype, int cc, bool bwd, size_t smem, int* blocks_per_sm);

}  // namespace fsweep
