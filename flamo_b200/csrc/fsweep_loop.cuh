// fsweep_loop.cuh — pattern-specialised sweep for the FDN structure, flamo's flagship system
// (examples/e8_colorless_fdn.py, auxiliary/reverb.py HomogeneousFDN):
//
//     [GAIN pre]  ->  RECURSION( F = diagonal ops (delays, attenuation filters, gains),
//                                Fb = ONE constant real matrix (dsp.Matrix / dsp.Gain) )  ->  [GAIN post]
//
// Same math, mapping and data layout as the generic kernels in fsweep_kernels.cuh (one group of G
// lanes per bin, lane r = row r), but the program structure is known at compile time, so
//   * row `lane` of the feedback matrix and its gradient accumulator live in REGISTERS for the whole
//     kernel (the generic interpreter re-reads them from shared memory per bin and accumulates
//     gradients with shared-memory read-modify-writes),
//   * A = I - D(w) W is formed with 2 FMAs per entry (no matrix-valued op application),
//   * the single-output-channel gain (1 x N) is a dot product done as one butterfly instead of N
//     broadcasts, and the single-input-channel gain (N x 1) needs no shuffle at all,
//   * there is no step-table interpretation in the per-bin path.
// The diagonal chain is still evaluated by the generic stage_op / backprop_op (any mix of PDELAY,
// PGAIN, PSOS, PTABLE), so attenuation filters inside the loop are covered.
#pragma once
#include "fsweep_kernels.cuh"

namespace fsweep {

// indices into ProgK::ops of the pattern's pieces (host-checked in fsweep_plan_create)
struct LoopInfo {
  int pre;       // GAIN before the recursion, or -1
  int ff_begin;  // first feedforward (diagonal) op
  int n_ff;
  int fb;        // the feedback GAIN
  int post;      // GAIN after the recursion, or -1
  int pad_[3];
};

template <typename T, int G>
struct LoopRegs {
  T wfb[G];  // row `lane` of the feedback matrix (zero beyond the live width)
};

template <typename T, int G>
__device__ __forceinline__ void load_fb_row(const OpK& op, int lane, T (&w)[G]) {
  static_for<0, G>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    w[j] = (lane < op.n_out && j < op.n_in) ? __ldg(reinterpret_cast<const T*>(op.coef) + lane * op.n_in + j) : T(0);
  });
}

// D_lane = product of the diagonal chain's responses at this bin (staged in hc by the caller)
template <typename T>
__device__ __forceinline__ cx<T> chain_product(const ProgK& P, const LoopInfo& L, const cx<T>* hc, int tid, int skip) {
  cx<T> D = mk<T>(1, 0);
  for (int i = 0; i < L.n_ff; ++i) {
    if (i == skip) continue;
    D = cmul(D, hc[(size_t)P.ops[L.ff_begin + i].h_off * BLOCK + tid]);
  }
  return D;
}

template <typename T, int G>
__device__ __forceinline__ void build_fdn(const T (&wfb)[G], cx<T> D, int lane, LU<T, G>& lu) {
  static_for<0, G>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    lu.a[j] = mk<T>((j == lane ? T(1) : T(0)) - D.x * wfb[j], -D.y * wfb[j]);
  });
  lu.factor(lane);
}

// recursion input s = W_pre x (or x itself), row-distributed
template <typename T, int G>
__device__ __forceinline__ cx<T> loop_input(const ProgK& P, const LoopInfo& L, const cx<T>* x, long long xbs,
                                            long long bl, int b, int cc, int cols, int lane, const cx<T>* hc, int tid,
                                            cx<T>& xin) {
  xin = mk<T>(0, 0);
  if (L.pre < 0) {
    if (lane < P.in_ch) xin = ld_cx(x + (size_t)b * xbs + ((size_t)bl * P.in_ch + lane) * cols + cc);
    return xin;
  }
  const OpK& op = P.ops[L.pre];
  const cx<T>* hrow = hc + (size_t)op.h_off * BLOCK + tid;
  if (op.n_in == 1) {  // every lane reads the one input channel: no shuffle
    xin = ld_cx(x + (size_t)b * xbs + (size_t)bl * cols + cc);
    const T w = hrow[0].x;
    return mk<T>(w * xin.x, w * xin.y);
  }
  if (lane < op.n_in) xin = ld_cx(x + (size_t)b * xbs + ((size_t)bl * op.n_in + lane) * cols + cc);
  cx<T> s = mk<T>(0, 0);
  for (int n = 0; n < op.n_in; ++n) {
    const cx<T> v = shfl<G>(xin, n);
    const T w = hrow[(size_t)n * BLOCK].x;
    s.x = fma(w, v.x, s.x);
    s.y = fma(w, v.y, s.y);
  }
  return s;
}

// output o = W_post y (or y itself), row-distributed; wpc = W_post[0][lane] when n_out == 1
template <typename T, int G>
__device__ __forceinline__ cx<T> loop_output(const ProgK& P, const LoopInfo& L, cx<T> y, int lane, T wpc,
                                             const cx<T>* hc, int tid) {
  if (L.post < 0) return y;
  const OpK& op = P.ops[L.post];
  if (op.n_out == 1) return group_sum<G>(mk<T>(wpc * y.x, wpc * y.y));  // all lanes hold the sum
  const cx<T>* hrow = hc + (size_t)op.h_off * BLOCK + tid;
  cx<T> o = mk<T>(0, 0);
  for (int n = 0; n < op.n_in; ++n) {
    const cx<T> v = shfl<G>(y, n);
    const T w = hrow[(size_t)n * BLOCK].x;
    o.x = fma(w, v.x, o.x);
    o.y = fma(w, v.y, o.y);
  }
  return o;
}

// ------------------------------------------------------------------------------------ forward
template <typename T, int G>
__global__ void __launch_bounds__(BLOCK) fsweep_loop_fwd_kernel(const __grid_constant__ ProgK P,
                                                               const __grid_constant__ LoopInfo L, const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* hc = reinterpret_cast<cx<T>*>(smem_raw);
  unsigned* gmask = reinterpret_cast<unsigned*>(hc + (size_t)P.h_total * BLOCK);
  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const long long groups_total = (long long)gridDim.x * (BLOCK / G);
  const long long gg = (long long)blockIdx.x * (BLOCK / G) + tid / G;
  const long long n_iter = (A.n_bins + groups_total - 1) / groups_total;
  const int ncols_total = A.batch * A.cols;
  const cx<T>* x = reinterpret_cast<const cx<T>*>(A.x);

  T wfb[G];
  load_fb_row<T, G>(P.ops[L.fb], lane, wfb);
  T wpc = T(0);
  {
    const Ctx<T> c0 = make_ctx<T>(P, A.bin_begin);
    stage_ops<T>(P, c0, lane, hc, gmask, tid, true, L.fb);  // input/output gains once per kernel (W_fb is in registers)
    if (L.post >= 0 && P.ops[L.post].n_out == 1 && lane < P.ops[L.post].n_in)
      wpc = __ldg(reinterpret_cast<const T*>(P.ops[L.post].coef) + lane);
  }
  const int out_rows = P.out_ch;
  double lacc = 0.0;

  for (long long it = 0; it < n_iter; ++it) {
    long long bl = it * groups_total + gg;
    const bool valid = bl < A.n_bins;
    if (!valid) bl = A.n_bins - 1;
    const Ctx<T> ctx = make_ctx<T>(P, A.bin_begin + bl);
    for (int i = 0; i < L.n_ff; ++i) stage_op<T>(P.ops[L.ff_begin + i], ctx, lane, hc, gmask, L.ff_begin + i, tid);
    const cx<T> D = chain_product<T>(P, L, hc, tid, -1);
    LU<T, G> lu;
    build_fdn<T, G>(wfb, D, lane, lu);

    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      cx<T> xin;
      cx<T> s = loop_input<T, G>(P, L, x, A.xbs, bl, b, cc, A.cols, lane, hc, tid, xin);
      cx<T> y[1] = {cmul(D, s)};
      lu.template solve<1>(lane, y);
      const cx<T> o = loop_output<T, G>(P, L, y[0], lane, wpc, hc, tid);
      if (epi_fused(A.epilogue)) {
        crit_rowdist<T, G>(A, abs_t(o.x, o.y), lane < out_rows, lane, out_rows, bl, b, lane, valid, lacc);
      } else if (valid && lane < out_rows) {
        size_t off = (size_t)b * A.ybs + ((size_t)bl * out_rows + lane) * A.cols + cc;
        if (A.epilogue == FSWEEP_EPI_ABS)
          reinterpret_cast<T*>(A.y)[off] = abs_t(o.x, o.y);
        else
          st_cx(reinterpret_cast<cx<T>*>(A.y) + off, o);
      }
    }
  }
  if (epi_fused(A.epilogue)) block_loss_store<T>(lacc, A.loss_partial);
}

// ------------------------------------------------------------------------------------ backward
template <typename T, int G>
__global__ void __launch_bounds__(BLOCK) fsweep_loop_bwd_kernel(const __grid_constant__ ProgK P,
                                                               const __grid_constant__ LoopInfo L, const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* hc = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* save = hc + (size_t)P.h_total * BLOCK;                          // unused here (layout shared with generic)
  T* sacc = reinterpret_cast<T*>(save + (size_t)P.n_slots * 1 * BLOCK);  // [acc_per_lane][BLOCK]
  unsigned* gmask = reinterpret_cast<unsigned*>(sacc + (size_t)P.acc_per_lane * BLOCK);

  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const long long groups_total = (long long)gridDim.x * (BLOCK / G);
  const long long gg = (long long)blockIdx.x * (BLOCK / G) + tid / G;
  const long long n_iter = (A.n_bins + groups_total - 1) / groups_total;
  const int ncols_total = A.batch * A.cols;
  const cx<T>* x = reinterpret_cast<const cx<T>*>(A.x);

  for (int i = 0; i < P.acc_per_lane; ++i) sacc[(size_t)i * BLOCK + tid] = T(0);
  Acc<T> acc;
  acc.sacc = sacc;
  acc.gacc = reinterpret_cast<T*>(A.gacc);
  acc.tid = tid;

  const OpK& fbop = P.ops[L.fb];
  T wfb[G], gwfb[G];
  load_fb_row<T, G>(fbop, lane, wfb);
  static_for<0, G>([&](auto jc) { gwfb[decltype(jc)::value] = T(0); });
  T wpc = T(0), gwpc = T(0), gwp0 = T(0);
  const bool post1 = L.post >= 0 && P.ops[L.post].n_out == 1;
  const bool pre1 = L.pre >= 0 && P.ops[L.pre].n_in == 1;
  {
    const Ctx<T> c0 = make_ctx<T>(P, A.bin_begin);
    stage_ops<T>(P, c0, lane, hc, gmask, tid, true, L.fb);
    if (post1 && lane < P.ops[L.post].n_in) wpc = __ldg(reinterpret_cast<const T*>(P.ops[L.post].coef) + lane);
  }
  const int out_rows = P.out_ch;
  const bool want_fb = fbop.acc_mode == ACC_SMEM;
  const bool want_pre = L.pre >= 0 && P.ops[L.pre].acc_mode == ACC_SMEM;
  const bool want_post = L.post >= 0 && P.ops[L.post].acc_mode == ACC_SMEM;
  double lacc = 0.0;

  for (long long it = 0; it < n_iter; ++it) {
    long long bl = it * groups_total + gg;
    const bool valid = bl < A.n_bins;
    if (!valid) bl = A.n_bins - 1;
    acc.valid = valid;
    const T vmask = valid ? T(1) : T(0);
    const Ctx<T> ctx = make_ctx<T>(P, A.bin_begin + bl);
    for (int i = 0; i < L.n_ff; ++i) stage_op<T>(P.ops[L.ff_begin + i], ctx, lane, hc, gmask, L.ff_begin + i, tid);
    const cx<T> D = chain_product<T>(P, L, hc, tid, -1);
    LU<T, G> lu;
    build_fdn<T, G>(wfb, D, lane, lu);

    for (int q = 0; q < ncols_total; ++q) {
      const int b = (A.cols == 1) ? q : q / A.cols, cc = q - b * A.cols;
      const bool first_chunk = q == 0;
      // ---- forward recompute
      cx<T> xin;
      const cx<T> s = loop_input<T, G>(P, L, x, A.xbs, bl, b, cc, A.cols, lane, hc, tid, xin);
      cx<T> yv[1] = {cmul(D, s)};
      lu.template solve<1>(lane, yv);
      const cx<T> y = yv[0];
      const cx<T> o = loop_output<T, G>(P, L, y, lane, wpc, hc, tid);
      // ---- output gradient (row-distributed over out_rows; replicated on all lanes when post1)
      cx<T> go = mk<T>(0, 0);
      if (epi_fused(A.epilogue)) {
        // post1: o (row 0) is replicated on every lane of the group; only lane 0 counts the error
        const T mag = abs_t(o.x, o.y);
        const T ga = crit_rowdist<T, G>(A, mag, post1 ? lane == 0 : lane < out_rows, post1 ? 0 : lane, out_rows, bl, b,
                                        lane, valid, lacc);
        const T gab = post1 ? __shfl_sync(FULL, ga, 0, G > 32 ? 32 : G) : ga;
        if ((lane < out_rows || post1) && mag > T(0)) {
          const T r = gab * rcp_t(mag);
          go = mk<T>(r * o.x, r * o.y);
        }
      } else if (lane < out_rows || post1) {
        const int row = post1 ? 0 : lane;
        size_t off = (size_t)b * A.gybs + ((size_t)bl * out_rows + row) * A.cols + cc;
        if (A.epilogue == FSWEEP_EPI_ABS) {
          T ga = __ldg(reinterpret_cast<const T*>(A.gy) + off);
          T mag = abs_t(o.x, o.y);
          if (mag > T(0)) {
            T r = ga * rcp_t(mag);
            go = mk<T>(r * o.x, r * o.y);
          }
        } else {
          go = ld_cx(reinterpret_cast<const cx<T>*>(A.gy) + off);
        }
      }
      // ---- through the output gain: g_y = W_post^T g_o, dW_post = Re(g_o y^H)
      cx<T> gy[1];
      if (L.post < 0) {
        gy[0] = go;
      } else if (post1) {
        gy[0] = mk<T>(wpc * go.x, wpc * go.y);
        gwpc += vmask * (go.x * y.x + go.y * y.y);
      } else {
        const OpK& op = P.ops[L.post];
        const cx<T>* hrow = hc + (size_t)op.h_off * BLOCK + tid;
        cx<T> gin = mk<T>(0, 0);
        for (int n = 0; n < op.n_in; ++n) {
          const cx<T> yn = shfl<G>(y, n);
          const T w = hrow[(size_t)n * BLOCK].x;
          if (want_post && lane < op.n_out) acc.add(op, lane, n, go.x * yn.x + go.y * yn.y);
          const cx<T> t = group_sum<G>(mk<T>(w * go.x, w * go.y));
          gin.x = (lane == n) ? t.x : gin.x;
          gin.y = (lane == n) ? t.y : gin.y;
        }
        gy[0] = gin;
      }
      // ---- adjoint solve and the loop: lambda = A^-H g_y, g_u = conj(D) lambda, u = s + W y
      lu.template solve_adj<1>(lane, gy);
      const cx<T> lam = gy[0];
      const cx<T> gu = mk<T>(D.x * lam.x + D.y * lam.y, D.x * lam.y - D.y * lam.x);  // conj(D) * lam
      cx<T> u = s;
      const T gux = vmask * gu.x, guy = vmask * gu.y;
      static_for<0, G>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const cx<T> yj = shfl<G>(y, j);
        u.x = fma(wfb[j], yj.x, u.x);
        u.y = fma(wfb[j], yj.y, u.y);
        gwfb[j] = fma(gux, yj.x, fma(guy, yj.y, gwfb[j]));  // Re(g_u conj(y_j))
      });
      // ---- diagonal chain: op i sees input u * prod_{j<i} d_j and output gradient lam * conj(prod_{j>i} d_j)
      if (L.n_ff == 1) {
        cx<T> Sin[1] = {u}, g1[1] = {lam};
        backprop_op<T, G, 1>(P.ops[L.ff_begin], ctx, lane, Sin, g1, false, acc, first_chunk, hc,
                             gmask[L.ff_begin * BLOCK + tid], tid);
      } else {
        cx<T> pre = mk<T>(1, 0);
        for (int i = 0; i < L.n_ff; ++i) {
          cx<T> suf = mk<T>(1, 0);
          for (int j = i + 1; j < L.n_ff; ++j) suf = cmul(suf, hc[(size_t)P.ops[L.ff_begin + j].h_off * BLOCK + tid]);
          cx<T> Sin[1] = {cmul(u, pre)};
          cx<T> g1[1] = {cmul(lam, mk<T>(suf.x, -suf.y))};
          backprop_op<T, G, 1>(P.ops[L.ff_begin + i], ctx, lane, Sin, g1, false, acc, first_chunk, hc,
                               gmask[(L.ff_begin + i) * BLOCK + tid], tid);
          pre = cmul(pre, hc[(size_t)P.ops[L.ff_begin + i].h_off * BLOCK + tid]);
        }
      }
      // ---- through the input gain: dW_pre = Re(g_s x^H), g_x = W_pre^T g_s   (g_s = g_u)
      if (L.pre < 0) {
        if (A.gx != nullptr && valid && lane < P.in_ch)
          st_cx(reinterpret_cast<cx<T>*>(A.gx) + (size_t)b * A.gxbs + ((size_t)bl * P.in_ch + lane) * A.cols + cc, gu);
      } else if (pre1) {
        gwp0 += gux * xin.x + guy * xin.y;
        if (A.gx != nullptr) {
          const T w = hc[(size_t)P.ops[L.pre].h_off * BLOCK + tid].x;
          const cx<T> t = group_sum<G>(mk<T>(w * gu.x, w * gu.y));
          if (valid && lane == 0) st_cx(reinterpret_cast<cx<T>*>(A.gx) + (size_t)b * A.gxbs + (size_t)bl * A.cols + cc, t);
        }
      } else {
        const OpK& op = P.ops[L.pre];
        const cx<T>* hrow = hc + (size_t)op.h_off * BLOCK + tid;
        cx<T> gxl = mk<T>(0, 0);
        for (int n = 0; n < op.n_in; ++n) {
          const cx<T> xn = shfl<G>(xin, n);
          if (want_pre && lane < op.n_out) acc.add(op, lane, n, gu.x * xn.x + gu.y * xn.y);
          if (A.gx != nullptr) {
            const T w = hrow[(size_t)n * BLOCK].x;
            const cx<T> t = group_sum<G>(mk<T>(w * gu.x, w * gu.y));
            gxl.x = (lane == n) ? t.x : gxl.x;
            gxl.y = (lane == n) ? t.y : gxl.y;
          }
        }
        if (A.gx != nullptr && valid && lane < op.n_in)
          st_cx(reinterpret_cast<cx<T>*>(A.gx) + (size_t)b * A.gxbs + ((size_t)bl * op.n_in + lane) * A.cols + cc, gxl);
      }
    }
  }

  // ---- flush the register accumulators into the thread-private shared-memory columns
  if (want_fb && lane < fbop.n_out) {
    static_for<0, G>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      if (j < fbop.n_in) sacc[(size_t)(fbop.row_off + j) * BLOCK + tid] += gwfb[j];
    });
  }
  if (pre1 && want_pre && lane < P.ops[L.pre].n_out) sacc[(size_t)P.ops[L.pre].row_off * BLOCK + tid] += gwp0;
  __syncthreads();
  if (post1 && want_post && lane < P.ops[L.post].n_in)  // entry `lane` of row 0 lives in the column of the group's lane 0
    sacc[(size_t)(P.ops[L.post].row_off + lane) * BLOCK + (tid - lane)] += gwpc;

  // ---- block reduction of the thread-private accumulator columns: partial[block][i][row]
  __syncthreads();
  T* partial = reinterpret_cast<T*>(A.partial) + (size_t)blockIdx.x * P.acc_per_lane * G;
  for (int e = tid; e < P.acc_per_lane * G; e += BLOCK) {
    int i = e / G, row = e - i * G;
    T s = T(0);
    for (int j = 0; j < BLOCK / G; ++j) s += sacc[(size_t)i * BLOCK + j * G + row];
    partial[e] = s;
  }
  if (epi_fused(A.epilogue)) block_loss_store<T>(lacc, A.loss_partial);
}

template <int G>
cudaError_t launch_loop_fwd(int dtype, const LaunchCfg& cfg, const ProgK& P, const LoopInfo& L, const SweepArgs& A);
template <int G>
cudaError_t launch_loop_bwd(int dtype, const LaunchCfg& cfg, const ProgK& P, const LoopInfo& L, const SweepArgs& A);
template <int G>
cudaError_t occupancy_loop(int dtype, bool bwd, size_t smem, int* blocks_per_sm);

}  // namespace fsweep
