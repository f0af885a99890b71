// fsweep_params.cu — the PARAMETER-SIZED kernels of a training step: everything here works on O(#params) data and
// costs one launch, which is what matters inside a captured step (a step of the headline config is 17 launches).
//
//  * E = exp(S), S = triu(P,1) - triu(P,1)^T (or S = P), and its adjoint, for the orthogonal map of dsp.Matrix
//    (reference flamo/processor/dsp.py:649, functional.py:42-56).  torch.matrix_exp copies the matrix norm to the
//    host to pick its Pade degree, which synchronises and cannot be captured in a CUDA graph.  This kernel does the
//    whole thing on the device: one CTA, float64, scaling-and-squaring with a degree-12 Taylor polynomial
//    (Paterson-Stockmeyer, 7 products) on X = S / 2^s with ||X||_1 <= 1/4 (remainder 0.25^13/13! ~ 2e-18), matrices
//    resident in shared memory.  The adjoint uses the block-triangular identity
//    exp([[S^T, G], [0, S^T]]) = [[E^T, dS], [0, E^T]] (the same Frechet-derivative formula torch.autograd uses),
//    then folds dS through the skew map.
//  * sparsity_loss of the mapped feedback matrix (reference flamo/optimize/loss.py:36-63), forward and backward.
//  * the weighted total of the step's criteria (reference flamo/optimize/trainer.py:184-188).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/fsweep.h"
#include "fsweep_pdl.cuh"

using fsweep::launch_pdl;
using fsweep::pdl_sync;

namespace {

struct TotalArgs {
  const void* part[FSWEEP_MAX_CRITERIA];
  double alpha[FSWEEP_MAX_CRITERIA];
  double scale[FSWEEP_MAX_CRITERIA];
  int n;
};
// host_vals / host_seq (optional): the same values written straight into MAPPED PINNED host memory together with the
// number of this launch: the host polls the counter and has the step's losses the moment they exist, without a copy
// node in the graph and without waiting for the rest of the step.  float32: every value travels WITH the launch number
// in one aligned 8-byte store {value, seq} — single-copy atomic, so no fence is needed and the host simply waits until
// every pair carries the number it expects; float64: plain values, a system-wide fence, then the counter.
template <typename T>
__device__ __forceinline__ void total_job(const TotalArgs& a, T* vals, void* host_vals, volatile int* host_seq,
                                          int* seq_counter) {
  int s = 0;
  if (host_vals) {
    s = *seq_counter + 1;
    *seq_counter = s;
  }
  T tot = T(0);
  for (int i = 0; i <= a.n; ++i) {
    T v = tot;
    if (i < a.n) {
      v = (T)a.scale[i] * *reinterpret_cast<const T*>(a.part[i]);
      tot += (T)a.alpha[i] * v;
    }
    vals[i] = v;
    if (host_vals) {
      if (sizeof(T) == 4) {
        const unsigned long long pair = (unsigned long long)__float_as_uint((float)v) | ((unsigned long long)(unsigned)s << 32);
        reinterpret_cast<volatile unsigned long long*>(host_vals)[i] = pair;
      } else {
        reinterpret_cast<volatile double*>(host_vals)[i] = (double)v;
      }
    }
  }
  if (host_vals) {
    if (sizeof(T) != 4) __threadfence_system();
    *host_seq = s;
    // push the stores out NOW: as a rider this thread shares a kernel with blocks that run for microseconds more,
    // and without a fence its posted writes were seen by the host only around the end of the kernel
    __threadfence_system();
  }
}


// A RIDER: one extra block of a parameter-sized kernel evaluates the weighted total of the step's criteria and
// notifies the host (fsweep_weighted_total_notify) — inside a captured step the values are only read by the host, so they
// need no launch of their own on the critical path.
struct RiderArgs {
  TotalArgs total;
  void *vals, *host_vals, *host_seq, *seq_counter;
  int on;
};
template <typename T>
__device__ __forceinline__ void run_rider(const RiderArgs& rd) {
  if (threadIdx.x == 0)
    total_job<T>(rd.total, reinterpret_cast<T*>(rd.vals), rd.host_vals, reinterpret_cast<volatile int*>(rd.host_seq),
                 reinterpret_cast<int*>(rd.seq_counter));
}
bool fill_rider(const fsweep_total_job_t* total, RiderArgs* rd) {
  rd->on = 0;
  rd->vals = rd->host_vals = rd->host_seq = rd->seq_counter = nullptr;
  rd->total.n = 0;
  if (!total) return true;
  if (total->n < 1 || total->n > FSWEEP_MAX_CRITERIA || !total->vals) return false;
  if (total->host_vals && (!total->host_seq || !total->seq_counter)) return false;
  rd->total.n = total->n;
  for (int i = 0; i < total->n; ++i) {
    if (!total->parts[i]) return false;
    rd->total.part[i] = total->parts[i];
    rd->total.alpha[i] = total->alphas[i];
    rd->total.scale[i] = total->scales[i];
  }
  rd->vals = total->vals;
  rd->host_vals = total->host_vals;
  rd->host_seq = total->host_seq;
  rd->seq_counter = total->seq_counter;
  rd->on = 1;
  return true;
}

constexpr int EXPM_THREADS = 256;

// C = A * B (n x n, shared memory, row-major).  One element per thread-iteration; (r, c) advance without divisions.
__device__ __forceinline__ void matmul(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                       int n) {
  const int nn = n * n;
  int r = threadIdx.x / n, c = threadIdx.x - r * n;
  const int dr = blockDim.x / n, dc = blockDim.x - dr * n;
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const double* a = A + r * n;
    const double* b = B + c;
    double s0 = 0.0, s1 = 0.0;
    int k = 0;
    for (; k + 1 < n; k += 2) {
      s0 = fma(a[k], b[k * n], s0);
      s1 = fma(a[k + 1], b[(k + 1) * n], s1);
    }
    if (k < n) s0 = fma(a[k], b[k * n], s0);
    C[e] = s0 + s1;
    r += dr;
    c += dc;
    if (c >= n) {
      c -= n;
      ++r;
    }
  }
  __syncthreads();
}

// out = c0 I + c1 X + c2 X2 + c3 X3 + c4 X4 + c5 X5   (elementwise combination of precomputed powers)
__device__ __forceinline__ void combine6(double* out, const double* X, const double* X2, const double* X3,
                                         const double* X4, const double* X5, const double* c, int n) {
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, col = e - r * n;
    out[e] = ((r == col) ? c[0] : 0.0) + c[1] * X[e] + c[2] * X2[e] + c[3] * X3[e] + c[4] * X4[e] + c[5] * X5[e];
  }
}

// exp(X) for X (n x n) in shared memory.  Scaling and squaring with a degree-12 Taylor polynomial evaluated
// Paterson-Stockmeyer style: with X2..X6 (5 products), p = B0 + X6 (B1 + X6 / 12!) (2 products), B0/B1 = degree-5
// blocks.  ||X/2^s||_1 <= theta: the remainder is bounded by theta^13/13! (theta = 1/2: 2e-14, float64 results; theta = 1:
// 1.6e-10, float32 results).  For NORMAL matrices (the skew-symmetric argument of the orthogonal map) the bound is very
// pessimistic — numpy, 8 x 8 and 16 x 16 skew matrices with N(0, 1) .. N(0, 9) entries: 1e-14 at theta = 1/2, 3e-13 at 1,
// 1.5e-9 at 2 — so float32 results of the skew map use theta = 2.  Every unit of log2(theta) saves one squaring, i.e. one
// of ~13 dependent products of this latency-bound kernel.  Scratch: 7 n x n matrices; result in W[0].
__device__ double* expm_inplace(double* X, double* W, int n, double* red, double theta) {
  const int nn = n * n;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += fabs(X[r * n + c]);
    red[c] = s;
  }
  __syncthreads();
  double m = 0.0;
  for (int c = 0; c < n; ++c) m = fmax(m, red[c]);  // every thread: n <= 96 broadcast reads
  int s = 0;
  if (m > theta) {
    int ex;
    frexp(m / theta, &ex);  // m / theta = f * 2^ex, f in [0.5, 1)  ->  m / 2^ex <= theta
    s = ex;
  }
  if (s > 60) s = 60;
  const double scale = ldexp(1.0, -s);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) X[e] *= scale;
  __syncthreads();
  double *X2 = W, *X3 = W + nn, *X4 = W + 2 * nn, *X5 = W + 3 * nn, *X6 = W + 4 * nn, *T0 = W + 5 * nn, *T1 = W + 6 * nn;
  matmul(X, X, X2, n);
  matmul(X2, X, X3, n);
  matmul(X2, X2, X4, n);
  matmul(X3, X2, X5, n);
  matmul(X3, X3, X6, n);
  const double c0[6] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120};
  const double c1[6] = {1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800};
  const double c12 = 1.0 / 479001600;
  // T0 = B1 + c12 X6
  combine6(T0, X, X2, X3, X4, X5, c1, n);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) T0[e] += c12 * X6[e];
  __syncthreads();
  matmul(X6, T0, T1, n);  // X6 (B1 + c12 X6)
  combine6(T0, X, X2, X3, X4, X5, c0, n);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) T0[e] += T1[e];
  __syncthreads();
  double* cur = T0;
  double* other = T1;
  for (int i = 0; i < s; ++i) {
    matmul(cur, cur, other, n);
    double* t = cur;
    cur = other;
    other = t;
  }
  return cur;
}

// IO type T (float | double) is the caller's parameter dtype; the arithmetic is float64 throughout.
// sp (optional): sparsity_loss of the result (reference optimize/loss.py:36-63, one matrix):
//   (sum |E| - n sqrt n) / (n (1 - sqrt n)),  evaluated on the values as stored (rounded to T)
template <typename T>
__global__ void __launch_bounds__(EXPM_THREADS) expm_fwd_kernel(const T* __restrict__ Pin, T* __restrict__ E,
                                                               int n, int skew, T* __restrict__ sp) {
  extern __shared__ double sm[];
  double *X = sm, *W = sm + n * n, *red = sm + 8 * n * n;
  pdl_sync();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    double v = (double)Pin[e];
    if (skew) v = (c > r) ? (double)Pin[r * n + c] : ((c < r) ? -(double)Pin[c * n + r] : 0.0);
    X[e] = v;
  }
  __syncthreads();
  double* R = expm_inplace(X, W, n, red, sizeof(T) == 4 ? (skew ? 2.0 : 1.0) : 0.5);
  double asum = 0.0;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const T v = (T)R[e];
    E[e] = v;
    asum += fabs((double)v);
  }
  if (sp != nullptr) {  // (uniform)
    __shared__ double sred[EXPM_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = asum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < EXPM_THREADS / 32; ++w) t += sred[w];
      const double rn = sqrt((double)n);
      *sp = (T)((t - n * rn) / (n * (1.0 - rn)));
    }
  }
}

// G (optional): dL/dE;  Esp + gsp (optional): the sparsity_loss term above contributes gsp * sign(E) / (n (1 - sqrt n))
template <typename T>
__global__ void __launch_bounds__(EXPM_THREADS) expm_bwd_kernel(const T* __restrict__ Pin, const T* __restrict__ G,
                                                               T* __restrict__ gP, int n, int skew,
                                                               const T* __restrict__ Esp, const T* __restrict__ gsp,
                                                               const __grid_constant__ RiderArgs rd) {
  extern __shared__ double sm[];
  const int m = 2 * n;
  double *X = sm, *W = sm + m * m, *red = sm + 8 * m * m;
  pdl_sync();
  if (blockIdx.x == 1) {
    run_rider<T>(rd);
    return;
  }
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    int r = e / m, c = e - r * m;
    double v = 0.0;
    if (r < n && c >= n) {
      if (G != nullptr) v = (double)G[r * n + (c - n)];
      if (gsp != nullptr) {
        const T a = Esp[r * n + (c - n)];
        const double rn = sqrt((double)n);
        const double k = (double)*gsp / (n * (1.0 - rn));
        v += a > T(0) ? k : (a < T(0) ? -k : 0.0);
      }
    } else if ((r < n) == (c < n)) {
      int i = r % n, j = c % n;  // block (i, j) of S^T = S[j][i]
      if (skew)
        v = (i > j) ? (double)Pin[j * n + i] : ((i < j) ? -(double)Pin[i * n + j] : 0.0);
      else
        v = (double)Pin[j * n + i];
    }
    X[e] = v;
  }
  __syncthreads();
  // The Frechet derivative is LINEAR in G: normalise the G block to unit 1-norm and scale the result back, so that
  // the number of squarings depends on ||S|| only (a large loss gradient used to add one 16 x 16 product per factor
  // of two of its norm: this kernel was 12 us of a 101 us config-2 step).
  __shared__ double s_gn;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double cs = 0.0;
    for (int r = 0; r < n; ++r) cs += fabs(X[r * m + n + c]);
    red[c] = cs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double g = 0.0;
    for (int c = 0; c < n; ++c) g = fmax(g, red[c]);
    s_gn = g;
  }
  __syncthreads();
  const double gn = s_gn;
  if (gn > 0.0 && isfinite(gn)) {
    const double ign = 1.0 / gn;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      int r = e / n, c = e - r * n;
      X[r * m + n + c] *= ign;
    }
  }
  __syncthreads();
  const double back = (gn > 0.0 && isfinite(gn)) ? gn : 1.0;
  // (the block matrix is not normal: the general bound applies to the derivative block)
  double* R = expm_inplace(X, W, m, red, sizeof(T) == 4 ? 1.0 : 0.5);
  // dS = top-right block; skew map: gP[i][j] = dS[i][j] - dS[j][i] for i < j, else 0
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int i = e / n, j = e - i * n;
    double d = R[i * m + n + j];
    if (skew) d = (i < j) ? d - R[j * m + n + i] : 0.0;
    d *= back;
    gP[e] = (T)d;
  }
}


// ---- small matrices (the headline's 8 x 8 feedback matrix): ONE THREAD PER ELEMENT of an M x M matrix (M = 8 | 16,
// smaller problems zero padded: exp(diag(X, 0)) = diag(exp X, I)), the same degree-12 Paterson-Stockmeyer polynomial
// and scaling rule as expm_inplace, but
//   * the row a thread multiplies with lives in registers, the dot products are fully unrolled (no index arithmetic,
//     no division), every thread keeps its own element of X .. X6 in registers for the elementwise combinations;
//   * products that share their left factor run in the same phase (X3, X4 after X2; X5, X6 after X3): 4 + s
//     dependent product phases instead of 7 + s, 7 + s block barriers instead of ~16 + s;
//   * norms by warp shuffles instead of serial column loops.
// The generic kernel took 7.5 / 10.3 us (forward / adjoint) of a ~50 us config-2 step.
template <int M>
struct SmallExpm {
  static constexpr int NN = M * M;
  double* buf;  // 5 matrices in shared memory
  int r, c, t;

  __device__ __forceinline__ double rowdot(const double (&a)[M], const double* B) const {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < M; k += 2) {
      s0 = fma(a[k], B[k * M + c], s0);
      s1 = fma(a[k + 1], B[(k + 1) * M + c], s1);
    }
    return s0 + s1;
  }
  __device__ __forceinline__ void loadrow(double (&a)[M], const double* A) const {
#pragma unroll
    for (int k = 0; k < M; ++k) a[k] = A[r * M + k];
  }
  // max over the block of the row sums of v (one value per thread, rows = groups of M adjacent lanes)
  __device__ __forceinline__ double max_rowsum(double v, double* red) const {
#pragma unroll
    for (int o = 1; o < M; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#pragma unroll
    for (int o = M; o < 32; o <<= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    double m = 0.0;
#pragma unroll
    for (int w = 0; w < NN / 32; ++w) m = fmax(m, red[w]);
    return m;
  }

  // x: this thread's element of X (already stored in buf[0 .. NN) and published).  Returns the buffer holding exp(X).
  __device__ double* run(double x, double* red, double theta) {
    double *X = buf, *X2 = buf + NN, *X3 = buf + 2 * NN, *X6 = buf + 3 * NN, *T0 = buf + 4 * NN;
    // 1-norm: thread (r, c) takes |X[c][r]|, so the row sums of what the threads hold are the column sums of X
    const double m = max_rowsum(fabs(X[c * M + r]), red);
    int s = 0;
    if (m > theta) {
      int ex;
      frexp(m / theta, &ex);
      s = ex;
    }
    if (s > 60) s = 60;
    x *= ldexp(1.0, -s);
    __syncthreads();  // every thread has read X (transposed) before it is overwritten
    X[t] = x;
    __syncthreads();
    double a[M];
    loadrow(a, X);
    const double x2 = rowdot(a, X);
    X2[t] = x2;
    __syncthreads();
    loadrow(a, X2);
    const double x3 = rowdot(a, X), x4 = rowdot(a, X2);
    X3[t] = x3;
    __syncthreads();
    loadrow(a, X3);
    const double x5 = rowdot(a, X2), x6 = rowdot(a, X3);
    const double id = (r == c) ? 1.0 : 0.0;
    X6[t] = x6;
    T0[t] = id * (1.0 / 720) + x * (1.0 / 5040) + x2 * (1.0 / 40320) + x3 * (1.0 / 362880) + x4 * (1.0 / 3628800) +
            x5 * (1.0 / 39916800) + x6 * (1.0 / 479001600);
    __syncthreads();
    loadrow(a, X6);
    double e = rowdot(a, T0) + (id + x + x2 * (1.0 / 2) + x3 * (1.0 / 6) + x4 * (1.0 / 24) + x5 * (1.0 / 120));
    double *cur = X, *other = X2;  // (both free by now)
    cur[t] = e;
    __syncthreads();
    for (int i = 0; i < s; ++i) {
      loadrow(a, cur);
      e = rowdot(a, cur);
      other[t] = e;
      __syncthreads();
      double* tmp = cur;
      cur = other;
      other = tmp;
    }
    return cur;
  }
};

template <typename T, int M>
__global__ void __launch_bounds__(M * M) expm_fwd_small_kernel(const T* __restrict__ Pin, T* __restrict__ E, int n,
                                                               int skew, T* __restrict__ sp) {
  __shared__ double buf[5 * M * M];
  __shared__ double red[8];
  SmallExpm<M> ex;
  ex.buf = buf;
  ex.t = threadIdx.x;
  ex.r = threadIdx.x / M;
  ex.c = threadIdx.x % M;
  const int r = ex.r, c = ex.c;
  pdl_sync();
  double v = 0.0;
  if (r < n && c < n) {
    if (skew)
      v = (c > r) ? (double)Pin[r * n + c] : ((c < r) ? -(double)Pin[c * n + r] : 0.0);
    else
      v = (double)Pin[r * n + c];
  }
  buf[ex.t] = v;
  __syncthreads();
  const double* R = ex.run(v, red, sizeof(T) == 4 ? (skew ? 2.0 : 1.0) : 0.5);
  double asum = 0.0;
  if (r < n && c < n) {
    const T o = (T)R[ex.t];
    E[r * n + c] = o;
    asum = fabs((double)o);
  }
  if (sp != nullptr) {  // (uniform)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    __syncthreads();  // red was read by every thread inside run()
    if ((ex.t & 31) == 0) red[ex.t >> 5] = asum;
    __syncthreads();
    if (ex.t == 0) {
      double tot = 0.0;
      for (int w = 0; w < M * M / 32; ++w) tot += red[w];
      const double rn = sqrt((double)n);
      *sp = (T)((tot - n * rn) / (n * (1.0 - rn)));
    }
  }
}

// adjoint, 2n <= 16: the block matrix [[S^T, G], [0, S^T]] on the 16 x 16 kernel (n < 8: the blocks sit at offset 0 and
// 8, the padding rows / columns are zero)
template <typename T>
__global__ void __launch_bounds__(256) expm_bwd_small_kernel(const T* __restrict__ Pin, const T* __restrict__ G,
                                                             T* __restrict__ gP, int n, int skew,
                                                             const T* __restrict__ Esp, const T* __restrict__ gsp,
                                                             const __grid_constant__ RiderArgs rd) {
  constexpr int M = 16, H = 8;
  __shared__ double buf[5 * M * M];
  __shared__ double red[8];
  SmallExpm<M> ex;
  ex.buf = buf;
  ex.t = threadIdx.x;
  ex.r = threadIdx.x / M;
  ex.c = threadIdx.x % M;
  const int r = ex.r, c = ex.c;
  const int i = r % H, j = c % H;
  pdl_sync();
  if (blockIdx.x == 1) {
    run_rider<T>(rd);
    return;
  }
  double v = 0.0;
  const bool gblock = r < H && c >= H && i < n && j < n;
  if (gblock) {
    if (G != nullptr) v = (double)G[i * n + j];
    if (gsp != nullptr) {
      const T a = Esp[i * n + j];
      const double rn = sqrt((double)n);
      const double k = (double)*gsp / (n * (1.0 - rn));
      v += a > T(0) ? k : (a < T(0) ? -k : 0.0);
    }
  } else if ((r < H) == (c < H) && i < n && j < n) {  // block (i, j) of S^T = S[j][i]
    if (skew)
      v = (i > j) ? (double)Pin[j * n + i] : ((i < j) ? -(double)Pin[i * n + j] : 0.0);
    else
      v = (double)Pin[j * n + i];
  }
  // The Frechet derivative is LINEAR in G: normalise the G block (here by n * max |G_ij| >= its 1-norm) and scale the
  // result back, so that the number of squarings depends on ||S|| only.
  double g = gblock ? fabs(v) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g = fmax(g, __shfl_xor_sync(0xffffffffu, g, o));
  if ((ex.t & 31) == 0) red[ex.t >> 5] = g;
  __syncthreads();
  double gn = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) gn = fmax(gn, red[w]);
  gn *= (double)n;
  const bool norm = gn > 0.0 && isfinite(gn);
  if (norm && gblock) v *= 1.0 / gn;
  buf[ex.t] = v;
  __syncthreads();  // (also: everyone has read red before run() writes it again)
  const double* R = ex.run(v, red, sizeof(T) == 4 ? 1.0 : 0.5);
  // dS = top-right block; skew map: gP[i][j] = dS[i][j] - dS[j][i] for i < j, else 0
  if (r < H && c >= H && i < n && j < n) {
    double d = R[i * M + H + j];
    if (skew) d = (i < j) ? d - R[j * M + H + i] : 0.0;
    if (norm) d *= gn;
    gP[i * n + j] = (T)d;
  }
}

}  // namespace

extern "C" FSWEEP_API int fsweep_expm_max_n(void) { return 28; }  // 8 matrices of (2n)^2 doubles in shared memory

template <typename T>
static int expm_forward_t(const T* P, T* E, int n, int skew, T* sp, cudaStream_t st) {
  static const bool small_ok = [] {
    const char* e = getenv("FSWEEP_EXPM_SMALL");
    return !(e && e[0] == '0');
  }();
  if (small_ok && n <= 16) {
    if (n <= 8)
      launch_pdl(expm_fwd_small_kernel<T, 8>, dim3(1), dim3(64), 0, st, P, E, n, skew, sp);
    else
      launch_pdl(expm_fwd_small_kernel<T, 16>, dim3(1), dim3(256), 0, st, P, E, n, skew, sp);
    return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
  }
  size_t smem = (size_t)(8 * n * n + n + 8) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(expm_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return FSWEEP_E_CUDA;
  }
  launch_pdl(expm_fwd_kernel<T>, dim3(1), dim3(EXPM_THREADS), smem, st, P, E, n, skew, sp);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

template <typename T>
static int expm_backward_t(const T* P, const T* G, T* gP, int n, int skew, const T* Esp, const T* gsp,
                           const fsweep_total_job_t* total, cudaStream_t st) {
  RiderArgs rd;
  if (!fill_rider(total, &rd)) return FSWEEP_E_BADARG;
  const dim3 grid(rd.on ? 2 : 1);
  static const bool small_ok = [] {
    const char* e = getenv("FSWEEP_EXPM_SMALL");
    return !(e && e[0] == '0');
  }();
  if (small_ok && n <= 8) {
    launch_pdl(expm_bwd_small_kernel<T>, grid, dim3(256), 0, st, P, G, gP, n, skew, Esp, gsp, rd);
    return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
  }
  const int m = 2 * n;
  size_t smem = (size_t)(8 * m * m + m + 8) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(expm_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return FSWEEP_E_CUDA;
  }
  launch_pdl(expm_bwd_kernel<T>, grid, dim3(EXPM_THREADS), smem, st, P, G, gP, n, skew, Esp, gsp, rd);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_expm_forward_sp(const void* P, void* E, int n, int skew, int dtype, void* sparsity,
                                                 void* stream) {
  if (!P || !E || n < 1 || n > 2 * fsweep_expm_max_n() || (sparsity && n < 2)) return FSWEEP_E_BADARG;
  if (dtype == FSWEEP_C64)
    return expm_forward_t<float>((const float*)P, (float*)E, n, skew, (float*)sparsity, (cudaStream_t)stream);
  if (dtype == FSWEEP_C128)
    return expm_forward_t<double>((const double*)P, (double*)E, n, skew, (double*)sparsity, (cudaStream_t)stream);
  return FSWEEP_E_BADARG;
}

extern "C" FSWEEP_API int fsweep_expm_forward(const void* P, void* E, int n, int skew, int dtype, void* stream) {
  return fsweep_expm_forward_sp(P, E, n, skew, dtype, nullptr, stream);
}

extern "C" FSWEEP_API int fsweep_expm_backward_sp_total(const void* P, const void* G, void* gP, int n, int skew,
                                                        int dtype, const void* E, const void* gsparsity,
                                                        const fsweep_total_job_t* total, void* stream) {
  if (!P || (!G && !gsparsity) || !gP || n < 1 || n > fsweep_expm_max_n() || (gsparsity && (!E || n < 2)))
    return FSWEEP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSWEEP_C64)
    return expm_backward_t<float>((const float*)P, (const float*)G, (float*)gP, n, skew, (const float*)E,
                                  (const float*)gsparsity, total, st);
  if (dtype == FSWEEP_C128)
    return expm_backward_t<double>((const double*)P, (const double*)G, (double*)gP, n, skew, (const double*)E,
                                   (const double*)gsparsity, total, st);
  return FSWEEP_E_BADARG;
}

extern "C" FSWEEP_API int fsweep_expm_backward_sp(const void* P, const void* G, void* gP, int n, int skew, int dtype,
                                                  const void* E, const void* gsparsity, void* stream) {
  return fsweep_expm_backward_sp_total(P, G, gP, n, skew, dtype, E, gsparsity, nullptr, stream);
}

extern "C" FSWEEP_API int fsweep_expm_backward(const void* P, const void* G, void* gP, int n, int skew, int dtype,
                                               void* stream) {
  if (!G) return FSWEEP_E_BADARG;
  return fsweep_expm_backward_sp(P, G, gP, n, skew, dtype, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------
// sparsity_loss (reference flamo/optimize/loss.py:36-63) of the mapped feedback matrix A (n_mats x n x n):
//   loss = mean_i ( (sum |A_i| - n sqrt n) / (n (1 - sqrt n)) ),   dloss/dA = sign(A) / (n (1 - sqrt n) n_mats)
// One launch each way instead of ~6 elementwise / reduction launches of a few microseconds each — inside a
// captured training step the launch count is what these parameter-sized ops cost.
namespace {

template <typename T>
__global__ void __launch_bounds__(256) sparsity_fwd_kernel(const T* __restrict__ A, int n_mats, int n, T* loss) {
  __shared__ double red[8];
  const long long total = (long long)n_mats * n * n;
  double s = 0.0;
  for (long long e = threadIdx.x; e < total; e += blockDim.x) s += fabs((double)A[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    const double rn = sqrt((double)n);
    *loss = (T)((t / n_mats - n * rn) / (n * (1.0 - rn)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) sparsity_bwd_kernel(const T* __restrict__ A, const T* __restrict__ gloss,
                                                           int n_mats, int n, T* __restrict__ gA) {
  const long long total = (long long)n_mats * n * n;
  const double rn = sqrt((double)n);
  const double k = (double)*gloss / (n * (1.0 - rn) * n_mats);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const T a = A[e];
    gA[e] = (T)(a > T(0) ? k : (a < T(0) ? -k : 0.0));
  }
}

}  // namespace

extern "C" FSWEEP_API int fsweep_sparsity_forward(const void* A, int n_mats, int n, int dtype, void* loss, void* stream) {
  if (!A || !loss || n_mats < 1 || n < 2) return FSWEEP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSWEEP_C64)
    sparsity_fwd_kernel<float><<<1, 256, 0, st>>>((const float*)A, n_mats, n, (float*)loss);
  else if (dtype == FSWEEP_C128)
    sparsity_fwd_kernel<double><<<1, 256, 0, st>>>((const double*)A, n_mats, n, (double*)loss);
  else
    return FSWEEP_E_BADARG;
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_sparsity_backward(const void* A, const void* gloss, int n_mats, int n, int dtype,
                                                   void* gA, void* stream) {
  if (!A || !gloss || !gA || n_mats < 1 || n < 2) return FSWEEP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n_mats * n * n;
  const int grid = (int)((total + 255) / 256 > 1024 ? 1024 : (total + 255) / 256);
  if (dtype == FSWEEP_C64)
    sparsity_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)A, (const float*)gloss, n_mats, n, (float*)gA);
  else if (dtype == FSWEEP_C128)
    sparsity_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)A, (const double*)gloss, n_mats, n, (double*)gA);
  else
    return FSWEEP_E_BADARG;
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------
// Weighted total of the step's criteria (reference optimize/trainer.py:184-188: loss += alpha * criterion):
//   vals[i] = scale_i * part_i  (i < n),   vals[n] = sum_i alpha_i * vals[i]
// (scale_i: the multi-GPU trainer's shard weights; 1 otherwise)
// in ONE launch instead of a mul + add per criterion and a stack; the values land in the buffer the Trainer reads
// back once per step.
namespace {
template <typename T>
__global__ void weighted_total_kernel(const __grid_constant__ TotalArgs a, T* vals, void* host_vals,
                                      volatile int* host_seq, int* seq_counter) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  total_job<T>(a, vals, host_vals, host_seq, seq_counter);
}
}  // namespace

extern "C" FSWEEP_API int fsweep_weighted_total_notify(const void* const* parts, const double* alphas,
                                                       const double* scales, int n, int dtype, void* vals,
                                                       void* host_vals, void* host_seq, void* seq_counter,
                                                       void* stream) {
  if (!parts || !alphas || !scales || !vals || n < 1 || n > FSWEEP_MAX_CRITERIA) return FSWEEP_E_BADARG;
  if (host_vals && (!host_seq || !seq_counter)) return FSWEEP_E_BADARG;
  TotalArgs a;
  a.n = n;
  for (int i = 0; i < n; ++i) {
    if (!parts[i]) return FSWEEP_E_BADARG;
    a.part[i] = parts[i];
    a.alpha[i] = alphas[i];
    a.scale[i] = scales[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSWEEP_C64)
    launch_pdl(weighted_total_kernel<float>, dim3(1), dim3(32), 0, st, a, (float*)vals, host_vals,
               (volatile int*)host_seq, (int*)seq_counter);
  else if (dtype == FSWEEP_C128)
    launch_pdl(weighted_total_kernel<double>, dim3(1), dim3(32), 0, st, a, (double*)vals, host_vals,
               (volatile int*)host_seq, (int*)seq_counter);
  else
    return FSWEEP_E_BADARG;
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_weighted_total(const void* const* parts, const double* alphas, const double* scales,
                                                int n, int dtype, void* vals, void* stream) {
  return fsweep_weighted_total_notify(parts, alphas, scales, n, dtype, vals, nullptr, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------
// One-shot all-reduce of the (small) flat gradient buffer over NVLink peer memory: the one exchange of a multi-GPU
// training step (80 - 4224 floats, pure latency).  Every rank's buffer lives in symmetric memory
// (torch.distributed._symmetric_memory provides the allocation and the peer pointers — plumbing); ONE kernel per
// rank does the whole collective: announce "my gradients are final" in every peer's signal pad (release), wait for
// every peer's announcement (acquire), read ALL peers' buffers with P2P loads and sum them in rank order (every rank
// gets bit-identical sums), announce "done reading", wait for everyone, then overwrite its own buffer with
// scale * sum.  Epochs increase monotonically in device memory, so the launch can be replayed inside a CUDA graph.
// A bounded spin (about a second) turns a missing peer into an ERROR FLAG (epoch_ctr[1] = 1 + the missing rank, sticky;
// the host checks it, flamo_b200/parallel.py) instead of a hung GPU.
namespace {
constexpr int AR_THREADS = 1024;
constexpr int AR_PER = 8;          // values per thread: n <= 8192
constexpr int AR_PAD_OFF = 256;    // uint32 offset of our flags inside the signal pad (torch's own barriers use the head)

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_p2p_kernel(float* const* __restrict__ bufs,
                                                                  unsigned* const* __restrict__ pads, int rank, int world,
                                                                  int n, float scale, unsigned* epoch_ctr) {
  __shared__ unsigned s_epoch;
  const int tid = threadIdx.x;
  if (tid == 0) s_epoch = *epoch_ctr + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  auto barrier = [&](int phase) {
    // every thread's peer loads / stores of the preceding phase must have completed before ANY thread of this block
    // releases its flag: without this, a peer could pass the barrier and overwrite a buffer that slower warps of
    // this block are still reading
    __syncthreads();
    if (tid < world) {
      __threadfence_system();
      st_release_sys(pads[tid] + AR_PAD_OFF + phase * 64 + rank, epoch);
      const unsigned* mine = pads[rank] + AR_PAD_OFF + phase * 64 + tid;
      unsigned spins = 0;
      while ((int)(ld_acquire_sys(mine) - epoch) < 0 && ++spins < (1u << 26)) {
      }
      if ((int)(ld_acquire_sys(mine) - epoch) < 0) atomicExch(epoch_ctr + 1, 1u + (unsigned)tid);  // sticky: peer `tid` never arrived
    }
    __syncthreads();
  };
  barrier(0);  // every rank's gradients are final and visible
  float acc[AR_PER];
#pragma unroll
  for (int i = 0; i < AR_PER; ++i) {
    const int idx = tid + i * AR_THREADS;
    float s = 0.f;
    if (idx < n)
      for (int r = 0; r < world; ++r) s += __ldcv(bufs[r] + idx);  // fixed rank order: identical on every rank
    acc[i] = s;
  }
  barrier(1);  // everyone has read everyone: the buffers may be overwritten
#pragma unroll
  for (int i = 0; i < AR_PER; ++i) {
    const int idx = tid + i * AR_THREADS;
    if (idx < n) bufs[rank][idx] = acc[i] * scale;
  }
  if (tid == 0) *epoch_ctr = epoch;
}
}  // namespace

extern "C" FSWEEP_API int fsweep_allreduce_p2p_max_n(void) { return AR_THREADS * AR_PER; }

extern "C" FSWEEP_API int fsweep_allreduce_p2p(void* const* peer_buffers, void* const* peer_signal_pads, int rank,
                                               int world, int n, double scale, void* epoch_counter, void* stream) {
  if (!peer_buffers || !peer_signal_pads || !epoch_counter || world < 1 || world > 64 || rank < 0 || rank >= world ||
      n < 1 || n > AR_THREADS * AR_PER)
    return FSWEEP_E_BADARG;
  allreduce_p2p_kernel<<<1, AR_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float* const*>(peer_buffers), reinterpret_cast<unsigned* const*>(peer_signal_pads), rank, world, n,
      (float)scale, reinterpret_cast<unsigned*>(epoch_counter));
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------
// The same exchange as a PUSH, with the packing and unpacking inside: the step's gradients stay where autograd left
// them (one tensor per parameter) and the kernel (1) gathers the segments and STORES them straight into slot [rank] of
// every peer's receive area over NVLink (posted writes: nothing waits for a round trip), each value together with the
// step's epoch in one 8-byte store, (2) polls the `world` slots of its OWN receive area (local loads) until they carry
// the epoch, sums them in fixed rank order (bit-identical on every rank), scales, and scatters the result back into the
// segments.  No fence, no flag round (until the end of round 2: a system fence, one flag per peer, a barrier).  The
// receive areas are double buffered on the epoch's parity: a rank that writes parity p again (two steps later) has
// passed the step in between, which needs every peer's values of that step — sent only after the peer finished reading p.
// Replaces torch.cat + fsweep_allreduce_p2p (two flag rounds, peer loads) + views in flamo_b200/parallel.py.
namespace {
struct SegArgs {
  float* ptr[FSWEEP_AR_MAX_SEGS];
  int off[FSWEEP_AR_MAX_SEGS + 1];  // prefix offsets in floats; off[n_segs] = n
  int n_segs;
};

__global__ void __launch_bounds__(AR_THREADS) allreduce_push_kernel(const __grid_constant__ SegArgs sg,
                                                                   float* const* __restrict__ bufs,
                                                                   unsigned* const* __restrict__ pads, int rank, int world,
                                                                   int cap, float scale, unsigned* epoch_ctr,
                                                                   unsigned long long* host_vals, volatile int* host_seq,
                                                                   int* seq_counter) {
  __shared__ unsigned s_epoch;
  __shared__ int s_seq;
  const int tid = threadIdx.x;
  pdl_sync();
  if (tid == 0) {
    s_epoch = *epoch_ctr + 1u;
    if (host_vals) {
      s_seq = *seq_counter + 1;
      *seq_counter = s_seq;
    }
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int n = sg.off[sg.n_segs];
  // Every value travels WITH the epoch in one aligned 8-byte store {value, epoch} (the "LL" protocol of the
  // collective libraries): an 8-byte store is single-copy atomic, so a receiver that sees the epoch has the value — no
  // system fence, no flag round, no barrier between pushing and summing.  Receive areas: 8-byte slots.
  const size_t par = (size_t)(epoch & 1u) * world * cap;  // this step's half of every receive area
  float* seg_ptr[AR_PER];
  // (1) gather + push
#pragma unroll
  for (int i = 0; i < AR_PER; ++i) {
    const int idx = tid + i * AR_THREADS;
    seg_ptr[i] = nullptr;
    if (idx < n) {
      int s_ = 0;
      while (idx >= sg.off[s_ + 1]) ++s_;
      seg_ptr[i] = sg.ptr[s_] + (idx - sg.off[s_]);
      const unsigned long long pk = (unsigned long long)__float_as_uint(*seg_ptr[i]) | ((unsigned long long)epoch << 32);
      for (int r = 0; r < world; ++r)
        reinterpret_cast<volatile unsigned long long*>(bufs[r])[par + (size_t)rank * cap + idx] = pk;
    }
  }
  // (2) wait for every rank's slot of the own area (local loads), sum in rank order, scatter back
  const volatile unsigned long long* mybuf = reinterpret_cast<const volatile unsigned long long*>(bufs[rank]) + par;
#pragma unroll
  for (int i = 0; i < AR_PER; ++i) {
    const int idx = tid + i * AR_THREADS;
    if (idx < n) {
      float s_ = 0.f;
      for (int r = 0; r < world; ++r) {
        unsigned long long pk = mybuf[(size_t)r * cap + idx];
        unsigned spins = 0;
        while ((unsigned)(pk >> 32) != epoch && ++spins < (1u << 24)) pk = mybuf[(size_t)r * cap + idx];
        if ((unsigned)(pk >> 32) != epoch) atomicExch(epoch_ctr + 1, 1u + (unsigned)r);  // sticky error flag
        s_ += __uint_as_float((unsigned)pk);
      }
      *seg_ptr[i] = s_ * scale;
      // the LAST segment (the step's loss values) also goes to the host, each value with the launch number in one
      // aligned 8-byte store (fsweep_weighted_total_notify's float32 layout): the host has the exchanged losses while
      // the rest of this kernel and the optimizer are still running
      if (host_vals != nullptr && idx >= sg.off[sg.n_segs - 1]) {
        const int k = idx - sg.off[sg.n_segs - 1];
        reinterpret_cast<volatile unsigned long long*>(host_vals)[k] =
            (unsigned long long)__float_as_uint(s_ * scale) | ((unsigned long long)(unsigned)s_seq << 32);
        if (idx == n - 1) *host_seq = s_seq;
        __threadfence_system();
      }
    }
  }
  if (tid == 0) *epoch_ctr = epoch;
}
}  // namespace

extern "C" FSWEEP_API int fsweep_allreduce_push_notify(const fsweep_seg_t* segs, int n_segs, void* const* peer_buffers,
                                                       void* const* peer_signal_pads, int rank, int world, int cap,
                                                       double scale, void* epoch_counter, void* host_vals,
                                                       void* host_seq, void* seq_counter, void* stream) {
  if (host_vals && (!host_seq || !seq_counter)) return FSWEEP_E_BADARG;
  if (!segs || n_segs < 1 || n_segs > FSWEEP_AR_MAX_SEGS || !peer_buffers || !peer_signal_pads || !epoch_counter ||
      world < 1 || world > 64 || rank < 0 || rank >= world || cap < 1)
    return FSWEEP_E_BADARG;
  SegArgs a;
  a.n_segs = n_segs;
  int64_t off = 0;
  for (int i = 0; i < n_segs; ++i) {
    if (!segs[i].ptr || segs[i].numel < 1) return FSWEEP_E_BADARG;
    a.ptr[i] = reinterpret_cast<float*>(segs[i].ptr);
    a.off[i] = (int)off;
    off += segs[i].numel;
    if (off > AR_THREADS * AR_PER || off > cap) return FSWEEP_E_BADARG;
  }
  a.off[n_segs] = (int)off;
  launch_pdl(allreduce_push_kernel, dim3(1), dim3(AR_THREADS), 0, (cudaStream_t)stream, a,
             reinterpret_cast<float* const*>(peer_buffers), reinterpret_cast<unsigned* const*>(peer_signal_pads), rank,
             world, cap, (float)scale, reinterpret_cast<unsigned*>(epoch_counter),
             reinterpret_cast<unsigned long long*>(host_vals), reinterpret_cast<volatile int*>(host_seq),
             reinterpret_cast<int*>(seq_counter));
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_allreduce_push(const fsweep_seg_t* segs, int n_segs, void* const* peer_buffers,
                                                void* const* peer_signal_pads, int rank, int world, int cap, double scale,
                                                void* epoch_counter, void* stream) {
  return fsweep_allreduce_push_notify(segs, n_segs, peer_buffers, peer_signal_pads, rank, world, cap, scale,
                                      epoch_counter, nullptr, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam's update, reference optimize/trainer.py:42: the Trainer's optimizer) for ALL parameters of a
// model in ONE launch: one block per parameter tensor, its step counter in device memory (read, used, written back by
// that block only), learning rate read from device memory (schedulers fill it in place).  torch's capturable fused
// Adam is two launches (a foreach add on the step counters, then the update).
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
namespace {
struct AdamArgs {
  fsweep_adam_tensor_t t[FSWEEP_ADAM_MAX_TENSORS];
  int n;
  double beta1, beta2, eps;
  RiderArgs rider;  // optional: block n of the grid
};

template <typename T>
__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ AdamArgs a, const float* __restrict__ lr) {
  __shared__ float s_step;
  pdl_sync();
  if ((int)blockIdx.x == a.n) {
    run_rider<T>(a.rider);
    return;
  }
  const fsweep_adam_tensor_t& q = a.t[blockIdx.x];
  float* step = reinterpret_cast<float*>(q.step);
  T* p = reinterpret_cast<T*>(q.param);
  const T* g = reinterpret_cast<const T*>(q.grad);
  T* m = reinterpret_cast<T*>(q.exp_avg);
  T* v = reinterpret_cast<T*>(q.exp_avg_sq);
  // the operands of this thread's first element are requested together with the step counter: one DRAM round trip in
  // front of the arithmetic instead of two (the parameters of the headline model are 8 .. 64 elements per tensor)
  const long long i0 = threadIdx.x;
  T g0 = T(0), m0 = T(0), v0 = T(0), p0 = T(0);
  if (i0 < q.numel) {
    g0 = g[i0];
    m0 = m[i0];
    v0 = v[i0];
    p0 = p[i0];
  }
  const float lr_v = *lr;
  if (threadIdx.x == 0) s_step = *step + 1.0f;
  __syncthreads();
  const double t = (double)s_step;
  const double bc1 = 1.0 - pow(a.beta1, t), bc2 = 1.0 - pow(a.beta2, t);
  const double step_size = (double)lr_v / bc1, bc2_sqrt = sqrt(bc2);
  const T b1 = (T)a.beta1, b2 = (T)a.beta2;
  for (long long i = i0; i < q.numel; i += blockDim.x) {
    const T gi = (i == i0) ? g0 : g[i], mo = (i == i0) ? m0 : m[i], vo = (i == i0) ? v0 : v[i];
    const T po = (i == i0) ? p0 : p[i];
    const T mi = mo + (gi - mo) * (T(1) - b1);      // lerp, as torch does
    const T vi = b2 * vo + (T(1) - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const T denom = (T)(sqrt((double)vi) / bc2_sqrt) + (T)a.eps;
    p[i] = po - (T)step_size * (mi / denom);
  }
  if (threadIdx.x == 0) *step = s_step;
}
}  // namespace

extern "C" FSWEEP_API int fsweep_adam_step_total(const fsweep_adam_tensor_t* tensors, int n, int dtype, const void* lr,
                                                 double beta1, double beta2, double eps,
                                                 const fsweep_total_job_t* total, void* stream) {
  if (!tensors || !lr || n < 1 || n > FSWEEP_ADAM_MAX_TENSORS) return FSWEEP_E_BADARG;
  AdamArgs a;
  a.n = n;
  if (!fill_rider(total, &a.rider)) return FSWEEP_E_BADARG;
  const int blocks = n + (total ? 1 : 0);
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].param || !tensors[i].grad || !tensors[i].exp_avg || !tensors[i].exp_avg_sq || !tensors[i].step ||
        tensors[i].numel < 1)
      return FSWEEP_E_BADARG;
    a.t[i] = tensors[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FSWEEP_C64)
    launch_pdl(adam_step_kernel<float>, dim3(blocks), dim3(256), 0, st, a, (const float*)lr);
  else if (dtype == FSWEEP_C128)
    launch_pdl(adam_step_kernel<double>, dim3(blocks), dim3(256), 0, st, a, (const float*)lr);
  else
    return FSWEEP_E_BADARG;
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_adam_step(const fsweep_adam_tensor_t* tensors, int n, int dtype, const void* lr,
                                           double beta1, double beta2, double eps, void* stream) {
  return fsweep_adam_step_total(tensors, n, dtype, lr, beta1, beta2, eps, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------- batch upload
// The batch of a step (reference optimize/trainer.py:176 `move_to_device`: inputs, targets) from PINNED host memory into
// the captured step's static device buffers as ONE kernel that reads the host tensors directly over PCIe (unified
// addressing: the pinned host pointer is a device pointer) instead of one DMA copy per tensor: the copies of the
// headline step move 384 KB + 8 B, and each DMA is bracketed by a compute <-> copy engine switch that costs more than the
// transfer.  All 16-byte chunks are requested at once (one per thread), cache-volatile loads (the host rewrites the
// buffers between steps).
namespace {
struct UploadArgs {
  const unsigned char* src[FSWEEP_UPLOAD_MAX];
  unsigned char* dst[FSWEEP_UPLOAD_MAX];
  long long bytes[FSWEEP_UPLOAD_MAX];
  int n;
};
__global__ void __launch_bounds__(256) upload_kernel(const __grid_constant__ UploadArgs a) {
  pdl_sync();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (int s = 0; s < a.n; ++s) {
    const bool al = ((reinterpret_cast<uintptr_t>(a.src[s]) | reinterpret_cast<uintptr_t>(a.dst[s])) & 15) == 0;
    const long long n16 = al ? a.bytes[s] / 16 : 0;
    const int4* s4 = reinterpret_cast<const int4*>(a.src[s]);
    int4* d4 = reinterpret_cast<int4*>(a.dst[s]);
    for (long long i = tid; i < n16; i += nt) d4[i] = __ldcv(s4 + i);
    for (long long i = n16 * 16 + tid; i < a.bytes[s]; i += nt) a.dst[s][i] = __ldcv(a.src[s] + i);
  }
}
}  // namespace

extern "C" FSWEEP_API int fsweep_upload(const void* const* host_src, void* const* dev_dst, const int64_t* bytes, int n,
                                        void* stream) {
  if (!host_src || !dev_dst || !bytes || n < 1 || n > FSWEEP_UPLOAD_MAX) return FSWEEP_E_BADARG;
  UploadArgs a;
  a.n = n;
  long long most = 0;
  for (int i = 0; i < n; ++i) {
    if (!host_src[i] || !dev_dst[i] || bytes[i] < 1) return FSWEEP_E_BADARG;
    a.src[i] = reinterpret_cast<const unsigned char*>(host_src[i]);
    a.dst[i] = reinterpret_cast<unsigned char*>(dev_dst[i]);
    a.bytes[i] = bytes[i];
    most = bytes[i] > most ? bytes[i] : most;
  }
  const long long blocks = (most / 16 + 255) / 256;
  launch_pdl(upload_kernel, dim3((unsigned)(blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks))), dim3(256), 0,
             (cudaStream_t)stream, a);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

// ---------------------------------------------------------------------------------------------- FP32 FMA peak probe
// The roofline denominator for the compute-bound sweeps ("derive + measure"): every thread runs 16
// independent FFMA chains for `iters` rounds; the caller times the launch with CUDA events and divides
// fsweep_fma_probe_flops() by the duration.  Not part of the product path.
namespace {
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float x, float y) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // keeps the chains alive
}
}  // namespace

extern "C" FSWEEP_API double fsweep_fma_probe_flops(int blocks, int iters) { return 2.0 * 16 * 4 * (double)iters * 256.0 * blocks; }

extern "C" FSWEEP_API int fsweep_fma_probe(void* out, int blocks, int iters, void* stream) {
  if (!out || blocks < 1 || iters < 1) return FSWEEP_E_BADARG;
  fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float*>(out), iters, 0.999f, 1e-4f);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

// ---------------------------------------------------------------------------------------------- Biquad designer
// dsp.Biquad / parallelBiquad with a low-pass or high-pass prototype (reference dsp.py:1494-1563 -> functional.py
// lowpass_filter / highpass_filter, :376-470): raw parameter (K, 2, ...) -> bounded map (clamp cut-off to [0, 1],
// 20 log10|gain| to [-60, 60] dB) -> RBJ taps (Q = 1/sqrt 2) -> the packed Taylor blocks of FSWEEP_OP_SOS, as ONE
// launch (and one for the adjoint) instead of ~60 parameter-sized PyTorch launches: config 1's captured step was
// nothing but those.  float64 arithmetic; the parameter is read / its gradient written in its own dtype.
namespace {
struct BqTaps {
  double b[3], a[3];
  double db0[3], da0[3];  // d taps / d x0 (raw cut-off)
  double db1[3];          // d b / d x1 (raw gain); a does not depend on it
};

__device__ __forceinline__ BqTaps biquad_taps(double x0, double x1, int highpass) {
  BqTaps t;
  const bool in0 = x0 >= 0.0 && x0 <= 1.0;
  const double v0 = fmin(fmax(x0, 0.0), 1.0);
  const double gdb_raw = 20.0 * log10(fabs(x1));
  const bool in1 = gdb_raw >= -60.0 && gdb_raw <= 60.0;
  const double gdb = fmin(fmax(gdb_raw, -60.0), 60.0);
  const double PI = 3.14159265358979323846;
  double sw, cw;
  sincos(PI * v0, &sw, &cw);
  const double alpha = sw * 0.70710678118654752440;
  const double g = pow(10.0, gdb / 20.0);
  const double sgn = highpass ? 1.0 : -1.0;      // h = (1 + sgn c) / 2
  const double h = 0.5 * (1.0 + sgn * cw);
  t.b[0] = g * h;
  t.b[1] = g * (highpass ? -(1.0 + cw) : (1.0 - cw));
  t.b[2] = g * h;
  t.a[0] = 1.0 + alpha;
  t.a[1] = -2.0 * cw;
  t.a[2] = 1.0 - alpha;
  // derivatives (torch.clamp passes the gradient on the closed interval)
  const double dw = in0 ? PI : 0.0;
  const double dc = -sw * dw, dal = cw * 0.70710678118654752440 * dw;
  const double dh = 0.5 * sgn * dc;
  t.db0[0] = g * dh;
  t.db0[1] = g * (highpass ? -dc : -dc);
  t.db0[2] = g * dh;
  t.da0[0] = dal;
  t.da0[1] = -2.0 * dc;
  t.da0[2] = -dal;
  // d g / d x1 = g ln10/20 * 20 / (ln10 x1) = g / x1 inside the clamp
  const double dg = (in1 && x1 != 0.0) ? g / x1 : 0.0;
  t.db1[0] = dg * h;
  t.db1[1] = dg * (highpass ? -(1.0 + cw) : (1.0 - cw));
  t.db1[2] = dg * h;
  return t;
}

// section e = (k, n_out index m, n_in index n) of a (K, 2, n_out, n_in) parameter (parallel: (K, 2, n), m = n)
template <typename T, bool BWD>
__global__ void __launch_bounds__(128) biquad_design_kernel(const T* __restrict__ param, double* __restrict__ packed,
                                                            const double* __restrict__ gpacked, T* __restrict__ gparam,
                                                            int K, int n_out, int n_in, int parallel, int highpass) {
  const int per = parallel ? n_in : n_out * n_in;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * per) return;
  const int k = e / per, r = e - k * per;
  const int m = parallel ? r : r / n_in, n = parallel ? r : r - m * n_in;
  const double x0 = (double)param[((size_t)k * 2 + 0) * per + r], x1 = (double)param[((size_t)k * 2 + 1) * per + r];
  const BqTaps t = biquad_taps(x0, x1, highpass);
  // packed layout [K][n_in][n_out][2][8] (parallel: [K][n][2][8])
  const size_t o = parallel ? ((size_t)k * n_in + n) * 16 : (((size_t)k * n_in + n) * n_out + m) * 16;
  if constexpr (!BWD) {
    double* q = packed + o;
    q[0] = t.b[0] + t.b[1] + t.b[2];
    q[1] = t.b[1] + 2.0 * t.b[2];
    q[2] = t.b[2];
    q[3] = 0.0;
    q[4] = t.a[0] + t.a[1] + t.a[2];
    q[5] = t.a[1] + 2.0 * t.a[2];
    q[6] = t.a[2];
    q[7] = 0.0;
    q[8] = t.b[0] - t.b[1] + t.b[2];
    q[9] = t.b[1] - 2.0 * t.b[2];
    q[10] = t.b[2];
    q[11] = 0.0;
    q[12] = t.a[0] - t.a[1] + t.a[2];
    q[13] = t.a[1] - 2.0 * t.a[2];
    q[14] = t.a[2];
    q[15] = 0.0;
  } else {
    const double* g = gpacked + o;
    // adjoint of the packing: gradient w.r.t. the taps
    const double gb0 = g[0] + g[8], gb1 = g[0] + g[1] - g[8] + g[9], gb2 = g[0] + 2.0 * g[1] + g[2] + g[8] - 2.0 * g[9] + g[10];
    const double ga0 = g[4] + g[12], ga1 = g[4] + g[5] - g[12] + g[13], ga2 = g[4] + 2.0 * g[5] + g[6] + g[12] - 2.0 * g[13] + g[14];
    const double d0 = gb0 * t.db0[0] + gb1 * t.db0[1] + gb2 * t.db0[2] + ga0 * t.da0[0] + ga1 * t.da0[1] + ga2 * t.da0[2];
    const double d1 = gb0 * t.db1[0] + gb1 * t.db1[1] + gb2 * t.db1[2];
    gparam[((size_t)k * 2 + 0) * per + r] = (T)d0;
    gparam[((size_t)k * 2 + 1) * per + r] = (T)d1;
  }
}
}  // namespace

extern "C" FSWEEP_API int fsweep_biquad_design(const void* param, int K, int n_out, int n_in, int parallel, int highpass,
                                               int dtype, void* packed, const void* gpacked, void* gparam, void* stream) {
  if (!param || K < 1 || n_out < 1 || n_in < 1 || (!packed && !(gpacked && gparam))) return FSWEEP_E_BADARG;
  const int total = K * (parallel ? n_in : n_out * n_in);
  const int grid = (total + 127) / 128;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bwd = packed == nullptr;
  if (dtype == FSWEEP_C64) {
    if (bwd)
      biquad_design_kernel<float, true><<<grid, 128, 0, st>>>((const float*)param, nullptr, (const double*)gpacked, (float*)gparam, K, n_out, n_in, parallel, highpass);
    else
      biquad_design_kernel<float, false><<<grid, 128, 0, st>>>((const float*)param, (double*)packed, nullptr, nullptr, K, n_out, n_in, parallel, highpass);
  } else if (dtype == FSWEEP_C128) {
    if (bwd)
      biquad_design_kernel<double, true><<<grid, 128, 0, st>>>((const double*)param, nullptr, (const double*)gpacked, (double*)gparam, K, n_out, n_in, parallel, highpass);
    else
      biquad_design_kernel<double, false><<<grid, 128, 0, st>>>((const double*)param, (double*)packed, nullptr, nullptr, K, n_out, n_in, parallel, highpass);
  } else {
    return FSWEEP_E_BADARG;
  }
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

// ---------------------------------------------------------------------------------------------- SVF designer
// dsp.SVF / parallelSVF with the GENERAL mixing (filter_type = None; reference dsp.py:2214-2232, 2343-2347): raw
// parameter (5, K, ...) = (f, R, mLP, mBP, mHP) before their activations
//     f = tan(pi/2 sigmoid(p0)),  R = softplus(p1) / ln 2,  mLP = p2 + 1,  mBP = p3 + 2,  mHP = p4 + 1
//     b = [f^2 mLP + f mBP + mHP, 2 f^2 mLP - 2 mHP, f^2 mLP - f mBP + mHP],  a = [f^2 + 2 R f + 1, 2 f^2 - 2, f^2 - 2 R f + 1]
// -> packed Taylor blocks of FSWEEP_OP_SOS, forward and adjoint, one launch each (float64 inside).
namespace {
template <typename T, bool BWD>
__global__ void __launch_bounds__(128) svf_design_kernel(const T* __restrict__ param, double* __restrict__ packed,
                                                         const double* __restrict__ gpacked, T* __restrict__ gparam,
                                                         int K, int n_out, int n_in, int parallel) {
  const int per = parallel ? n_in : n_out * n_in;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * per) return;
  const int k = e / per, r = e - k * per;
  const int m = parallel ? r : r / n_in, n = parallel ? r : r - m * n_in;
  const size_t cs = (size_t)K * per;  // stride between the five parameter planes
  const size_t pi_ = (size_t)k * per + r;
  const double p0 = (double)param[pi_], p1 = (double)param[cs + pi_], p2 = (double)param[2 * cs + pi_],
               p3 = (double)param[3 * cs + pi_], p4 = (double)param[4 * cs + pi_];
  const double HALF_PI = 1.57079632679489661923, LN2 = 0.69314718055994530942;
  const double sg = 1.0 / (1.0 + exp(-p0));
  const double f = tan(HALF_PI * sg);
  const double sp = p1 > 20.0 ? p1 : log1p(exp(p1));  // torch's softplus (threshold 20)
  const double R = sp / LN2;
  const double mLP = p2 + 1.0, mBP = p3 + 2.0, mHP = p4 + 1.0;
  const double f2 = f * f;
  const size_t o = parallel ? ((size_t)k * n_in + n) * 16 : (((size_t)k * n_in + n) * n_out + m) * 16;
  if constexpr (!BWD) {
    const double b0 = f2 * mLP + f * mBP + mHP, b1 = 2.0 * f2 * mLP - 2.0 * mHP, b2 = f2 * mLP - f * mBP + mHP;
    const double a0 = f2 + 2.0 * R * f + 1.0, a1 = 2.0 * f2 - 2.0, a2 = f2 - 2.0 * R * f + 1.0;
    double* q = packed + o;
    q[0] = b0 + b1 + b2;
    q[1] = b1 + 2.0 * b2;
    q[2] = b2;
    q[3] = 0.0;
    q[4] = a0 + a1 + a2;
    q[5] = a1 + 2.0 * a2;
    q[6] = a2;
    q[7] = 0.0;
    q[8] = b0 - b1 + b2;
    q[9] = b1 - 2.0 * b2;
    q[10] = b2;
    q[11] = 0.0;
    q[12] = a0 - a1 + a2;
    q[13] = a1 - 2.0 * a2;
    q[14] = a2;
    q[15] = 0.0;
  } else {
    const double* g = gpacked + o;
    const double gb0 = g[0] + g[8], gb1 = g[0] + g[1] - g[8] + g[9], gb2 = g[0] + 2.0 * g[1] + g[2] + g[8] - 2.0 * g[9] + g[10];
    const double ga0 = g[4] + g[12], ga1 = g[4] + g[5] - g[12] + g[13], ga2 = g[4] + 2.0 * g[5] + g[6] + g[12] - 2.0 * g[13] + g[14];
    // d / d f, R, mLP, mBP, mHP
    const double gf = gb0 * (2.0 * f * mLP + mBP) + gb1 * (4.0 * f * mLP) + gb2 * (2.0 * f * mLP - mBP) +
                      ga0 * (2.0 * f + 2.0 * R) + ga1 * (4.0 * f) + ga2 * (2.0 * f - 2.0 * R);
    const double gR = ga0 * (2.0 * f) - ga2 * (2.0 * f);
    const double gLP = (gb0 + 2.0 * gb1 + gb2) * f2;
    const double gBP = (gb0 - gb2) * f;
    const double gHP = gb0 - 2.0 * gb1 + gb2;
    const double dfdp0 = HALF_PI * sg * (1.0 - sg) * (1.0 + f2);
    const double dRdp1 = (p1 > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-p1))) / LN2;
    gparam[pi_] = (T)(gf * dfdp0);
    gparam[cs + pi_] = (T)(gR * dRdp1);
    gparam[2 * cs + pi_] = (T)gLP;
    gparam[3 * cs + pi_] = (T)gBP;
    gparam[4 * cs + pi_] = (T)gHP;
  }
}
}  // namespace

extern "C" FSWEEP_API int fsweep_svf_design(const void* param, int K, int n_out, int n_in, int parallel, int dtype,
                                            void* packed, const void* gpacked, void* gparam, void* stream) {
  if (!param || K < 1 || n_out < 1 || n_in < 1 || (!packed && !(gpacked && gparam))) return FSWEEP_E_BADARG;
  const int total = K * (parallel ? n_in : n_out * n_in);
  const int grid = (total + 127) / 128;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bwd = packed == nullptr;
  if (dtype == FSWEEP_C64) {
    if (bwd)
      svf_design_kernel<float, true><<<grid, 128, 0, st>>>((const float*)param, nullptr, (const double*)gpacked, (float*)gparam, K, n_out, n_in, parallel);
    else
      svf_design_kernel<float, false><<<grid, 128, 0, st>>>((const float*)param, (double*)packed, nullptr, nullptr, K, n_out, n_in, parallel);
  } else if (dtype == FSWEEP_C128) {
    if (bwd)
      svf_design_kernel<double, true><<<grid, 128, 0, st>>>((const double*)param, nullptr, (const double*)gpacked, (double*)gparam, K, n_out, n_in, parallel);
    else
      svf_design_kernel<double, false><<<grid, 128, 0, st>>>((const double*)param, (double*)packed, nullptr, nullptr, K, n_out, n_in, parallel);
  } else {
    return FSWEEP_E_BADARG;
  }
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}
