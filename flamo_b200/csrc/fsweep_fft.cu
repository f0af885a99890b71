// fsweep_fft.cu — the input transform of a training step: one real-to-complex FFT of the excitation per step
// (reference flamo/processor/dsp.py:69-93 dsp.FFT, :122-163 dsp.FFTAntiAlias: torch.fft.rfft(x, n=nfft, dim=1)).
//
// cuFFT plans nfft = 96000 as FIVE launches (fill, radix 3, 125, 128, real post-processing): 20.5 us of the headline
// step's 63 us of device time for 384 KB of data — launch bound.  Here the same transform is TWO launches of a
// four-step FFT on the half-length complex sequence z[n] = x[2n] + i x[2n+1], M = nfft / 2 = M1 * M2:
//
//   pass 1  (one block per residue n2 and signal): the M1-point DFT over n1 of z[M2 n1 + n2], times W_M^(n2 k1),
//           stored as Y[k1][n2];
//   pass 2  (one block per pair k1, M1 - k1 and signal): the two M2-point DFTs over n2 give Z[k1 + M1 k2] and its
//           mirror Z[M - k]; X[k] = E + W_N^k O with E, O the even / odd parts — the real-input split needs exactly the
//           pair a block holds — written straight into X[batch][nfft/2 + 1][channels].
//
// Inside a block an L-point DFT (L = a * b, radices <= 64) is two rounds of direct small DFTs in shared memory with
// twiddles from tables that are computed once per (nfft, device) in float64 (fsweep_rfft_table) and laid out in the
// order the threads read them: no sincos in the step, every twiddle exact to float32 rounding, every load coalesced.  Zero padding / cropping to nfft and the anti-alias
// envelope gamma^-n are folded into the load.  Sizes whose half does not split into M1 * M2 with both <= 1024 and
// radices <= 64 are refused (FSWEEP_E_UNSUPPORTED): the caller keeps cuFFT for those.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/fsweep.h"
#include "fsweep_pdl.cuh"

using fsweep::launch_pdl;
using fsweep::pdl_sync;

namespace {

struct FftShape {
  int N, M, M1, M2, a1, b1, a2, b2;
};

// best split L = a * b with a, b <= 64, minimal a + b; false if none
bool split_radix(int L, int* a, int* b) {
  int best = 1 << 30;
  for (int p = 1; p * p <= L; ++p) {
    if (L % p) continue;
    const int q = L / p;
    if (q > 64) continue;
    if (p + q < best) {
      best = p + q;
      *a = p;
      *b = q;
    }
  }
  return best < (1 << 30);
}

bool plan_shape_search(int64_t nfft, FftShape* s);

// (the search below walks ~1000 candidate splits: remembered per thread for the size last asked about — a model has one)
bool plan_shape(int64_t nfft, FftShape* s) {
  thread_local int64_t last_nfft = -1;
  thread_local bool last_ok = false;
  thread_local FftShape last_shape;
  if (nfft != last_nfft) {
    last_ok = plan_shape_search(nfft, &last_shape);
    last_nfft = nfft;
  }
  if (last_ok) *s = last_shape;
  return last_ok;
}

bool plan_shape_search(int64_t nfft, FftShape* s) {
  if (nfft < 512 || nfft > (int64_t)1 << 21 || (nfft & 1)) return false;
  const int N = (int)nfft, M = N / 2;
  int best = 1 << 30;
  for (int M2 = 8; M2 <= 1024; ++M2) {
    if (M % M2) continue;
    const int M1 = M / M2;
    if (M1 > 1024 || M1 < M2) continue;  // pass 2 pairs k1 with its mirror: the longer factor there keeps its grid wide
    int a1, b1, a2, b2;
    if (!split_radix(M1, &a1, &b1) || !split_radix(M2, &a2, &b2)) continue;
    const int cost = (a1 + b1) + 2 * (a2 + b2);
    if (cost < best) {
      best = cost;
      *s = FftShape{N, M, M1, M2, a1, b1, a2, b2};
    }
  }
  return best < (1 << 30);
}

struct FftArgs {
  const float* x;
  const float* env;  // optional envelope, nfft entries
  const float2* TA;  // [M1]          exp(-2 pi i j / M1): the radix twiddles of pass 1
  const float2* TB;  // [M2]          exp(-2 pi i j / M2): the radix twiddles of pass 2
  const float2* TC;  // [M2][M1]      W_M^(n2 k1(t)): the inter-pass twiddle of thread t of block n2
  const float2* TD;  // [M1/2+1][2][M2] W_N^k of the two output bins of thread t of block k1
  float2* Y;         // [signals][M1][M2]
  float2* X;         // [batch][M + 1][C]
  long long xbs;     // batch stride of x, in elements (time stride C, channel stride 1)
  int n_time, C;
  FftShape s;
  float scale;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ void cfma(float2& acc, float2 v, float2 w) {
  acc.x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc.x));
  acc.y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc.y));
}

// L-point DFTs of NSEQ sequences held in buf0[s * L + n] (published by a barrier), L = a * b = blockDim.x:
// thread t = hi * b + lo ends up with output bin k = hi + a * lo of every sequence.  buf0 is free after the inner barrier.
template <int NSEQ>
__device__ __forceinline__ int block_dft(const float2* buf0, float2* buf1, const float2* WL, int L, int a, int b,
                                         float2* out) {
  const int t = threadIdx.x;
  const int hi = t / b, lo = t - hi * b;
  float2 acc[NSEQ];
  // round A: (k_a, n_b) = (hi, lo): sum over n_a of in[b n_a + lo] W_a^(n_a k_a); W_a^j = WL[b j]
#pragma unroll
  for (int s = 0; s < NSEQ; ++s) acc[s] = make_float2(0.f, 0.f);
  {
    int widx = 0;
    const int wstep = b * hi;
#pragma unroll 4
    for (int na = 0; na < a; ++na) {
      const float2 w = WL[widx];
#pragma unroll
      for (int s = 0; s < NSEQ; ++s) cfma(acc[s], buf0[s * L + b * na + lo], w);
      widx += wstep;
      if (widx >= L) widx -= L;
    }
    const float2 tw = WL[lo * hi];
#pragma unroll
    for (int s = 0; s < NSEQ; ++s) buf1[s * L + t] = cmul(acc[s], tw);
  }
  __syncthreads();
  // round B: (k_a, k_b) = (hi, lo): sum over n_b of U[k_a][n_b] W_b^(n_b k_b); W_b^j = WL[a j]
#pragma unroll
  for (int s = 0; s < NSEQ; ++s) acc[s] = make_float2(0.f, 0.f);
  {
    int widx = 0;
    const int wstep = a * lo;
    const float2* u = buf1 + hi * b;
#pragma unroll 4
    for (int nb = 0; nb < b; ++nb) {
      const float2 w = WL[widx];
#pragma unroll
      for (int s = 0; s < NSEQ; ++s) cfma(acc[s], u[s * L + nb], w);
      widx += wstep;
      if (widx >= L) widx -= L;
    }
  }
#pragma unroll
  for (int s = 0; s < NSEQ; ++s) out[s] = acc[s];
  return hi + a * lo;
}

__global__ void __launch_bounds__(1024) rfft_pass1_kernel(const FftArgs A) {
  extern __shared__ float2 fsm[];
  const int L = A.s.M1, t = threadIdx.x, n2 = blockIdx.x, sig = blockIdx.y;
  float2* WL = fsm;
  float2* buf0 = fsm + L;
  float2* buf1 = buf0 + L;
  // the twiddle tables are constant — read ahead of the wait for the preceding kernel of the step (fsweep_pdl.cuh) —
  // and laid out in thread order: every load of a block is one contiguous run (the first kernels of a step run cold)
  WL[t] = __ldg(A.TA + t);
  const float2 tw = __ldg(A.TC + (size_t)n2 * L + t);  // W_M^(n2 k1) of the bin this thread ends up with
  pdl_sync();
  {
    const int b = sig / A.C, c = sig - b * A.C;
    const float* xs = A.x + (size_t)b * A.xbs + c;
    const int n = 2 * (t * A.s.M2 + n2);
    float re = 0.f, im = 0.f;
    if (n < A.n_time) re = __ldg(xs + (size_t)n * A.C);
    if (n + 1 < A.n_time) im = __ldg(xs + (size_t)(n + 1) * A.C);
    if (A.env != nullptr) {
      re *= __ldg(A.env + n);
      im *= __ldg(A.env + n + 1);
    }
    buf0[t] = make_float2(re, im);
  }
  __syncthreads();
  float2 v;
  const int k1 = block_dft<1>(buf0, buf1, WL, L, A.s.a1, A.s.b1, &v);
  A.Y[((size_t)sig * L + k1) * A.s.M2 + n2] = cmul(v, tw);
}

__global__ void __launch_bounds__(1024) rfft_pass2_kernel(const FftArgs A) {
  extern __shared__ float2 fsm[];
  const int L = A.s.M2, M1 = A.s.M1, M = A.s.M, t = threadIdx.x, sig = blockIdx.y;
  const int k1 = blockIdx.x, k1m = (M1 - k1) % M1;
  float2* WL = fsm;
  float2* buf0 = fsm + L;       // [2][L]
  float2* buf1 = buf0 + 2 * L;  // [2][L]
  WL[t] = __ldg(A.TB + t);
  const float2 wk0 = __ldg(A.TD + (size_t)(2 * k1) * L + t), wk1 = __ldg(A.TD + (size_t)(2 * k1 + 1) * L + t);  // W_N^k
  pdl_sync();  // pass 1 is complete from here on
  buf0[t] = A.Y[((size_t)sig * M1 + k1) * L + t];
  buf0[L + t] = A.Y[((size_t)sig * M1 + k1m) * L + t];
  __syncthreads();
  float2 z[2];
  const int k2 = block_dft<2>(buf0, buf1, WL, L, A.s.a2, A.s.b2, z);
  buf0[k2] = z[0];  // (round B reads buf1 only: buf0 is free)
  buf0[L + k2] = z[1];
  __syncthreads();
  const int b = sig / A.C, c = sig - b * A.C;
  float2* Xs = A.X + (size_t)b * (M + 1) * A.C + c;
  const int n_seq = (k1m == k1) ? 1 : 2;
  for (int s = 0; s < n_seq; ++s) {
    const int k = (s ? k1m : k1) + M1 * t;
    const float2 zk = buf0[s * L + t];
    const float2 zm = (k1 == 0) ? buf0[(L - t) % L] : buf0[(1 - s) * L + (L - 1 - t)];  // Z[M - k]
    // E = (Z[k] + conj Z[M-k]) / 2, O = (Z[k] - conj Z[M-k]) / (2i), X[k] = E + W_N^k O
    const float ex = 0.5f * (zk.x + zm.x), ey = 0.5f * (zk.y - zm.y);
    const float ox = 0.5f * (zk.y + zm.y), oy = -0.5f * (zk.x - zm.x);
    const float2 w = s ? wk1 : wk0;
    float2 r = make_float2(ex + (w.x * ox - w.y * oy), ey + (w.x * oy + w.y * ox));
    r.x *= A.scale;
    r.y *= A.scale;
    Xs[(size_t)k * A.C] = r;
    if (k == 0) Xs[(size_t)M * A.C] = make_float2((zk.x - zk.y) * A.scale, 0.f);  // Nyquist bin
  }
}

// table sections, in this order: TA | TB | TC | TD (see FftArgs); every entry exp(-2 pi i * num / den) in float64
__host__ __device__ inline size_t table_entries(const FftShape& s) {
  return (size_t)s.M1 + s.M2 + (size_t)s.M + (size_t)(s.M1 / 2 + 1) * 2 * s.M2;
}
__global__ void rfft_table_kernel(float2* T, const FftShape sh) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= table_entries(sh)) return;
  long long num, den;
  size_t e = j;
  if (e < (size_t)sh.M1) {
    num = (long long)e;
    den = sh.M1;
  } else if ((e -= sh.M1) < (size_t)sh.M2) {
    num = (long long)e;
    den = sh.M2;
  } else if ((e -= sh.M2) < (size_t)sh.M) {
    const int n2 = (int)(e / sh.M1), t = (int)(e % sh.M1);
    const int k1 = (t / sh.b1) + sh.a1 * (t % sh.b1);
    num = (long long)n2 * k1;
    den = sh.M;
  } else {
    e -= sh.M;
    const int k1 = (int)(e / (2 * (size_t)sh.M2)), r = (int)(e % (2 * (size_t)sh.M2));
    const int which = r / sh.M2, t = r % sh.M2;
    const int kk1 = which ? (sh.M1 - k1) % sh.M1 : k1;
    num = (long long)kk1 + (long long)sh.M1 * t;
    den = sh.N;
  }
  double sn, cs;
  sincospi(2.0 * (double)(num % den) / (double)den, &sn, &cs);
  T[j] = make_float2((float)cs, (float)-sn);
}

}  // namespace

extern "C" FSWEEP_API int fsweep_rfft_supported(int64_t nfft) {
  FftShape s;
  return plan_shape(nfft, &s) ? 1 : 0;
}

extern "C" FSWEEP_API size_t fsweep_rfft_workspace_bytes(int64_t nfft, int64_t signals) {
  return (size_t)(nfft / 2) * (size_t)signals * sizeof(float2);
}

extern "C" FSWEEP_API int64_t fsweep_rfft_table_entries(int64_t nfft) {
  FftShape s;
  return plan_shape(nfft, &s) ? (int64_t)table_entries(s) : 0;
}

extern "C" FSWEEP_API int fsweep_rfft_table(void* table, int64_t nfft, void* stream) {
  FftShape s;
  if (!table || !plan_shape(nfft, &s)) return FSWEEP_E_BADARG;
  const size_t n = table_entries(s);
  rfft_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(table), s);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_rfft(const void* x, int64_t batch, int64_t n_time, int64_t channels,
                                      int64_t x_batch_stride, int64_t nfft, double scale, const void* envelope,
                                      const void* table, void* workspace, size_t workspace_bytes, void* X,
                                      void* stream) {
  FftShape s;
  if (!x || !table || !workspace || !X || batch < 1 || channels < 1 || n_time < 1) return FSWEEP_E_BADARG;
  if (!plan_shape(nfft, &s)) return FSWEEP_E_UNSUPPORTED;
  const int64_t signals = batch * channels;
  if (signals > 65535) return FSWEEP_E_UNSUPPORTED;
  if (workspace_bytes < fsweep_rfft_workspace_bytes(nfft, signals)) return FSWEEP_E_WORKSPACE;
  FftArgs A;
  A.x = reinterpret_cast<const float*>(x);
  A.env = reinterpret_cast<const float*>(envelope);
  A.TA = reinterpret_cast<const float2*>(table);
  A.TB = A.TA + s.M1;
  A.TC = A.TB + s.M2;
  A.TD = A.TC + s.M;
  A.Y = reinterpret_cast<float2*>(workspace);
  A.X = reinterpret_cast<float2*>(X);
  A.xbs = x_batch_stride;
  A.n_time = (int)(n_time < nfft ? n_time : nfft);
  A.C = (int)channels;
  A.s = s;
  A.scale = (float)scale;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(rfft_pass1_kernel, dim3(s.M2, (unsigned)signals), dim3(s.M1), 3 * s.M1 * sizeof(float2), st, A);
  launch_pdl(rfft_pass2_kernel, dim3(s.M1 / 2 + 1, (unsigned)signals), dim3(s.M2), 5 * s.M2 * sizeof(float2), st, A);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}
