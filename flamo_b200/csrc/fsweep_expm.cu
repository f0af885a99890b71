// fsweep_expm.cu — E = exp(S), S = triu(P,1) - triu(P,1)^T (or S = P), and its adjoint, for the
// orthogonal map of dsp.Matrix (reference flamo/processor/dsp.py:649, functional.py:42-56).
//
// torch.matrix_exp copies the matrix norm to the host to pick its Pade degree, which synchronises
// and cannot be captured in a CUDA graph.  This kernel does the whole thing on the device: one CTA,
// float64, scaling-and-squaring with a degree-14 Taylor polynomial evaluated by Horner's rule on
// X = S / 2^s with ||X||_1 <= 1/2 (remainder 0.5^15/15! ~ 2e-17), matrices resident in shared memory.
// The adjoint uses the block-triangular identity  exp([[S^T, G], [0, S^T]]) = [[E^T, dS], [0, E^T]]
// (the same Frechet-derivative formula torch.autograd uses), then folds dS through the skew map.
#include <cuda_runtime.h>

#include "../../include/fsweep.h"

namespace {

constexpr int EXPM_THREADS = 256;
constexpr int TAYLOR_DEGREE = 14;

// C = A * B (n x n, shared memory, row-major)
__device__ void matmul(const double* A, const double* B, double* C, int n) {
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    double s = 0.0;
    for (int k = 0; k < n; ++k) s = fma(A[r * n + k], B[k * n + c], s);
    C[e] = s;
  }
  __syncthreads();
}

// in: X (n x n) in shared memory; out: exp(X) left in `P`; T is scratch.  Returns pointer to result.
__device__ double* expm_inplace(double* X, double* P, double* T, int n, double* red) {
  // 1-norm = max column sum
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += fabs(X[r * n + c]);
    red[c] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int c = 0; c < n; ++c) m = fmax(m, red[c]);
    int s = 0;
    if (m > 0.5) s = (int)ceil(log2(m / 0.5));
    if (s > 60) s = 60;
    red[n] = (double)s;
  }
  __syncthreads();
  const int s = (int)red[n];
  const double scale = ldexp(1.0, -s);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    X[e] *= scale;
    int r = e / n, c = e - r * n;
    P[e] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  // Horner: P <- I + X P / j, j = m .. 1
  for (int j = TAYLOR_DEGREE; j >= 1; --j) {
    matmul(X, P, T, n);
    const double inv = 1.0 / (double)j;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      int r = e / n, c = e - r * n;
      P[e] = ((r == c) ? 1.0 : 0.0) + T[e] * inv;
    }
    __syncthreads();
  }
  double* cur = P;
  double* other = T;
  for (int i = 0; i < s; ++i) {
    matmul(cur, cur, other, n);
    double* t = cur;
    cur = other;
    other = t;
  }
  return cur;
}

__global__ void __launch_bounds__(EXPM_THREADS) expm_fwd_kernel(const double* __restrict__ Pin, double* __restrict__ E,
                                                               int n, int skew) {
  extern __shared__ double sm[];
  double *X = sm, *P = sm + n * n, *T = sm + 2 * n * n, *red = sm + 3 * n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    double v = Pin[e];
    if (skew) v = (c > r) ? Pin[r * n + c] : ((c < r) ? -Pin[c * n + r] : 0.0);
    X[e] = v;
  }
  __syncthreads();
  double* R = expm_inplace(X, P, T, n, red);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) E[e] = R[e];
}

__global__ void __launch_bounds__(EXPM_THREADS) expm_bwd_kernel(const double* __restrict__ Pin,
                                                               const double* __restrict__ G,
                                                               double* __restrict__ gP, int n, int skew) {
  extern __shared__ double sm[];
  const int m = 2 * n;
  double *X = sm, *P = sm + m * m, *T = sm + 2 * m * m, *red = sm + 3 * m * m;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    int r = e / m, c = e - r * m;
    double v = 0.0;
    if (r < n && c >= n) {
      v = G[r * n + (c - n)];
    } else if ((r < n) == (c < n)) {
      int i = r % n, j = c % n;  // block (i, j) of S^T = S[j][i]
      if (skew)
        v = (i > j) ? Pin[j * n + i] : ((i < j) ? -Pin[i * n + j] : 0.0);
      else
        v = Pin[j * n + i];
    }
    X[e] = v;
  }
  __syncthreads();
  double* R = expm_inplace(X, P, T, m, red);
  // dS = top-right block; skew map: gP[i][j] = dS[i][j] - dS[j][i] for i < j, else 0
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int i = e / n, j = e - i * n;
    double d = R[i * m + n + j];
    if (skew) d = (i < j) ? d - R[j * m + n + i] : 0.0;
    gP[e] = d;
  }
}

}  // namespace

extern "C" FSWEEP_API int fsweep_expm_max_n(void) { return 48; }

extern "C" FSWEEP_API int fsweep_expm_forward(const double* P, double* E, int n, int skew, void* stream) {
  if (!P || !E || n < 1 || n > 2 * fsweep_expm_max_n()) return FSWEEP_E_BADARG;
  size_t smem = (size_t)(3 * n * n + n + 8) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(expm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return FSWEEP_E_CUDA;
  }
  expm_fwd_kernel<<<1, EXPM_THREADS, smem, (cudaStream_t)stream>>>(P, E, n, skew);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}

extern "C" FSWEEP_API int fsweep_expm_backward(const double* P, const double* G, double* gP, int n, int skew,
                                               void* stream) {
  if (!P || !G || !gP || n < 1 || n > fsweep_expm_max_n()) return FSWEEP_E_BADARG;
  const int m = 2 * n;
  size_t smem = (size_t)(3 * m * m + m + 8) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(expm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return FSWEEP_E_CUDA;
  }
  expm_bwd_kernel<<<1, EXPM_THREADS, smem, (cudaStream_t)stream>>>(P, G, gP, n, skew);
  return cudaGetLastError() == cudaSuccess ? FSWEEP_OK : FSWEEP_E_CUDA;
}
