// tc_probe.cu — hardware probe for the tcgen05 pieces fsweep_tc.cuh relies on (B200, sm_100a):
//   * TMEM alloc / tcgen05.st / tcgen05.ld lane+column mapping,
//   * kind::tf32 MMA (M = 128, N multiple of 16, K = 8 per instruction) with K-major, no-swizzle shared-memory
//     descriptors in the layout  offset(row, k) = (k / 4) * LBO + row * 16 + (k % 4) * 4  (SBO = 128),
//   * the 3 x TF32 split (hi*hi + lo*hi + hi*lo) against a float64 host product,
//   * cycle counts of an MMA chain and of tcgen05.ld sweeps (design calibration).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_probe tools/tc_probe.cu
// Run  :  tools/tc_probe <variant 0|1> <N> <K>      (variant 1 swaps the LBO / SBO descriptor fields)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = no swizzle, base offset 0
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

constexpr int TCOLS = 128;

// A: [128][K] row-major, B: [N][K] row-major (D = D0 + A * B^T), D: [128][N]
__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* A, const float* B, float* D, long long* cyc, int N, int K, int variant, int split) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x, warp = t >> 5;
  const int nch = K / 4;
  const uint32_t a_bytes = 128 * K * 4, b_bytes = N * K * 4;
  float* sAhi = reinterpret_cast<float*>(smem);
  float* sAlo = reinterpret_cast<float*>(smem + a_bytes);
  float* sBhi = reinterpret_cast<float*>(smem + 2 * a_bytes);
  float* sBlo = reinterpret_cast<float*>(smem + 2 * a_bytes + b_bytes);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // operands -> canonical layout, split into hi (tf32-exact) and lo
  for (int e = t; e < 128 * K; e += 128) {
    const int row = e / K, k = e % K;
    const float a = A[e];
    const float hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
    const int off = (k / 4) * (128 * 4) + row * 4 + (k % 4);
    sAhi[off] = split ? hi : a;
    sAlo[off] = a - hi;
  }
  for (int e = t; e < N * K; e += 128) {
    const int row = e / K, k = e % K;
    const float b = B[e];
    const float hi = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
    const int off = (k / 4) * (N * 4) + row * 4 + (k % 4);
    sBhi[off] = split ? hi : b;
    sBlo[off] = b - hi;
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = s_tmem;
  const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
  // D0[lane][col] = lane / 128 + col / 16384 through tcgen05.st
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (float)t * (1.f / 128.f) + (float)(c0 + i) * (1.f / 16384.f);
    tmem_st8(tlane + c0, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (t == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t lboA = 128 * 16, lboB = N * 16, sbo = 128;
    t0 = clock64();
    const int passes = split ? 3 : 1;
    for (int p = 0; p < passes; ++p) {
      const float* pa = (p == 1) ? sAlo : sAhi;
      const float* pb = (p == 2) ? sBlo : sBhi;
      for (int kk = 0; kk < K / 8; ++kk) {
        const uint32_t aaddr = smem_u32(pa) + kk * 2 * lboA, baddr = smem_u32(pb) + kk * 2 * lboB;
        const uint64_t ad = variant ? make_desc(aaddr, sbo, lboA) : make_desc(aaddr, lboA, sbo);
        const uint64_t bd = variant ? make_desc(baddr, sbo, lboB) : make_desc(baddr, lboB, sbo);
        mma_tf32(tbase, ad, bd, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar))
                 : "memory");
  }
  while (!mbar_try(smem_u32(&s_bar), 0)) {
  }
  if (t == 0) t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  long long t2 = clock64();
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tmem_ld8(tlane + c0, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) D[(size_t)t * N + c0 + i] = v[i];
  }
  long long t3 = clock64();
  // tcgen05.ld sweep timing: 64 sweeps of all N columns by all four warps
  float sink = 0.f;
  __syncthreads();
  long long t4 = clock64();
  for (int it = 0; it < 64; ++it)
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tmem_ld8(tlane + c0, v);
      sink += v[it & 7];
    }
  __syncthreads();
  long long t5 = clock64();
  if (sink == 12345.678f) D[0] = sink;
  if (t == 0) {
    cyc[0] = t1 - t0;
    cyc[1] = t3 - t2;
    cyc[2] = (t5 - t4) / 64;
    (void)nch;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(TCOLS) : "memory");
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = argc > 2 ? atoi(argv[2]) : 16;
  const int K = argc > 3 ? atoi(argv[3]) : 8;
  const int split = argc > 4 ? atoi(argv[4]) : 1;
  if (N % 16 || N < 16 || N > 128 || K % 8 || K < 8 || K > 32) {
    printf("bad N / K\n");
    return 1;
  }
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dC, 8 * sizeof(long long)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 2 * (size_t)128 * K * 4 + 2 * (size_t)N * K * 4;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dC, N, K, variant, split);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  long long cyc[8];
  CK(cudaMemcpy(cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  int bad_r = -1, bad_c = -1;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < N; ++c) {
      double ref = (double)r / 128.0 + (double)c / 16384.0;
      for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * (double)B[c * K + k];
      const double e = fabs(ref - (double)D[(size_t)r * N + c]);
      if (e > maxerr) {
        maxerr = e;
        bad_r = r;
        bad_c = c;
      }
      maxref = fmax(maxref, fabs(ref));
    }
  printf("variant %d N %d K %d split %d: max abs err %.3e (at %d,%d; max |ref| %.1f)  mma chain %lld cyc, ld %lld cyc, "
         "ld sweep of N cols x 4 warps %lld cyc\n",
         variant, N, K, split, maxerr, bad_r, bad_c, maxref, cyc[0], cyc[1], cyc[2]);
  return 0;
}
