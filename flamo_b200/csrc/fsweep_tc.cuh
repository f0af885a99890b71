// fsweep_tc.cuh — tensor-core P A = L U of the per-bin 64 x 64 complex loop matrix (tcgen05 / TMEM, sm_100a).
//
// Used by fsweep_cta_kernel<BWD, /*TC=*/true> (fsweep_cta.cuh) in place of the SIMT warp-pipeline elimination.  The
// matrix A = I - D W lives in TENSOR MEMORY for the whole factorisation, as a real 128 x 64 accumulator tile:
//     TMEM lane r      (r < 64) : Re A[r][0..63]        lane 64 + r : Im A[r][0..63]
// A blocked right-looking elimination with IMPLICIT partial pivoting (rows never move; finished rows get zero
// multipliers) walks over 8 panels of NB = 8 columns:
//   1. the panel columns come back with tcgen05.ld (8 floats per thread) and go to shared memory;
//   2. ONE warp factors the 64 x 8 panel in registers (two rows per lane; pivot = one REDUX over a packed |.|^2 / row
//      key, pivot row through shuffles) — no barrier inside the panel;
//   3. every thread writes its row of the panel back to TMEM (the tile ends up holding L\U), builds its row of the MMA
//      A-operand  -L  (real 128 x 16: [-Re L | Im L ; -Im L | -Re L], K-major canonical layout, split into a TF32
//      "hi" and an exact remainder "lo"), and reads the trailing columns of its row so that the 8 pivot rows can be
//      collected in shared memory;
//   4. 64 threads (one per column) finish the block row U12 = L11^-1 A12[pivot rows] and write it as the B-operand
//      (real 64 x 16: [Re U ; Im U], hi and lo), zero for the columns that are already final;
//   5. one thread issues the rank-8 complex update of ALL rows as 3 x 2 tcgen05.mma.kind::tf32 instructions
//      (hi*hi + lo*hi + hi*lo, K = 16: the "3 x TF32" scheme — plain TF32 is 3.6e-3 off, the split 7e-7:
//      tools/tc_probe.cu), M = 128, N = 64, accumulating in place in TMEM.  Because the pivot rows' A-operand rows
//      hold their within-panel multipliers, the same instruction also turns them into the final U12 rows.
// The tile is read once more at the end and scattered into shared memory in pivot order for the substitutions.
// Reference op: torch.linalg.solve on the (B, M, N, N) loop matrices, flamo/processor/system.py:417-425.
#pragma once
#include <cstdint>

#include "fsweep_tpc.cuh"

namespace fsweep {
namespace tc {

constexpr int T = 128;         // threads per block = TMEM lanes
constexpr int NB = 8;          // panel width
constexpr int BLOCKS_PER_SM = 5;  // register / shared-memory budget the kernel is compiled for (TMEM allows 8)
constexpr int NCOL = 64;       // TMEM columns (= loop width, padded)
constexpr int PLD = NB + 1;    // row stride of the panel buffer (floats)
constexpr int ULD = 64;        // row stride of the pivot-row buffer (floats)
constexpr uint32_t LBO_A = 128 * 16, LBO_B = 64 * 16, SBO = 128;
constexpr int AOP_FLOATS = 128 * 2 * NB;  // 128 x 16
constexpr int BOP_FLOATS = 64 * 2 * NB;   // 64 x 16
// bytes of the factorisation scratch (must fit in the L\U area it shares: 64 x 65 float2 = 33280 B)
constexpr size_t scratch_bytes() {
  return (size_t)(2 * AOP_FLOATS + 2 * BOP_FLOATS + 2 * 64 * PLD + 2 * NB * ULD) * sizeof(float);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); no swizzle, base offset 0
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  const uint32_t accum = 1u;  // D += A B
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

// two 8-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t taddr, float* v, bool second) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  if (second)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr + 8)
                 : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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

// nearest TF32 value, ties away from zero (the tensor core truncates the 13 low mantissa bits of what it is given:
// round first).  Two integer instructions; cvt.rna.tf32.f32 measured 12 % of the kernel's instructions.
__device__ __forceinline__ float tf32_hi(float a) {
  return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
}

// Block-level state that lives across bins.
struct State {
  uint32_t tbase;   // TMEM base address of the 64-column tile
  uint32_t mbar;    // shared-memory address of the MMA-completion mbarrier
  uint32_t parity;  // its next phase
};

// once per block: TMEM allocation (warp 0) and mbarrier init
__device__ __forceinline__ void setup(State& S, uint32_t* s_tmem, uint64_t* s_bar) {
  const int t = threadIdx.x;
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(NCOL)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  S.tbase = *s_tmem;
  S.mbar = smem_u32(s_bar);
  S.parity = 0;
}

__device__ __forceinline__ void teardown(const State& S) {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(S.tbase), "r"(NCOL) : "memory");
}

// P A = L U of A = I - D W for one bin.
//   scratch : the (16-byte aligned) L\U area [64][CLD] float2; used as operand / panel scratch during the elimination
//             and filled with L\U in pivot order (reciprocal pivots on the diagonal) on return
//   sD      : [64] diagonal chain response (rows >= N: zero)
//   sPiv/sPos/sFin : [64] ints; sInv : [64] reciprocal pivots
template <int CLD>
__device__ __forceinline__ void lu(State& S, float2* scratch, const float2* sD, const float* __restrict__ Wfb, int N,
                                   int* sPiv, int* sPos, int* sFin, float2* sInv) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int r = t & 63, part = t >> 6;  // this thread's TMEM lane: Re (part 0) or Im (part 1) of row r
  float* fs = reinterpret_cast<float*>(scratch);
  float* sAhi = fs;
  float* sAlo = sAhi + AOP_FLOATS;
  float* sBhi = sAlo + AOP_FLOATS;
  float* sBlo = sBhi + BOP_FLOATS;
  float* sP = sBlo + BOP_FLOATS;   // [2][64][PLD]  panel, Re then Im
  float* sU = sP + 2 * 64 * PLD;   // [2][NB][ULD]  pivot rows of the panel (trailing columns), Re then Im
  const uint32_t tlane = S.tbase + ((uint32_t)(warp * 32) << 16);

  // ---- A = I - D W into the TMEM tile (this thread: one real row of 64 entries)
  {
    const float2 d = sD[r];
    const float dv = part ? d.y : d.x;
#pragma unroll 1
    for (int c0 = 0; c0 < NCOL; c0 += 8) {
      float v[8];
      float w8[8];
      if (N == NCOL) {  // full width: the row is 64 contiguous, 16-byte aligned floats
        const float4 wa = __ldg(reinterpret_cast<const float4*>(Wfb + r * NCOL + c0));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(Wfb + r * NCOL + c0 + 4));
        w8[0] = wa.x, w8[1] = wa.y, w8[2] = wa.z, w8[3] = wa.w, w8[4] = wb.x, w8[5] = wb.y, w8[6] = wb.z, w8[7] = wb.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = (r < N && c0 + i < N) ? __ldg(Wfb + r * N + c0 + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ((part == 0 && c0 + i == r) ? 1.f : 0.f) - dv * w8[i];
      tmem_st8(tlane + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  if (t < 64) {
    sPos[t] = 255;  // not a pivot row yet
    sFin[t] = NB;   // active: all NB multipliers of a panel are real multipliers
  }
  unsigned act = 3u;  // warp 0: bit i = row lane + 32 i has not been a pivot row yet
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);

#pragma unroll 1
  for (int p = 0; p < NCOL / NB; ++p) {
    const int c0 = p * NB;
    // ---- 1. panel columns: TMEM -> shared memory
    {
      float v[8];
      tmem_ld8(tlane + c0, v);
      float* dst = sP + (part * 64 + r) * PLD;
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = v[i];
    }
    __syncthreads();
    // ---- 2. warp 0 factors the 64 x 8 panel in registers (rows lane, lane + 32)
    if (warp == 0) {
      float2 c[2][NB];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j)
          c[i][j] = f2(sP[(lane + 32 * i) * PLD + j], sP[(64 + lane + 32 * i) * PLD + j]);
      int fin0 = (act & 1u) ? NB : 0, fin1 = (act & 2u) ? NB : 0;  // number of leading entries that are multipliers
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float m0 = c[0][j].x * c[0][j].x + c[0][j].y * c[0][j].y;
        const float m1 = c[1][j].x * c[1][j].x + c[1][j].y * c[1][j].y;
        const unsigned k0 = (act & 1u) ? ((__float_as_uint(m0) & ~63u) | (unsigned)(63 - lane)) + 64u : 0u;
        const unsigned k1 = (act & 2u) ? ((__float_as_uint(m1) & ~63u) | (unsigned)(31 - lane)) + 64u : 0u;
        // every lane forms the reciprocals of its own two candidates while the REDUX is in flight; the pivot's one is
        // then a select + shuffle away (no MUFU on the dependent chain)
        const float r0 = rcp_t(m0), r1 = rcp_t(m1);
        const float2 ic0 = f2(c[0][j].x * r0, -c[0][j].y * r0), ic1 = f2(c[1][j].x * r1, -c[1][j].y * r1);
        const unsigned key = __reduce_max_sync(FULL, max(k0, k1));
        const int pr = 63 - (int)(key & 63u);
        float2 inv = (pr & 32) ? ic1 : ic0;
        inv.x = __shfl_sync(FULL, inv.x, pr & 31);
        inv.y = __shfl_sync(FULL, inv.y, pr & 31);
        float2 l0 = f2(0.f, 0.f), l1 = f2(0.f, 0.f);
        if ((act & 1u) && lane != pr) c[0][j] = l0 = cmul2(c[0][j], inv);
        if ((act & 2u) && lane + 32 != pr) c[1][j] = l1 = cmul2(c[1][j], inv);
        if (lane == (pr & 31)) {
          act &= ~(1u << (pr >> 5));
          if (pr & 32)
            fin1 = j;
          else
            fin0 = j;
        }
        if (lane == 0) {
          sPiv[c0 + j] = pr;
          sPos[pr] = c0 + j;
          sInv[c0 + j] = inv;
        }
#pragma unroll
        for (int jj = j + 1; jj < NB; ++jj) {
          float2 u = (pr & 32) ? c[1][jj] : c[0][jj];
          u.x = __shfl_sync(FULL, u.x, pr & 31);
          u.y = __shfl_sync(FULL, u.y, pr & 31);
          c[0][jj] = cnma2(c[0][jj], l0, u);
          c[1][jj] = cnma2(c[1][jj], l1, u);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          sP[(lane + 32 * i) * PLD + j] = c[i][j].x;
          sP[(64 + lane + 32 * i) * PLD + j] = c[i][j].y;
        }
      sFin[lane] = fin0;
      sFin[lane + 32] = fin1;
    }
    __syncthreads();
    // ---- 3. every thread: panel row back to TMEM; its row of the A-operand; pivot rows' trailing columns out
    {
      const float* mine = sP + (part * 64 + r) * PLD;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = mine[i];
      tmem_st8(tlane + c0, v);
    }
    const bool last = (p == NCOL / NB - 1);
    if (!last) {
      const int fin = sFin[r];  // entries j < fin of the panel row are multipliers (NB: all; 0: finished before)
      const int pos = sPos[r];
      {
        // A-operand row t of  -L :  part 0: [-Re l | +Im l],  part 1: [-Im l | -Re l]
        const float* lre = sP + r * PLD;
        const float* lim = sP + (64 + r) * PLD;
        float a16[16];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const bool on = j < fin;
          const float re = on ? lre[j] : 0.f, im = on ? lim[j] : 0.f;
          a16[j] = part ? -im : -re;
          a16[NB + j] = part ? -re : im;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 hi, lo;
          hi.x = tf32_hi(a16[4 * q + 0]);
          hi.y = tf32_hi(a16[4 * q + 1]);
          hi.z = tf32_hi(a16[4 * q + 2]);
          hi.w = tf32_hi(a16[4 * q + 3]);
          lo.x = tf32_hi(a16[4 * q + 0] - hi.x);
          lo.y = tf32_hi(a16[4 * q + 1] - hi.y);
          lo.z = tf32_hi(a16[4 * q + 2] - hi.z);
          lo.w = tf32_hi(a16[4 * q + 3] - hi.w);
          *reinterpret_cast<float4*>(sAhi + q * (128 * 4) + t * 4) = hi;
          *reinterpret_cast<float4*>(sAlo + q * (128 * 4) + t * 4) = lo;
        }
      }
      // trailing columns of this row; the NB pivot rows of the panel park theirs in sU
      const bool is_piv = pos >= c0 && pos < c0 + NB;
      float* urow = sU + (part * NB + (pos - c0)) * ULD;
#pragma unroll 1
      for (int cc = c0 + NB; cc < NCOL; cc += 16) {
        float v[16];
        const bool two = cc + 8 < NCOL;  // (block-uniform)
        tmem_ld8x2(tlane + cc, v, two);
        if (is_piv) {
#pragma unroll
          for (int i = 0; i < 8; ++i) urow[cc + i] = v[i];
          if (two) {
#pragma unroll
            for (int i = 8; i < 16; ++i) urow[cc + i] = v[i];
          }
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    if (last) break;
    __syncthreads();
    // ---- 4. U12 = L11^-1 A12[pivot rows], one column per thread (64 threads) -> B-operand [Re U ; Im U]
    if (t < 64) {
      const int c = t;
      float4 hi[4], lo[4];
      if (c >= c0 + NB) {
        float2 u[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) u[j] = f2(sU[j * ULD + c], sU[(NB + j) * ULD + c]);
#pragma unroll
        for (int j = 1; j < NB; ++j) {
          const int pj = sPiv[c0 + j];
          const float* lre = sP + pj * PLD;
          const float* lim = sP + (64 + pj) * PLD;
#pragma unroll
          for (int i = 0; i < j; ++i) u[j] = cnma2(u[j], f2(lre[i], lim[i]), u[i]);
        }
        float b16[16];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          b16[j] = u[j].x;
          b16[NB + j] = u[j].y;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          hi[q].x = tf32_hi(b16[4 * q + 0]);
          hi[q].y = tf32_hi(b16[4 * q + 1]);
          hi[q].z = tf32_hi(b16[4 * q + 2]);
          hi[q].w = tf32_hi(b16[4 * q + 3]);
          lo[q].x = tf32_hi(b16[4 * q + 0] - hi[q].x);
          lo[q].y = tf32_hi(b16[4 * q + 1] - hi[q].y);
          lo[q].z = tf32_hi(b16[4 * q + 2] - hi[q].z);
          lo[q].w = tf32_hi(b16[4 * q + 3] - hi[q].w);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) hi[q] = lo[q] = make_float4(0.f, 0.f, 0.f, 0.f);  // final columns: untouched
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<float4*>(sBhi + q * (64 * 4) + c * 4) = hi[q];
        *reinterpret_cast<float4*>(sBlo + q * (64 * 4) + c * 4) = lo[q];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    // ---- 5. rank-8 complex update of the whole tile: D += (-L) U as 3 x TF32
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const float* pa = (pass == 1) ? sAlo : sAhi;
        const float* pb = (pass == 2) ? sBlo : sBhi;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t ad = make_desc(smem_u32(pa) + kk * 2 * LBO_A, LBO_A, SBO);
          const uint64_t bd = make_desc(smem_u32(pb) + kk * 2 * LBO_B, LBO_B, SBO);
          mma_tf32(S.tbase, ad, bd, idesc);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(S.mbar)
                   : "memory");
    }
    while (!mbar_try(S.mbar, S.parity)) {
    }
    S.parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  __syncthreads();  // (the scratch is dead from here on: every MMA has completed)
  // ---- L\U out of TMEM into shared memory in pivot order; reciprocal pivots on the diagonal
  {
    float* dstrow = reinterpret_cast<float*>(scratch + sPos[r] * CLD) + part;
#pragma unroll 1
    for (int cc = 0; cc < NCOL; cc += 8) {
      float v[8];
      tmem_ld8(tlane + cc, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) dstrow[2 * (cc + i)] = v[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (t < 64) scratch[t * CLD + t] = sInv[t];
  __syncthreads();
}

}  // namespace tc
}  // namespace fsweep
