// Warp-level 3xTF32 GEMM for the small vector-channel contractions (K <= 40, N <= 48): mma.sync.m16n8k8 TF32 with the same
// error compensation as the tcgen05 path (tc.cuh): a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate.
// These matrices are far too small for tcgen05 tiles (M128 x N>=32 x K8 per instruction, 64-cycle floor); on the CUDA cores
// they were instruction-bound (38 % FMA-pipe, ncu r01d).  Operands come straight from shared memory:
//   A [rows][lda] row-major (rows = edge x plane), W [K][ldw] row-major with ldw % 32 == 8 (conflict-free fragment loads).
#pragma once
#include "tc.cuh"

namespace fm {

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[mt][nt][4] += A[m0 + 16 mt .. +16][0:K) x W[0:K)[n0 + 8 nt .. +8];  K multiple of 8.
// Fragment ownership (PTX ISA, m16n8k8 .tf32): g = lane / 4, t = lane % 4
//   a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4);   b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
template <int MT, int NTL>
__device__ __forceinline__ void warp_gemm_3xtf32(const float* __restrict__ A, int lda, int m0, const float* __restrict__ W, int ldw,
                                                 int n0, int K, float (&acc)[MT][NTL][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t bh[NTL][2], bl[NTL][2];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float hi, lo;
        tc::split_tf32(W[(k0 + t + 4 * i) * ldw + n0 + 8 * nt + g], hi, lo);
        bh[nt][i] = __float_as_uint(hi); bl[nt][i] = __float_as_uint(lo);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t ah[4], al[4];
      const float* ap = A + (m0 + 16 * mt + g) * lda + k0 + t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float hi, lo;
        tc::split_tf32(ap[(i & 1) * 8 * lda + (i >> 1) * 4], hi, lo);
        ah[i] = __float_as_uint(hi); al[i] = __float_as_uint(lo);
      }
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        mma_tf32_16x8x8(acc[mt][nt], al, bh[nt]);
        mma_tf32_16x8x8(acc[mt][nt], ah, bl[nt]);
        mma_tf32_16x8x8(acc[mt][nt], ah, bh[nt]);
      }
    }
  }
}


// ---- fp16x3 variant ----------------------------------------------------------------------------------------------------------
// Same error compensation on fp16 (hi, lo) operands (tc.cuh: "scaled fp16 hi/lo"): mma.sync.m16n8k16 consumes 16 k values per
// instruction and the fp32 -> (hi, lo) split is 3 instructions per element (F2FP / HADD2.F32 / FADD) instead of the ~12 the two
// cvt.rna.tf32 of a TF32 split compile to on sm_100 (ncu r01k: the splits were ~60 % of k_vec_b's instructions).  The weights are
// split ONCE per CTA into packed (k, k+1) half2 words, scaled by a power of two so that their `lo` parts stay normal:
//   Wh / Wl [ceil(K / 2)][ldw] words, ldw % 32 == 8 (conflict-free B fragments);  acc holds (A W) * scale.
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16_16x8x8(float (&d)[4], const uint32_t (&a)[2], const uint32_t b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(b));
}
// Fragment ownership (PTX ISA, m16n8k16 .f16): g = lane / 4, t = lane % 4; every register holds k and k + 1
//   a0 (g, 2t)  a1 (g + 8, 2t)  a2 (g, 2t + 8)  a3 (g + 8, 2t + 8);   b0 (k = 2t, n = g)  b1 (k = 2t + 8, n = g);   c as m16n8k8
// K is a multiple of 8 (a trailing half step runs as one m16n8k8); A [rows][lda] fp32 with lda even.
template <int MT, int NTL>
__device__ __forceinline__ void warp_gemm_h16x3(const float* __restrict__ A, int lda, int m0, const uint32_t* __restrict__ Wh,
                                                const uint32_t* __restrict__ Wl, int ldw, int n0, int K, float (&acc)[MT][NTL][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  int k0 = 0;
  for (; k0 + 16 <= K; k0 += 16) {
    uint32_t bh[NTL][2], bl[NTL][2];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int w = ((k0 >> 1) + t + 4 * i) * ldw + n0 + 8 * nt + g;
        bh[nt][i] = Wh[w]; bl[nt][i] = Wl[w];
      }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t ah[4], al[4];
      const float* ap = A + (m0 + 16 * mt + g) * lda + k0 + 2 * t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(ap + (i & 1) * 8 * lda + (i >> 1) * 8);
        tc::split_h16x2(v.x, v.y, ah[i], al[i]);
      }
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        mma_f16_16x8x16(acc[mt][nt], al, bh[nt]);
        mma_f16_16x8x16(acc[mt][nt], ah, bl[nt]);
        mma_f16_16x8x16(acc[mt][nt], ah, bh[nt]);
      }
    }
  }
  if (k0 < K) {                                             // 8 remaining k values
    uint32_t bh[NTL], bl[NTL];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      const int w = ((k0 >> 1) + t) * ldw + n0 + 8 * nt + g;
      bh[nt] = Wh[w]; bl[nt] = Wl[w];
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      uint32_t ah[2], al[2];
      const float* ap = A + (m0 + 16 * mt + g) * lda + k0 + 2 * t;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(ap + i * 8 * lda);
        tc::split_h16x2(v.x, v.y, ah[i], al[i]);
      }
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        mma_f16_16x8x8(acc[mt][nt], al, bh[nt]);
        mma_f16_16x8x8(acc[mt][nt], ah, bl[nt]);
        mma_f16_16x8x8(acc[mt][nt], ah, bh[nt]);
      }
    }
  }
}

}  // namespace fm
