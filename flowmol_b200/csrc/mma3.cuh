// Warp-level error-compensated GEMM for the small vector-channel contractions (K <= 40, N <= 48): mma.sync with the same
// compensation as the tcgen05 path (tc.cuh): a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate.
// These matrices are far too small for tcgen05 tiles (M128 x N>=32 per instruction, 64-cycle floor); on the CUDA cores
// they were instruction-bound (38 % FMA-pipe, ncu r01d).  Operands come straight from shared memory:
//   A [rows][lda] fp32 row-major (rows = edge x plane), W pre-split per CTA (below).  A first version used TF32 operands
//   (m16n8k8): cvt.rna.tf32 compiles to ~6 instructions on sm_100 and the splits were 60 % of the kernels' instructions (r01k).
#pragma once
#include "tc.cuh"

namespace fm {

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
