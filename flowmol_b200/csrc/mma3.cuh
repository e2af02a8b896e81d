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

}  // namespace fm
