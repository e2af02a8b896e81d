// Stand-alone check of the tcgen05 building blocks: D^T[f][e] = sum_k W[f][k] * X[e][k] with 3xTF32,
// W: 128 x K (UMMA A operand, M = 128), X: 64 x K (UMMA B operand, N = 64), K a multiple of 32.
#pragma once
#include "tc.cuh"

namespace fm {

constexpr int TCT_M = 128, TCT_N = 64;

__global__ void __launch_bounds__(128, 1) k_tc_gemm_test(const float* __restrict__ W, const float* __restrict__ X, int K,
                                                         float* __restrict__ out /*[128][64]*/, int passes /*1 or 3*/) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t tsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nslab = K / 32;
  // per slab: W_hi (16 KB) | W_lo (16 KB) | X_hi (8 KB) | X_lo (8 KB)
  const uint32_t SLAB = 48 * 1024;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm) + 1023) & ~uintptr_t(1023));
  for (int s = 0; s < nslab; ++s) {
    uint8_t* sb = base + (size_t)s * SLAB;
    for (int idx = tid; idx < TCT_M * 32; idx += 128) {
      const int r = idx >> 5, k = idx & 31;
      float hi, lo;
      tc::split_tf32(W[(size_t)r * K + s * 32 + k], hi, lo);
      *reinterpret_cast<float*>(sb + tc::sw128_off(r, k)) = hi;
      *reinterpret_cast<float*>(sb + 16384 + tc::sw128_off(r, k)) = lo;
    }
    for (int idx = tid; idx < TCT_N * 32; idx += 128) {
      const int r = idx >> 5, k = idx & 31;
      float hi, lo;
      tc::split_tf32(X[(size_t)r * K + s * 32 + k], hi, lo);
      *reinterpret_cast<float*>(sb + 32768 + tc::sw128_off(r, k)) = hi;
      *reinterpret_cast<float*>(sb + 40960 + tc::sw128_off(r, k)) = lo;
    }
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(TCT_M, TCT_N);
    uint32_t accum = 0;
    for (int s = 0; s < nslab; ++s) {
      const uint32_t sb = tc::smem_u32(base + (size_t)s * SLAB);
      for (int j = 0; j < 4; ++j) {
        const uint32_t ko = j * 32;            // 8 tf32 = 32 bytes per k-step inside the swizzle atom
        const uint64_t whi = tc::desc_sw128(sb + ko), wlo = tc::desc_sw128(sb + 16384 + ko);
        const uint64_t xhi = tc::desc_sw128(sb + 32768 + ko), xlo = tc::desc_sw128(sb + 40960 + ko);
        if (passes == 3) {
          tc::umma_tf32(tm, wlo, xhi, idesc, accum); accum = 1;
          tc::umma_tf32(tm, whi, xlo, idesc, accum);
        }
        tc::umma_tf32(tm, whi, xhi, idesc, accum); accum = 1;
      }
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  float v[32];
  for (int c = 0; c < TCT_N; c += 32) {
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[(size_t)tid * TCT_N + c + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 64);
}

// Same check for the scaled fp16 hi/lo operands ("fp16x3", kind::f16): K a multiple of 64, slabs of 64 k values with the byte
// geometry of the TF32 slabs.  `w_scale` is the power-of-two weight scale the host packer would choose.
__global__ void __launch_bounds__(128, 1) k_tc_gemm_test_h16(const float* __restrict__ W, const float* __restrict__ X, int K,
                                                             float* __restrict__ out /*[128][64]*/, float w_scale) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t tsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nslab = K / 64;
  const uint32_t SLAB = 48 * 1024;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm) + 1023) & ~uintptr_t(1023));
  for (int s = 0; s < nslab; ++s) {
    uint8_t* sb = base + (size_t)s * SLAB;
    for (int idx = tid; idx < TCT_M * 32; idx += 128) {          // two k values per step
      const int r = idx >> 5, k = (idx & 31) * 2;
      uint32_t hi, lo;
      tc::split_h16x2(W[(size_t)r * K + s * 64 + k] * w_scale, W[(size_t)r * K + s * 64 + k + 1] * w_scale, hi, lo);
      *reinterpret_cast<uint32_t*>(sb + tc::sw128_off_h(r, k)) = hi;
      *reinterpret_cast<uint32_t*>(sb + 16384 + tc::sw128_off_h(r, k)) = lo;
    }
    for (int idx = tid; idx < TCT_N * 32; idx += 128) {
      const int r = idx >> 5, k = (idx & 31) * 2;
      uint32_t hi, lo;
      tc::split_h16x2(X[(size_t)r * K + s * 64 + k] * tc::ACT_SCALE_H16, X[(size_t)r * K + s * 64 + k + 1] * tc::ACT_SCALE_H16, hi, lo);
      *reinterpret_cast<uint32_t*>(sb + 32768 + tc::sw128_off_h(r, k)) = hi;
      *reinterpret_cast<uint32_t*>(sb + 40960 + tc::sw128_off_h(r, k)) = lo;
    }
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_f16(TCT_M, TCT_N);
    uint32_t accum = 0;
    for (int s = 0; s < nslab; ++s) {
      const uint32_t sb = tc::smem_u32(base + (size_t)s * SLAB);
      for (int j = 0; j < 4; ++j) {
        const uint32_t ko = j * 32;            // 16 fp16 = 32 bytes per k-step inside the swizzle atom
        const uint64_t whi = tc::desc_sw128(sb + ko), wlo = tc::desc_sw128(sb + 16384 + ko);
        const uint64_t xhi = tc::desc_sw128(sb + 32768 + ko), xlo = tc::desc_sw128(sb + 40960 + ko);
        tc::umma_f16(tm, wlo, xhi, idesc, accum); accum = 1;
        tc::umma_f16(tm, whi, xlo, idesc, accum);
        tc::umma_f16(tm, whi, xhi, idesc, accum);
      }
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const float unscale = 1.0f / (w_scale * tc::ACT_SCALE_H16);
  float v[32];
  for (int c = 0; c < TCT_N; c += 32) {
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[(size_t)tid * TCT_N + c + i] = v[i] * unscale;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 64);
}


// Probe of the tcgen05.ld shapes: every TMEM (lane, column) cell of a 128 x 32 block holds lane * 256 + column (written with the
// 32x32b shape whose mapping is known: thread = lane); every warp then reads lanes [32 w, 32 w + 16) with 16x256b.x2 / 16x128b.x2 /
// 16x64b.x2 and reports the cells its registers received: out[warp][lane][0..7] = 16x256b, [8..11] = 16x128b, [12..13] = 16x64b.
__global__ void __launch_bounds__(128, 1) k_tmem_shape_probe(int* __restrict__ out) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc(&tmem_base, 32);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t la = (uint32_t)(warp * 32) << 16;
  uint32_t v[16];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (uint32_t)((warp * 32 + lane) * 256 + h * 16 + i);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(tmem + la + h * 16),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  uint32_t a[8], b[4], c[2];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
               : "r"(tmem + la));
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(tmem + la));
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x2.b32 {%0, %1}, [%2];\n" : "=r"(c[0]), "=r"(c[1]) : "r"(tmem + la));
  tc::tmem_ld_wait();
  int* o = out + (warp * 32 + lane) * 16;
  for (int i = 0; i < 8; ++i) o[i] = (int)a[i];
  for (int i = 0; i < 4; ++i) o[8 + i] = (int)b[i];
  for (int i = 0; i < 2; ++i) o[12 + i] = (int)c[i];
  o[14] = o[15] = 0;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

}  // namespace fm
