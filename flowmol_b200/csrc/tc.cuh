// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; encodings follow cute/arch/mma_sm100_desc.hpp).
//
// Error-compensated TF32 ("3xTF32"): x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi);
//   x*w ~= hi_x*hi_w + lo_x*hi_w + hi_x*lo_w   (dropped lo*lo term ~2^-24 relative) accumulated in fp32 in TMEM,
// which keeps the fused edge kernels inside the fp32 parity budget while running on the 5th-gen tensor cores.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the hardware parks the thread until the phase completes or the time limit expires, so a
// waiting warp does not burn issue slots of the scheduler it shares with the warps it is waiting for (profiles/r01g: spin loops
// were ~40 % of all warp instructions of the persistent kernel)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap (and surface as a CUDA error), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (++spins > 400000u) __trap();          // >= seconds
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same copy delivered to the same shared-memory offset of every CTA in `cta_mask` of this cluster; each destination CTA's
// mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; the allocated base address is written to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 columns: thread (lane l of warp w) receives row 32*(w%4)+l, columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA -----------------------------------------------------------------------------------------------------------
// K-major operand tile, 128-byte swizzle: rows of 32 fp32 (128 B), 8-row atoms 1024 B apart (SBO), one atom along K.
// The tile base must be 1024-byte aligned; `byte_off` advances along K inside the atom (32 B per tf32 k-step of 8).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3FFF);          // start address  [0,14)
  d |= (uint64_t)0 << 16;                                      // LBO (unused: one swizzle atom along K)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;               // SBO = 1024 B  [32,46)
  d |= (uint64_t)1 << 46;                                      // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                                      // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor, kind::f16 with fp16 operands, fp32 accumulate, both operands K-major (a_format = b_format = 0: F16)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of the (converged) warp: the MMA-issuing warps stay converged so that descriptors / addresses live in uniform
// registers and the issue loop is a handful of instructions per MMA (an `if (lane == 0)` region forces per-MMA R2UR moves and a
// compiler-generated ELECT loop around every tcgen05 instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// SW128 K-major descriptor from its low word ((smem address >> 4) & 0x3FFF): the high word is the same for every operand tile
__device__ __forceinline__ uint64_t desc_sw128_lo(uint32_t lo) { return ((uint64_t)0x40004040u << 32) | lo; }
// completion of all previously issued MMAs of this thread -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ... -> one arrival on the mbarrier at this shared-memory offset in EVERY CTA of `cta_mask` (weight slots shared by a cluster)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- 3xTF32 split -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - hi);
}
// byte offset of element (row, k) inside a [rows][32] fp32 SW128 K-major slab
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 2) ^ (row & 7)) << 4) | ((k & 3) << 2)));
}


// ---- fp16x3 split ("scaled fp16 hi/lo") ----------------------------------------------------------------------------------------
// x * 2^s = hi + lo with hi = fp16_rn(x 2^s), lo = fp16_rn(x 2^s - hi): the same 22 significand bits as the TF32 split, in
// operands half as wide that the tensor cores consume at twice the rate.  Weights carry a per-matrix power-of-two scale chosen on
// the host so that max |w| 2^s lies in [2^13, 2^14): `lo` stays out of fp16's subnormal range for every weight that matters; the
// scale is undone exactly in the epilogue.  Activations are not scaled: their `lo` parts go subnormal only for |x| < 2^-3, an
// absolute error <= 2^-25 per element, below the fp32 accumulation noise of the products (simulated in tests/test_host_logic.py).
// |x| must stay below 65504 (checked by the loaders; reported through the status word).
constexpr float ACT_SCALE_H16 = 1.0f;      // activations: no scale needed (|x| < 2^-3 loses low bits at an absolute 2^-25, below fp32 noise)
constexpr float ACT_LIMIT_H16 = 65504.0f / ACT_SCALE_H16;
// two consecutive k values -> packed (hi, hi) and (lo, lo) half2 words; element k sits in the low half (lower address)
__device__ __forceinline__ void split_h16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi2 = *reinterpret_cast<const uint32_t*>(&h);
  lo2 = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of fp16 element (row, k), k in [0, 64), inside a [rows][64] fp16 SW128 K-major slab (same byte geometry as the
// [rows][32] fp32 slab: 128-byte rows, 16-byte chunks XOR-swizzled by row % 8)
__device__ __forceinline__ uint32_t sw128_off_h(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

}  // namespace tc
}  // namespace fm
