// Common device helpers for the FlowMol B200 sampling kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fm {

constexpr int NT = 256;     // threads per CTA for every tile kernel
constexpr int NWARP = 8;
constexpr int TM = 64;      // rows (edges / upper edges / nodes) per CTA tile
constexpr int RPW = 8;      // rows owned by one warp inside a tile
constexpr int KC = 16;      // K-chunk of the weight stream staged through shared memory
constexpr int WSTAGE_FLOATS = 2 * KC * 256;   // double buffer, widest weight matrix has 256 columns

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__host__ __device__ __forceinline__ int pad4(int x) { return (x + 3) & ~3; }
__host__ __device__ __forceinline__ int pad32(int x) { return (x + 31) & ~31; }

// ---- async copy (LDGSTS) -------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- programmatic dependent launch ---------------------------------------------------------------------------------------------------
// Every kernel of the pipeline is launched with cudaLaunchAttributeProgrammaticStreamSerialization (api.cu:launch_k): its CTAs may
// become resident while the kernel before it in the stream is still draining, run their prologue (barrier init, tensor-memory
// allocation, weight split) and then block in pdl_wait() until the predecessor has completed and its writes are visible.  Nothing
// that reads or writes activation data may come before pdl_wait(); pdl_launch() lets the NEXT kernel's CTAs start early in turn.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- math (accurate versions: the parity bar is the reference's fp32 noise floor) ---------------------------------
// 1 / v for v in [1, inf]: MUFU.RCP (1 ulp), branch-free.  The IEEE division / __frcp_rn compile to a call with a slow path
// behind a reconvergence barrier, which serialises unrolled epilogues into one ~130-cycle dependent chain per element
// (measured on k_egemm_tc: 34 k cycles for 256 elements per thread, profiles/r01c).
__device__ __forceinline__ float rcp_fast(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float silu_f(float x) { return x * rcp_fast(1.0f + expf(-x)); }   // torch.nn.SiLU: x / (1 + exp(-x))
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_fast(1.0f + expf(-x)); }    // torch.sigmoid
// tensor-core pipeline epilogues: ex2.approx based exponential as well (error ~1e-7 of the activation, below 3xTF32's)
// ex2.approx.ftz: without .ftz the instruction is wrapped in a range-scaling sequence (FSETP + 2 FMUL per element, ~15 % of the
// epilogues' instructions, ncu r01k); a flushed denormal exponential is absorbed by the "1 +" anyway, so the result is the same.
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  return rcp_fast(1.0f + e);
}

// flowmol/models/gvp.py:14-21 -- sqrt(clamp(x^2+y^2+z^2, 1e-8))
__device__ __forceinline__ float norm_no_nan3(float x, float y, float z) {
  float q = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  return sqrtf(fmaxf(q, 1e-8f));
}

// branch-free square root for q >= 1e-8 (MUFU.RSQ + one Newton step in fma arithmetic: within 1 ulp of sqrtf, whose IEEE routine
// carries a slow-path branch that breaks up unrolled loops)
__device__ __forceinline__ float sqrt_pos(float q) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
  const float s = q * r;
  return fmaf(0.5f * r, fmaf(-s, s, q), s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// v[i] of every lane summed over the 32 lanes, lane i receives the total of element i: 31 shuffles instead of 32 x 5
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32-10 (same function as oracle/philox.py) -------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}
__device__ __forceinline__ float u24(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }  // 2^-24

// ---- complete-graph index algebra ---------------------------------------------------------------------------------
// Internal directed-edge order inside a molecule with n atoms: local edge le = j*(n-1) + i', dst = j,
// src = i' + (i' >= j).  (The reference order -- upper triangle row-major, then its mirror,
// flowmol/data_processing/utils.py:4-17 -- is only used at the C-ABI boundary for the upper-edge arrays.)
__device__ __forceinline__ void edge_src_dst(int le, int n, int& i, int& j) {
  j = le / (n - 1);
  int ip = le - j * (n - 1);
  i = ip + (ip >= j ? 1 : 0);
}
__device__ __forceinline__ int edge_pos(int i, int j, int n) {  // local index of directed edge src i -> dst j
  return j * (n - 1) + (i < j ? i : i - 1);
}
// upper-triangle index lu (row-major over i<j) -> (i, j)
__device__ __forceinline__ void upper_ij(int lu, int n, int& i, int& j) {
  float b = (float)(2 * n - 1);
  int ii = (int)floorf((b - sqrtf(fmaxf(b * b - 8.0f * (float)lu, 0.0f))) * 0.5f);
  ii = max(0, min(ii, n - 2));
  // offset(i) = i*(2n-i-1)/2 ; fix rounding of the float sqrt
  while (ii > 0 && (ii * (2 * n - ii - 1)) / 2 > lu) --ii;
  while (((ii + 1) * (2 * n - ii - 2)) / 2 <= lu) ++ii;
  i = ii;
  j = lu - (ii * (2 * n - ii - 1)) / 2 + ii + 1;
}

}  // namespace fm
