// Block-level FP32 GEMM on a 64-row activation tile held in shared memory.
//
//   C[p][r][col] = sum_k A[p][r][k] * W[k][col]        p < PLANES, r < 64, col < 32*CPT
//
// * A lives in shared memory (row-major, leading dimension lda, planes plane_stride apart; PLANES = 3 is the
//   x/y/z view of GVP vector features).  Warp w owns rows [8w, 8w+8) of every plane, so every A read is a
//   warp-wide broadcast (one wavefront).
// * W lives in global memory / L2 as [K][32*CPT] row-major (zero padded by the host packer) and is streamed through a
//   double-buffered shared-memory stage with cp.async (LDGSTS) in chunks of KC = 16 rows; every lane reads its own
//   columns with 128-bit conflict-free loads.
// * Each thread keeps an 8 x CPT accumulator micro-tile per plane in registers: 64 FFMA per (4 + 2) shared loads
//   for the 256-wide layers -- the fp32 CUDA-core path that reproduces the reference's fp32 arithmetic.
//
// Every thread of the CTA must call it (it contains __syncthreads()).  K must be a multiple of 4 and the A columns
// [0, K) must hold finite values (pad columns are zero-filled by the callers).
#pragma once
#include "common.cuh"

namespace fm {

template <int CPT>
struct ColMap {
  static constexpr int VEC = CPT >= 4 ? 4 : CPT;
  static constexpr int NP = 32 * CPT;
  __device__ static __forceinline__ int col(int lane, int c) { return (c / VEC) * (32 * VEC) + lane * VEC + (c % VEC); }
};

template <int CPT>
__device__ __forceinline__ void load_w(const float* __restrict__ wrow, int lane, float (&w)[CPT]) {
  if constexpr (CPT >= 4) {
#pragma unroll
    for (int g = 0; g < CPT / 4; ++g) {
      float4 t = *reinterpret_cast<const float4*>(wrow + g * 128 + lane * 4);
      w[g * 4 + 0] = t.x; w[g * 4 + 1] = t.y; w[g * 4 + 2] = t.z; w[g * 4 + 3] = t.w;
    }
  } else if constexpr (CPT == 2) {
    float2 t = *reinterpret_cast<const float2*>(wrow + lane * 2);
    w[0] = t.x; w[1] = t.y;
  } else {
    w[0] = wrow[lane];
  }
}

template <int PLANES, int CPT, int RPW = 8, int WCAP = WSTAGE_FLOATS>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int lda, int plane_stride, int K,
                                          const float* __restrict__ Wg, float* __restrict__ wstage,
                                          float (&acc)[PLANES][RPW][CPT]) {
  constexpr int NP = 32 * CPT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int p = 0; p < PLANES; ++p)
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[p][r][c] = 0.0f;

  if (K <= 64 && K * NP <= WCAP) {
    // small operand (the vector-channel GEMMs): stage the whole matrix once -- a chunked pipeline would expose one L2
    // round trip per 16 rows because there is almost no math to hide it behind
    __syncthreads();
    for (int i = tid; i < K * NP / 4; i += NT) cp_async16(wstage + i * 4, Wg + i * 4);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int kk = 0; kk < K; kk += 4) {
#pragma unroll
      for (int p = 0; p < PLANES; ++p) {
        float4 a[RPW];
        const float* ap = A + p * plane_stride + (warp * RPW) * lda + kk;
#pragma unroll
        for (int r = 0; r < RPW; ++r) a[r] = *reinterpret_cast<const float4*>(ap + r * lda);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w[CPT];
          load_w<CPT>(wstage + (kk + j) * NP, lane, w);
#pragma unroll
          for (int r = 0; r < RPW; ++r) {
            const float av = j == 0 ? a[r].x : (j == 1 ? a[r].y : (j == 2 ? a[r].z : a[r].w));
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) acc[p][r][cc] = fmaf(av, w[cc], acc[p][r][cc]);
          }
        }
      }
    }
    __syncthreads();
    return;
  }
  const int nchunks = (K + KC - 1) / KC;
  auto issue = [&](int c) {
    const int k0 = c * KC;
    const int rows = min(KC, K - k0);
    const int nvec = rows * NP / 4;
    float* dst = wstage + (c & 1) * (KC * NP);
    const float* src = Wg + (size_t)k0 * NP;
    for (int i = tid; i < nvec; i += NT) cp_async16(dst + i * 4, src + i * 4);
    cp_async_commit();
  };

  __syncthreads();   // A tile complete; previous users of wstage done
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      issue(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int k0 = c * KC;
    const int rows = min(KC, K - k0);
    const float* wb = wstage + (c & 1) * (KC * NP);
    for (int kk = 0; kk < rows; kk += 4) {
#pragma unroll
      for (int p = 0; p < PLANES; ++p) {
        float4 a[RPW];
        const float* ap = A + p * plane_stride + (warp * RPW) * lda + k0 + kk;
#pragma unroll
        for (int r = 0; r < RPW; ++r) a[r] = *reinterpret_cast<const float4*>(ap + r * lda);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w[CPT];
          load_w<CPT>(wb + (kk + j) * NP, lane, w);
#pragma unroll
          for (int r = 0; r < RPW; ++r) {
            const float av = j == 0 ? a[r].x : (j == 1 ? a[r].y : (j == 2 ? a[r].z : a[r].w));
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) acc[p][r][cc] = fmaf(av, w[cc], acc[p][r][cc]);
          }
        }
      }
    }
    __syncthreads();
  }
}

// Same contraction with the (small) weight matrix already resident in shared memory: no staging, one barrier on entry
// (A complete) and one on exit.  Used by the persistent vector-stage kernels, which load their two matrices once per CTA.
template <int PLANES, int CPT, int RPW = 8>
__device__ __forceinline__ void tile_gemm_resident(const float* __restrict__ A, int lda, int plane_stride, int K,
                                                   const float* __restrict__ Wsm, float (&acc)[PLANES][RPW][CPT]) {
  constexpr int NP = 32 * CPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < PLANES; ++p)
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[p][r][c] = 0.0f;
  __syncthreads();
  for (int kk = 0; kk < K; kk += 4) {
#pragma unroll
    for (int p = 0; p < PLANES; ++p) {
      float4 a[RPW];
      const float* ap = A + p * plane_stride + (warp * RPW) * lda + kk;
#pragma unroll
      for (int r = 0; r < RPW; ++r) a[r] = *reinterpret_cast<const float4*>(ap + r * lda);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float w[CPT];
        load_w<CPT>(Wsm + (kk + j) * NP, lane, w);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const float av = j == 0 ? a[r].x : (j == 1 ? a[r].y : (j == 2 ? a[r].z : a[r].w));
#pragma unroll
          for (int cc = 0; cc < CPT; ++cc) acc[p][r][cc] = fmaf(av, w[cc], acc[p][r][cc]);
        }
      }
    }
  }
  __syncthreads();
}

// LayerNorm (eps 1e-5, affine) of rows held as an [8][CPT] register micro-tile: the 32 lanes of a warp hold the
// 32*CPT columns of the same 8 rows.  torch.nn.LayerNorm: biased variance.
template <int CPT>
__device__ __forceinline__ void rows_layernorm(float (&v)[RPW][CPT], const float* __restrict__ gamma,
                                               const float* __restrict__ beta) {
  constexpr float inv_n = 1.0f / (32 * CPT);
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) s += v[r][c];
    const float mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) { const float d = v[r][c] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * inv_n + 1e-5f);
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = ColMap<CPT>::col(lane, c);
      v[r][c] = (v[r][c] - mean) * rstd * gamma[col] + beta[col];
    }
  }
}

}  // namespace fm
