// Kernels of one vector-field evaluation (EndpointVectorField.forward / denoise_graph,
// flowmol/models/vector_field.py:212-369) on batched complete molecular graphs.
//
// Kernel                what it replaces in the reference
//   k_node_embed        token/time embedding MLP + LN (vector_field.py:227-244) + self-conditioning node residual
//                       (self_conditioning.py:46-57,76) + per-node pre-activations of conv 0's first message linear
//   k_edge_init         edge token embedding (vector_field.py:257-261, a 5-row table) + self-conditioning edge residual
//                       on upper edges mirrored to both directions (self_conditioning.py:59-82)
//   k_conv_edge   (HOT) GVPConv edge phase: gather s/v/x of src (and dst), recompute x_diff / rbf(d), 3 message GVPs,
//                       segment-sum over in-edges of every dst (gvp.py:476,491-492,523-543; vector_field.py:371-386)
//   k_node_update       residual + GVPLayerNorm, 3 update GVPs, residual + GVPLayerNorm (gvp.py:509-519), then the per-node
//                       halves of the next edge phases, then NodePositionUpdate (vector_field.py:813-842)
//   k_dst_proj          dst_feat_msg_projection GVP (gvp.py:472-473), dev config only
//   k_edge_update       EdgeUpdate (vector_field.py:844-880) with distances recomputed from the new positions
//   k_node_head / k_edge_head / k_com   output heads, softmax, COM removal (vector_field.py:336-367)
#pragma once
#include "gvp.cuh"
#include "mma3.cuh"

namespace fm {

// ------------------------------------------------------------------------------------------------------------------
// shared-memory carve-up (same plan for every tile kernel)
// ------------------------------------------------------------------------------------------------------------------
template <class D>
struct Smem {
  float *Xs, *Va, *Vb, *G, *wstage;
  int *src, *dst;
  float* dist;
  int* aux;
  __device__ explicit Smem(float* base) {
    Xs = base;
    Va = Xs + D::SM_XS;
    Vb = Va + D::SM_VA;
    G = Vb + D::SM_VB;
    wstage = G + D::SM_G;
    float* misc = wstage + WSTAGE_FLOATS;
    src = reinterpret_cast<int*>(misc);
    dst = src + TM;
    dist = misc + 2 * TM;
    aux = reinterpret_cast<int*>(misc + 3 * TM);
  }
};

// compact plan of the two upper-edge kernels (k_edge_init, k_edge_head): one activation tile [64][2F + R + 12] and a weight stage
// for F-wide matrices -- 94 KB at flowmol3 instead of the generic 183 KB, i.e. 2 CTAs per SM (these kernels are chains of
// gather -> GEMM -> GEMM -> store with nothing to overlap inside one CTA)
template <class D>
struct EdgeSmem {
  static constexpr int XE = 2 * D::F + D::R + 12;
  static constexpr int WST = 2 * KC * D::F;
  static constexpr int FLOATS = TM * XE + TM + WST + D::SM_MISC;
  static constexpr size_t BYTES = (size_t)FLOATS * 4;
  __device__ static Smem<D> carve(float* base) {
    Smem<D> sm(base);
    sm.Xs = base;
    sm.Va = nullptr; sm.Vb = nullptr;
    sm.G = base + TM * XE;             // [64] scratch
    sm.wstage = sm.G + TM;
    float* misc = sm.wstage + WST;
    sm.src = reinterpret_cast<int*>(misc);
    sm.dst = sm.src + TM;
    sm.dist = misc + 2 * TM;
    sm.aux = reinterpret_cast<int*>(misc + 3 * TM);
    return sm;
  }
};

// rbf(d)_k = exp(-((d - mu_k)/sigma)^2), sigma = dmax / R        (flowmol/utils/embedding.py:19-34)
__device__ __forceinline__ float rbf_f(float d, float mu, float sigma) {
  const float z = __fdiv_rn(__fsub_rn(d, mu), sigma);
  return expf(-__fmul_rn(z, z));
}

// rbf for the operand loaders of the tensor-core linears: the value is split into fp16 (hi, lo) / TF32 operands (22 significand
// bits) right after, so the IEEE division and the accurate expf of rbf_f (~35 instructions with slow-path branches, 32 of them
// per edge: ~17 % of k_egemm_p<MSG0>'s samples, profiles/r01s) buy nothing there.  Relative error ~1e-7.
__device__ __forceinline__ float rbf_fast(float d, float mu, float inv_sigma) {
  const float z = (d - mu) * inv_sigma;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  return e;
}

// distance of an ordered node pair: ||x_a - x_b|| clamped + 1e-8   (vector_field.py:381-382)
__device__ __forceinline__ float pair_dist(const float* __restrict__ x, int a, int b, float& dx, float& dy, float& dz) {
  dx = __fsub_rn(x[a * 3 + 0], x[b * 3 + 0]);
  dy = __fsub_rn(x[a * 3 + 1], x[b * 3 + 1]);
  dz = __fsub_rn(x[a * 3 + 2], x[b * 3 + 2]);
  return __fadd_rn(norm_no_nan3(dx, dy, dz), 1e-8f);
}

// ------------------------------------------------------------------------------------------------------------------
// tile_gemm_h16: the 64-row tile GEMM of tile_gemm.cuh, C = Xs[64][K] W[K][256], on mma.sync.m16n8k16 with the error-compensated
// fp16 (hi, lo) operands of mma3.cuh (a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulate) -- for k_node_embed, whose five
// 256-wide linears were 0.62 ms of fp32 FFMA per evaluation (profiles/r02_kprof_last.txt).  The weights stay fp32 in global
// memory: every CTA converts them on the fly, 32 k rows per chunk, into packed (k, k + 1) half2 words (double-buffered: the next
// chunk's global loads are in flight during this chunk's MMAs).  Fixed power-of-two weight scale 2^10: lo parts are exact to
// 2^-35 absolute, |w| must stay below 64 (fm_create checks the five matrices and keeps the fp32 path otherwise).
// Warp w owns output columns [32 w, 32 w + 32) of all 64 rows (4 x 4 accumulator tiles); the result goes through shared memory
// (Ys) into the caller's row-per-warp register layout (ColMap), so the epilogues are those of the fp32 path.
// ------------------------------------------------------------------------------------------------------------------
constexpr int NE_LDW = 264;                 // words per k-pair row of the converted weights (256 + 8: conflict-free B fragments)
constexpr int NE_LDY = 264;                 // floats per row of the result tile
constexpr int NE_WBUF = 2 * 16 * NE_LDW;    // one chunk: 16 k-pair rows of hi words | 16 of lo words
constexpr float NE_WSCALE = 1024.0f;
constexpr float NE_WMAX = 32.0f;            // fm_create: largest |w| the mma path accepts

// not volatile: a pure function of its operands, ptxas may interleave it with the splits and loads around it
__device__ __forceinline__ void ne_mma16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// amax: running max |a| over the activation values this thread split into fp16 (hi, lo) -- the caller reports >= 65504 through the
// status word like every other fp16x3 kernel (an overflowing hi part would turn into inf - inf = NaN inside the products)
template <int S>
__device__ __forceinline__ void tile_gemm_h16(const float* __restrict__ Xs, int lda, int K, const float* __restrict__ Wg,
                                              uint32_t* __restrict__ wbuf, float* __restrict__ Ys, float (&acc)[1][RPW][S / 32],
                                              float& amax) {
  static_assert(S == 256 && NT == 256 && TM == 64, "8 warps x 32 columns, 4 m16 tiles");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int n0 = warp * 32;
  float c[4][4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
  // the next chunk's weights are fetched in two halves (k pairs 0-7 during k-step 0, 8-15 during k-step 1): 16 registers in flight
  float4 wr[2][2];
  auto wload = [&](int ch, int half) {       // item = (k pair, 4 columns): rows k, k + 1 of the chunk
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int item = tid + NT * (2 * half + j), kp = item >> 6, n4 = item & 63, k = ch * 32 + 2 * kp;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      wr[j][0] = k < K ? __ldg(reinterpret_cast<const float4*>(Wg + (size_t)k * S) + n4) : z;
      wr[j][1] = k + 1 < K ? __ldg(reinterpret_cast<const float4*>(Wg + (size_t)(k + 1) * S) + n4) : z;
    }
  };
  auto wstore = [&](int buf, int half) {
    uint32_t* hi = wbuf + buf * NE_WBUF;
    uint32_t* lo = hi + 16 * NE_LDW;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int item = tid + NT * (2 * half + j), kp = item >> 6, n4 = item & 63;
      uint4 h4, l4;
      tc::split_h16x2(wr[j][0].x * NE_WSCALE, wr[j][1].x * NE_WSCALE, h4.x, l4.x);
      tc::split_h16x2(wr[j][0].y * NE_WSCALE, wr[j][1].y * NE_WSCALE, h4.y, l4.y);
      tc::split_h16x2(wr[j][0].z * NE_WSCALE, wr[j][1].z * NE_WSCALE, h4.z, l4.z);
      tc::split_h16x2(wr[j][0].w * NE_WSCALE, wr[j][1].w * NE_WSCALE, h4.w, l4.w);
      *reinterpret_cast<uint4*>(hi + kp * NE_LDW + 4 * n4) = h4;
      *reinterpret_cast<uint4*>(lo + kp * NE_LDW + 4 * n4) = l4;
    }
  };
  const int nch = (K + 31) / 32;
  __syncthreads();                           // the A tile is complete; Ys / wbuf of a previous call have been read
  wload(0, 0); wstore(0, 0);
  wload(0, 1); wstore(0, 1);
  __syncthreads();
  for (int ch = 0; ch < nch; ++ch) {
    const uint32_t* hi = wbuf + (ch & 1) * NE_WBUF;
    const uint32_t* lo = hi + 16 * NE_LDW;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int k0 = ch * 32 + 16 * ks;
      if (ch + 1 < nch) wload(ch + 1, ks);
      if (k0 < K) {
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int w = (8 * ks + t + 4 * i) * NE_LDW + n0 + 8 * nt + g;
            bh[nt][i] = hi[w]; bl[nt][i] = lo[w];
          }
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {      // a0 (g, 2t)  a1 (g + 8, 2t)  a2 (g, 2t + 8)  a3 (g + 8, 2t + 8)
            const int col = k0 + 2 * t + (i >> 1) * 8;
            float2 v = make_float2(0.f, 0.f);
            if (col < K) v = *reinterpret_cast<const float2*>(Xs + (16 * mt + g + (i & 1) * 8) * lda + col);
            amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
            tc::split_h16x2(v.x, v.y, ah[mt][i], al[mt][i]);
          }
        // the three products of an accumulator are 16 instructions apart (same order per accumulator as mma3.cuh: a_lo b_hi, a_hi b_lo, a_hi b_hi)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) ne_mma16(c[mt][nt], al[mt], bh[nt]);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) ne_mma16(c[mt][nt], ah[mt], bl[nt]);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) ne_mma16(c[mt][nt], ah[mt], bh[nt]);
      }
      if (ch + 1 < nch) wstore((ch + 1) & 1, ks);
    }
    __syncthreads();
  }
  constexpr float inv = 1.0f / NE_WSCALE;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* yp = Ys + (16 * mt + g) * NE_LDY + n0 + 8 * nt + 2 * t;
      *reinterpret_cast<float2*>(yp) = make_float2(c[mt][nt][0] * inv, c[mt][nt][1] * inv);
      *reinterpret_cast<float2*>(yp + 8 * NE_LDY) = make_float2(c[mt][nt][2] * inv, c[mt][nt][3] * inv);
    }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const float* yr = Ys + (warp * RPW + r) * NE_LDY + lane * 4;
    const float4 y0 = *reinterpret_cast<const float4*>(yr), y1 = *reinterpret_cast<const float4*>(yr + 128);
    acc[0][r][0] = y0.x; acc[0][r][1] = y0.y; acc[0][r][2] = y0.z; acc[0][r][3] = y0.w;
    acc[0][r][4] = y1.x; acc[0][r][5] = y1.y; acc[0][r][6] = y1.z; acc[0][r][7] = y1.w;
  }
}

template <class D>
struct NodeEmbedMmaSmem {                    // Xs | Ys | converted weight chunks (the fp32 path's vector planes / weight stage are unused)
  static constexpr size_t BYTES = ((size_t)D::SM_XS + (size_t)TM * NE_LDY + 2 * NE_WBUF) * 4;
  static_assert(BYTES <= 232448, "227 KB of shared memory per CTA");
};

// ------------------------------------------------------------------------------------------------------------------
// k_node_embed
// ------------------------------------------------------------------------------------------------------------------
struct PredPtr {            // predicted endpoint ("dst_dict"): x [N,3], a [N,A], c [N,C], e [U,EB]   (device pointers)
  float *x, *a, *c, *e;
};

// MMA = 1 (option node_embed_tc, S = 256): the five linears on mma.sync fp16x3 (tile_gemm_h16) instead of fp32 FFMA
template <class D, int MMA = 0>
__global__ void __launch_bounds__(NT, 1)
k_node_embed(const ModelRT m, const BatchRT bt, const float* __restrict__ x_t, const uint8_t* __restrict__ a_t,
             const uint8_t* __restrict__ c_t, float t, const PredPtr prev, int has_prev,
             float* __restrict__ s_out, float* __restrict__ v_out, float* __restrict__ P0, int* __restrict__ status) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float amax = 0.f;
  auto gemm = [&](const float* __restrict__ W, int K, float (&out)[1][RPW][D::CPT_S]) {
    if constexpr (MMA) {
      float* Ys = smem_raw + D::SM_XS;
      tile_gemm_h16<D::S>(sm.Xs, D::XLD, K, W, reinterpret_cast<uint32_t*>(Ys + TM * NE_LDY), Ys, out, amax);
    } else {
      tile_gemm<1, D::CPT_S>(sm.Xs, D::XLD, 0, K, W, sm.wstage, out);
    }
  };
  const int g0 = blockIdx.x * TM;
  // [Emb_a | Emb_c | time embedding]                                             (vector_field.py:232-243)
  const float* ea = m.g(G_EMB_A);
  const float* ec = m.g(G_EMB_C);
  const float* freq = m.g(G_TIME_FREQ);
  constexpr int K0 = 2 * D::TOK + D::TD;
  for (int idx = tid; idx < TM * K0; idx += NT) {
    const int row = idx / K0, c = idx - row * K0, g = g0 + row;
    float v = 0.f;
    if (g < bt.N) {
      if (c < D::TOK) v = ea[(int)a_t[g] * D::TOK + c];
      else if (c < 2 * D::TOK) v = ec[(int)c_t[g] * D::TOK + (c - D::TOK)];
      else {
        const int k = c - 2 * D::TOK;
        const float arg = __fmul_rn(__fmul_rn(t, 1000.0f), freq[k < D::TD / 2 ? k : k - D::TD / 2]);   // embedding.py:8-13
        v = k < D::TD / 2 ? sinf(arg) : cosf(arg);
      }
    }
    sm.Xs[row * D::XLD + c] = v;
  }
  float acc[1][RPW][D::CPT_S];
  gemm(m.g(G_SEMB0_W), K0, acc);
  {
    const float* b = m.g(G_SEMB0_B);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < D::CPT_S; ++c) {
        const int col = ColMap<D::CPT_S>::col(lane, c);
        sm.Xs[(warp * RPW + r) * D::XLD + col] = silu_f(acc[0][r][c] + b[col]);
      }
  }
  gemm(m.g(G_SEMB2_W), D::S, acc);
  {
    const float* b = m.g(G_SEMB2_B);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < D::CPT_S; ++c) acc[0][r][c] = silu_f(acc[0][r][c] + b[ColMap<D::CPT_S>::col(lane, c)]);
    rows_layernorm<D::CPT_S>(acc[0], m.g(G_SEMB_LN_W), m.g(G_SEMB_LN_B));
  }
  // self-conditioning node residual: s += MLP(cat[s, a_hat, c_hat, rbf(||x_t - x_hat||)])      (self_conditioning.py:46-57,76)
  if (has_prev) {
    const int A = m.A, C = m.C;
    const int K = D::S + A + C + D::R, KP = pad4(K);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < D::CPT_S; ++c) {
        const int col = ColMap<D::CPT_S>::col(lane, c);
        sm.Xs[(warp * RPW + r) * D::XLD + col] = acc[0][r][c];
        // MMA: the 64 values of s are parked in s_out across the two linears (the mma path needs the registers: with them live the
        // prefetched weight chunk was spilled, i.e. waited for, ahead of every MMA phase)
        if (MMA && g0 + warp * RPW + r < bt.N) s_out[(size_t)(g0 + warp * RPW + r) * D::S + col] = acc[0][r][c];
      }
    const float* mu = m.g(G_RBF_MU);
    const float sigma = m.rbf_dmax / (float)D::R;
    const int extra = KP - D::S;
    for (int idx = tid; idx < TM * extra; idx += NT) {
      const int row = idx / extra, c = idx - row * extra, g = g0 + row;
      float v = 0.f;
      if (g < bt.N) {
        if (c < A) v = prev.a[(size_t)g * A + c];
        else if (c < A + C) v = prev.c[(size_t)g * C + (c - A)];
        else if (c < A + C + D::R) {
          const float d = norm_no_nan3(__fsub_rn(x_t[g * 3 + 0], prev.x[g * 3 + 0]), __fsub_rn(x_t[g * 3 + 1], prev.x[g * 3 + 1]),
                                       __fsub_rn(x_t[g * 3 + 2], prev.x[g * 3 + 2]));
          v = rbf_f(d, mu[c - A - C], sigma);
        }
      }
      sm.Xs[row * D::XLD + D::S + c] = v;
    }
    float acc2[1][RPW][D::CPT_S];
    gemm(m.g(G_SCN0_W), KP, acc2);
    {
      const float* b = m.g(G_SCN0_B);
#pragma unroll
      for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int c = 0; c < D::CPT_S; ++c) {
          const int col = ColMap<D::CPT_S>::col(lane, c);
          sm.Xs[(warp * RPW + r) * D::XLD + col] = silu_f(acc2[0][r][c] + b[col]);
        }
    }
    gemm(m.g(G_SCN2_W), D::S, acc2);
    {
      const float* b = m.g(G_SCN2_B);
#pragma unroll
      for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int c = 0; c < D::CPT_S; ++c)
        {
          const int col = ColMap<D::CPT_S>::col(lane, c);
          float sv = acc[0][r][c];
          if constexpr (MMA) sv = g0 + warp * RPW + r < bt.N ? s_out[(size_t)(g0 + warp * RPW + r) * D::S + col] : 0.f;
          acc[0][r][c] = __fadd_rn(sv, silu_f(acc2[0][r][c] + b[col]));
        }
    }
  }
  // store s, zero v (vector_field.py:251), and the per-node half of conv 0's first message linear
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r, g = g0 + row;
#pragma unroll
    for (int c = 0; c < D::CPT_S; ++c) {
      const int col = ColMap<D::CPT_S>::col(lane, c);
      sm.Xs[row * D::XLD + col] = acc[0][r][c];
      if (g < bt.N) s_out[(size_t)g * D::S + col] = acc[0][r][c];
    }
  }
  for (int idx = tid; idx < TM * 3 * D::V; idx += NT) {
    const int row = idx / (3 * D::V), g = g0 + row;
    if (g < bt.N) v_out[(size_t)g * 3 * D::V + (idx - row * 3 * D::V)] = 0.f;
  }
  gemm(m.c(0, C_WSRC), D::S, acc);
  {
    const float* b = m.c(0, C_BSRC);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int g = g0 + warp * RPW + r;
      if (g < bt.N)
#pragma unroll
        for (int c = 0; c < D::CPT_S; ++c) {
          const int col = ColMap<D::CPT_S>::col(lane, c);
          P0[(size_t)g * D::S + col] = acc[0][r][c] + b[col];
        }
    }
  }
  if (MMA && !(amax < tc::ACT_LIMIT_H16) && status) atomicOr(status, 1);
}

// ------------------------------------------------------------------------------------------------------------------
// k_edge_table: edge_embedding applied to the (n_bond_types + 1) possible tokens -- a weights-only constant computed
// once at fm_create (vector_field.py:257-261: every directed edge carries one of 5 token embeddings).
// ------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void k_edge_table(const ModelRT m, float* __restrict__ table) {
  pdl_launch();
  pdl_wait();
  __shared__ float h0[D::F], h1[D::F];
  const int tok = blockIdx.x, tid = threadIdx.x;
  const float* emb = m.g(G_EMB_E) + tok * D::TOK;
  const float* w0 = m.g(G_EEMB0_W);
  const float* w2 = m.g(G_EEMB2_W);
  constexpr int NP = D::CPT_F * 32;
  if (tid < D::F) {
    float a = 0.f;
    for (int k = 0; k < D::TOK; ++k) a = fmaf(emb[k], w0[k * NP + tid], a);
    h0[tid] = silu_f(a + m.g(G_EEMB0_B)[tid]);
  }
  __syncthreads();
  if (tid < D::F) {
    float a = 0.f;
    for (int k = 0; k < D::F; ++k) a = fmaf(h0[k], w2[k * NP + tid], a);
    h1[tid] = silu_f(a + m.g(G_EEMB2_B)[tid]);
  }
  __syncthreads();
  if (tid < D::F) {
    float mean = 0.f;
    for (int k = 0; k < D::F; ++k) mean += h1[k];
    mean /= D::F;
    float var = 0.f;
    for (int k = 0; k < D::F; ++k) var += (h1[k] - mean) * (h1[k] - mean);
    var /= D::F;
    table[tok * D::F + tid] = (h1[tid] - mean) / sqrtf(var + 1e-5f) * m.g(G_EEMB_LN_W)[tid] + m.g(G_EEMB_LN_B)[tid];
  }
  // Second table (self-conditioning models): the part of the edge residual MLP's first linear that multiplies the embedded edge
  // features (self_conditioning.py:70-78: cat[e_feats, e_1_pred, rbf differences]) takes one of n_bond_types + 1 values as well:
  // T1[tok] = b1 + table[tok] . W1[0:F, :].  k_edge_init then contracts only the F + 4 + R -> 4 + R per-edge columns (-44 % of its
  // MACs; a different fp32 summation order for those F terms, inside the parity budget like the node-side folds).
  if (m.self_cond) {
    __syncthreads();
    if (tid < D::F) h0[tid] = table[tok * D::F + tid];
    __syncthreads();
    if (tid < D::F) {
      const float* w1 = m.g(G_SCE0_W);
      float a = 0.f;
      for (int k = 0; k < D::F; ++k) a = fmaf(h0[k], w1[k * NP + tid], a);
      table[(m.EB + 1 + tok) * D::F + tid] = a + m.g(G_SCE0_B)[tid];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_edge_init: one tile = 64 upper edges (i<j) of one molecule
// ------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_edge_init(const ModelRT m, const BatchRT bt, const float* __restrict__ x_t, const uint8_t* __restrict__ e_t,
            const PredPtr prev, int has_prev, float* __restrict__ ef) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = EdgeSmem<D>::carve(smem_raw);
  constexpr int XE = EdgeSmem<D>::XE, WST = EdgeSmem<D>::WST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, mol = bt.utile_mol[tile];
  const int n = bt.mol_n[mol], nb = bt.mol_node[mol], ucount = n * (n - 1) / 2;
  const int lu0 = (tile - bt.mol_utile[mol]) * TM;
  const long long ebase = (long long)bt.mol_etile[mol] * TM;
  const int ub = bt.mol_u[mol];
  // per-row: positions of the two directed edges in the internal order; sm.src = pos(i->j), sm.dst = pos(j->i)
  if (tid < TM) {
    const int lu = lu0 + tid;
    int p0 = -1, p1 = -1, tok = 0;
    float dd = 0.f, d1 = 0.f;
    if (lu < ucount) {
      int i, j;
      upper_ij(lu, n, i, j);
      p0 = edge_pos(i, j, n);
      p1 = edge_pos(j, i, n);
      tok = e_t[ub + lu];
      if (has_prev) {
        float dx, dy, dz;
        dd = pair_dist(x_t, nb + i, nb + j, dx, dy, dz);           // self_conditioning.py:88-103 (edge_distances)
        d1 = pair_dist(prev.x, nb + i, nb + j, dx, dy, dz);
      }
    }
    sm.src[tid] = p0; sm.dst[tid] = p1; sm.aux[tid] = tok; sm.dist[tid] = dd; sm.G[tid] = d1;
  }
  __syncthreads();
  // self-conditioning residual on the edge features (self_conditioning.py:70-82).  The embedded edge features are one of
  // n_bond_types + 1 table rows, so their share of the first linear is the precomputed T1[tok] (k_edge_table); the per-edge
  // operand is only [e_1_pred (EB) | rbf(d_1) - rbf(d_t) (R)].
  const int EB = m.EB;
  const int K = EB + D::R, KP = pad4(K);
  const float* tab = m.eemb_table;
  const float* tab1 = m.eemb_table + (size_t)(EB + 1) * D::F;
  float out[1][RPW][D::CPT_F];
  if (has_prev) {
    for (int idx = tid; idx < TM * KP; idx += NT) {
      const int row = idx / KP, c = idx - row * KP;
      float v = 0.f;
      if (sm.src[row] >= 0) {
        if (c < EB) v = prev.e[(size_t)(ub + lu0 + row) * EB + c];
        else if (c < K) {
          const float mu = m.g(G_RBF_MU)[c - EB], sigma = m.rbf_dmax / (float)D::R;
          v = __fsub_rn(rbf_f(sm.G[row], mu, sigma), rbf_f(sm.dist[row], mu, sigma));   // d_edge_1 - d_edge_t
        }
      }
      sm.Xs[row * XE + c] = v;
    }
    float acc[1][RPW][D::CPT_F];
    tile_gemm<1, D::CPT_F, RPW, WST>(sm.Xs, XE, 0, KP, m.g(G_SCE0_W) + (size_t)D::F * (D::CPT_F * 32), sm.wstage, acc);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int tok = sm.aux[warp * RPW + r];
#pragma unroll
      for (int c = 0; c < D::CPT_F; ++c) {
        const int col = ColMap<D::CPT_F>::col(lane, c);
        sm.Xs[(warp * RPW + r) * XE + KP + col] = silu_f(acc[0][r][c] + tab1[tok * D::F + col]);
      }
    }
    tile_gemm<1, D::CPT_F, RPW, WST>(sm.Xs + KP, XE, 0, D::F, m.g(G_SCE2_W), sm.wstage, acc);
    const float* b = m.g(G_SCE2_B);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int tok = sm.aux[warp * RPW + r];
#pragma unroll
      for (int c = 0; c < D::CPT_F; ++c) {
        const int col = ColMap<D::CPT_F>::col(lane, c);
        out[0][r][c] = __fadd_rn(tab[tok * D::F + col], silu_f(acc[0][r][c] + b[col]));
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int tok = sm.aux[warp * RPW + r];
#pragma unroll
      for (int c = 0; c < D::CPT_F; ++c) out[0][r][c] = tab[tok * D::F + ColMap<D::CPT_F>::col(lane, c)];
    }
  }
  // mirrored store (self_conditioning.py:79-82)
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r;
    const int p0 = sm.src[row], p1 = sm.dst[row];
    if (p0 >= 0) {
      if constexpr (D::CPT_F >= 4) {
#pragma unroll
        for (int g = 0; g < D::CPT_F / 4; ++g) {
          const float4 v = make_float4(out[0][r][g * 4], out[0][r][g * 4 + 1], out[0][r][g * 4 + 2], out[0][r][g * 4 + 3]);
          const int col = g * 128 + lane * 4;
          *reinterpret_cast<float4*>(ef + (size_t)(ebase + p0) * D::F + col) = v;
          *reinterpret_cast<float4*>(ef + (size_t)(ebase + p1) * D::F + col) = v;
        }
      } else {
#pragma unroll
        for (int c = 0; c < D::CPT_F; ++c) {
          const int col = ColMap<D::CPT_F>::col(lane, c);
          ef[(size_t)(ebase + p0) * D::F + col] = out[0][r][c];
          ef[(size_t)(ebase + p1) * D::F + col] = out[0][r][c];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_conv_edge (the hot kernel): one tile = 64 consecutive in-edges (dst-major) of one molecule
// ------------------------------------------------------------------------------------------------------------------
template <class D>
struct Msg0Pre {           // per-node pre-activations of message GVP 0 gathered per edge: P[src] (+ Q[dst])
  const float* P;
  const float* Q;
  const int* src;
  const int* dst;
  __device__ __forceinline__ float operator()(int row, int col) const {
    const int s = src[row];
    if (s < 0) return 0.f;
    float v = P[(size_t)s * D::S + col];
    if constexpr (D::SD > 0) v = __fadd_rn(v, Q[(size_t)dst[row] * D::S + col]);
    return v;
  }
};

template <class D>
__global__ void __launch_bounds__(NT, 1)
k_conv_edge(const ModelRT m, const BatchRT bt, int layer, const float* __restrict__ x, const float* __restrict__ v,
            const float* __restrict__ ef, const float* __restrict__ P, const float* __restrict__ Q,
            const float* __restrict__ vd, float* __restrict__ M, float* __restrict__ partF, float* __restrict__ partL) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm(smem_raw);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, mol = bt.etile_mol[tile];
  const int n = bt.mol_n[mol], nb = bt.mol_node[mol], ecount = n * (n - 1);
  const int le0 = (tile - bt.mol_etile[mol]) * TM;
  const size_t erow0 = (size_t)tile * TM;              // first padded edge slot of this tile
  // ---- gather + geometric features -----------------------------------------------------------------------------------
  if (tid < TM) {
    const int le = le0 + tid;
    int s = -1, d = -1;
    float dist = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
    if (le < ecount) {
      int i, j;
      edge_src_dst(le, n, i, j);
      s = nb + i; d = nb + j;
      float dx, dy, dz;
      dist = pair_dist(x, s, d, dx, dy, dz);                       // x_src - x_dst  (dgl u_sub_v)
      ux = __fdiv_rn(dx, dist); uy = __fdiv_rn(dy, dist); uz = __fdiv_rn(dz, dist);
    }
    sm.src[tid] = s; sm.dst[tid] = d; sm.dist[tid] = dist;
    sm.Va[(0 * TM + tid) * D::LDVA] = ux;
    sm.Va[(1 * TM + tid) * D::LDVA] = uy;
    sm.Va[(2 * TM + tid) * D::LDVA] = uz;
  }
  __syncthreads();
  {
    const float* mu = m.g(G_RBF_MU);
    const float sigma = m.rbf_dmax / (float)D::R;
    for (int idx = tid; idx < TM * D::R; idx += NT) {              // scalar cols [0, R): rbf(d)
      const int row = idx / D::R, k = idx - row * D::R;
      sm.Xs[row * D::XLD + k] = sm.src[row] >= 0 ? rbf_f(sm.dist[row], mu[k], sigma) : 0.f;
    }
    for (int idx = tid; idx < TM * (D::F / 4); idx += NT) {        // scalar cols [R, R+F): edge features (HBM stream)
      const int row = idx / (D::F / 4), c4 = idx - row * (D::F / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sm.src[row] >= 0) val = __ldg(reinterpret_cast<const float4*>(ef + (erow0 + row) * D::F) + c4);
      *reinterpret_cast<float4*>(sm.Xs + row * D::XLD + D::R + c4 * 4) = val;
    }
    constexpr int VW = D::LDVA - 1;                                 // vector cols [1, LDVA): v_src | v_dst_msg | zero pad
    for (int idx = tid; idx < 3 * TM * VW; idx += NT) {
      const int pr = idx / VW, c = idx - pr * VW;                   // pr = plane * TM + row
      const int p = pr / TM, row = pr - p * TM;
      float val = 0.f;
      const int s = sm.src[row];
      if (s >= 0) {
        if (c < D::V) val = v[((size_t)s * 3 + p) * D::V + c];
        else if (D::VD > 0 && c < D::V + D::VD) val = vd[((size_t)sm.dst[row] * 3 + p) * D::VD + (c - D::V)];
      }
      sm.Va[pr * D::LDVA + 1 + c] = val;
    }
  }
  // ---- three message GVPs ------------------------------------------------------------------------------------------------
  {
    const GvpShape s0{D::VIN0, D::H0, D::CP, D::V, D::R + D::F, D::S, true};
    gvp_tile<D::CPT_S, D::CPT_HC0, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, s0,
                                                           gvp_ptr_conv(m, layer, C_MSG0_WHCP),
                                                           Msg0Pre<D>{P, Q, sm.src, sm.dst});
    const GvpShape s1{D::V, D::V, D::CP, D::V, D::S, D::S, true};
    const GvpPtr w1 = gvp_ptr_conv(m, layer, C_MSG1_WHCP);
    gvp_tile<D::CPT_S, D::CPT_HC, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, s1, w1, BiasPre{w1.b});
    const GvpPtr w2 = gvp_ptr_conv(m, layer, C_MSG2_WHCP);
    gvp_tile<D::CPT_S, D::CPT_HC, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, s1, w2, BiasPre{w2.b});
  }
  // ---- segment-sum over the in-edges of every dst (rows are sorted by dst) -----------------------------------------------
  // complete segments go straight to M[dst]; a segment cut by the tile boundary goes to partL (its head) or partF
  // (a continuation); k_node_update adds the pieces in tile order => deterministic, no atomics.
  for (int col = tid; col < D::MW; col += NT) {
    const float* base;
    int stride;
    if (col < D::S) { base = sm.Xs + col; stride = D::XLD; }
    else { const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V; base = sm.Va + p * TM * D::LDVA + c; stride = D::LDVA; }
    float acc = 0.f;
    int seg_first = le0;
    for (int row = 0; row < TM; ++row) {
      const int d = sm.dst[row];
      if (d < 0) break;
      acc = __fadd_rn(acc, base[row * stride]);
      const bool last = (row == TM - 1) || (sm.dst[row + 1] != d);
      if (last) {
        const int j = d - nb, le_last = le0 + row;
        const bool head = seg_first == j * (n - 1), tail = le_last == j * (n - 1) + (n - 2);
        if (head && tail) M[(size_t)d * D::MW + col] = acc;
        else if (head) partL[(size_t)tile * D::MW + col] = acc;
        else partF[(size_t)tile * D::MW + col] = acc;
        acc = 0.f;
        seg_first = le_last + 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// helpers for node tiles
// ------------------------------------------------------------------------------------------------------------------
// GVPLayerNorm on a node tile held as Xs[row][0:S) / Va[plane][row][0:V)   (gvp.py:169-184); one warp per 8 rows.
template <class D>
__device__ __forceinline__ void tile_gvp_layernorm(float* __restrict__ Xs, float* __restrict__ Va,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r;
    float* xr = Xs + row * D::XLD;
    float s = 0.f;
    for (int c = lane; c < D::S; c += 32) s += xr[c];
    const float mean = warp_sum(s) * (1.0f / D::S);
    float q = 0.f;
    for (int c = lane; c < D::S; c += 32) { const float d = xr[c] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D::S) + 1e-5f);
    for (int c = lane; c < D::S; c += 32) xr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    // vectors: v / (sqrt(mean_c clamp(|v_c|^2, 1e-8) + eps) + eps)
    float vq = 0.f;
    if (lane < D::V) {
      const float a = Va[(0 * TM + row) * D::LDVA + lane], b = Va[(1 * TM + row) * D::LDVA + lane],
                  c = Va[(2 * TM + row) * D::LDVA + lane];
      vq = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)), 1e-8f);
    }
    const float vn = __fadd_rn(sqrtf(__fadd_rn(warp_sum(vq) * (1.0f / D::V), 1e-5f)), 1e-5f);
    if (lane < D::V) {
#pragma unroll
      for (int p = 0; p < 3; ++p) Va[(p * TM + row) * D::LDVA + lane] = __fdiv_rn(Va[(p * TM + row) * D::LDVA + lane], vn);
    }
  }
}

// aggregated message of node g, column col: direct or pieces in tile order (see k_conv_edge)
template <class D>
__device__ __forceinline__ float gather_message(const BatchRT& bt, const float* __restrict__ M, const float* __restrict__ partF,
                                                const float* __restrict__ partL, int g, int mol, int col, int agg_rows) {
  // agg_rows = rows per tile of the conv-edge kernel that produced the pieces (64: k_conv_edge, 32: k_conv_edge_tc)
  const int n = bt.mol_n[mol], j = g - bt.mol_node[mol];
  const int first = j * (n - 1), last = first + n - 2;
  const int tb = bt.mol_etile[mol] * (TM / agg_rows);
  const int t0 = tb + first / agg_rows, t1 = tb + last / agg_rows;
  if (t0 == t1) return M[(size_t)g * D::MW + col];
  float acc = partL[(size_t)t0 * D::MW + col];
  for (int t = t0 + 1; t <= t1; ++t) acc = __fadd_rn(acc, partF[(size_t)t * D::MW + col]);
  return acc;
}

// Y[g][0:N) = Xs[:, 0:K) x W + b  for the valid rows of a node tile
template <int CPT, int XLD>
__device__ __forceinline__ void tile_linear_store(const float* __restrict__ Xs, int K, const float* __restrict__ W,
                                                  const float* __restrict__ b, float* __restrict__ wstage,
                                                  float* __restrict__ out, int ldo, int g0, int N, int ncols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[1][RPW][CPT];
  tile_gemm<1, CPT>(Xs, XLD, 0, K, W, wstage, acc);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int g = g0 + warp * RPW + r;
    if (g < N)
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int col = ColMap<CPT>::col(lane, c);
        if (col < ncols) out[(size_t)g * ldo + col] = acc[0][r][c] + b[col];
      }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_node_update: one tile = 64 nodes
// ------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(NT, 1)
k_node_update(const ModelRT m, const BatchRT bt, int layer, int updater /* -1: no molecule update after this conv */,
              int has_next, int agg_rows, float* __restrict__ s, float* __restrict__ v, float* __restrict__ x,
              const float* __restrict__ M, const float* __restrict__ partF, const float* __restrict__ partL,
              float* __restrict__ Pnext, float* __restrict__ EAB) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm(smem_raw);
  const int tid = threadIdx.x;
  const int g0 = blockIdx.x * TM;
  // ---- s + msg, v + msg --------------------------------------------------------------------------------------------------
  const float znorm = m.msg_norm;
  for (int idx = tid; idx < TM * D::MW; idx += NT) {
    const int row = idx / D::MW, col = idx - row * D::MW, g = g0 + row;
    float val = 0.f;
    if (g < bt.N) {
      const int mol = bt.node_mol[g];
      float msg = gather_message<D>(bt, M, partF, partL, g, mol, col, agg_rows);
      if (znorm > 0.f) msg = __fdiv_rn(msg, znorm);
      else if (znorm < 0.f) msg = __fdiv_rn(msg, (float)(bt.mol_n[mol] - 1));      // 'mean' over in-edges
      const float cur = col < D::S ? s[(size_t)g * D::S + col] : v[(size_t)g * 3 * D::V + (col - D::S)];
      val = __fadd_rn(cur, msg);
    }
    if (col < D::S) sm.Xs[row * D::XLD + col] = val;
    else { const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V; sm.Va[(p * TM + row) * D::LDVA + c] = val; }
  }
  __syncthreads();
  tile_gvp_layernorm<D>(sm.Xs, sm.Va, m.c(layer, C_LN_MSG_W), m.c(layer, C_LN_MSG_B));
  __syncthreads();
  // keep s', v' (needed for the second residual) in global memory: this tile owns these rows
  for (int idx = tid; idx < TM * D::MW; idx += NT) {
    const int row = idx / D::MW, col = idx - row * D::MW, g = g0 + row;
    if (g < bt.N) {
      if (col < D::S) s[(size_t)g * D::S + col] = sm.Xs[row * D::XLD + col];
      else { const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V; v[(size_t)g * 3 * D::V + (col - D::S)] = sm.Va[(p * TM + row) * D::LDVA + c]; }
    }
  }
  // ---- three update GVPs ------------------------------------------------------------------------------------------------
  const GvpShape sh{D::V, D::V, D::CP, D::V, D::S, D::S, true};
  for (int i = 0; i < 3; ++i) {
    const GvpPtr w = gvp_ptr_conv(m, layer, C_UPD0_WHCP + i * GV_COUNT);
    gvp_tile<D::CPT_S, D::CPT_HC, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, sh, w, BiasPre{w.b});
  }
  // ---- second residual + norm ----------------------------------------------------------------------------------------------
  for (int idx = tid; idx < TM * D::MW; idx += NT) {
    const int row = idx / D::MW, col = idx - row * D::MW, g = g0 + row;
    if (col < D::S) {
      const float cur = g < bt.N ? s[(size_t)g * D::S + col] : 0.f;
      sm.Xs[row * D::XLD + col] = __fadd_rn(cur, sm.Xs[row * D::XLD + col]);
    } else {
      const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V;
      const float cur = g < bt.N ? v[(size_t)g * 3 * D::V + (col - D::S)] : 0.f;
      sm.Va[(p * TM + row) * D::LDVA + c] = __fadd_rn(cur, sm.Va[(p * TM + row) * D::LDVA + c]);
    }
  }
  __syncthreads();
  tile_gvp_layernorm<D>(sm.Xs, sm.Va, m.c(layer, C_LN_UPD_W), m.c(layer, C_LN_UPD_B));
  __syncthreads();
  for (int idx = tid; idx < TM * D::MW; idx += NT) {
    const int row = idx / D::MW, col = idx - row * D::MW, g = g0 + row;
    if (g < bt.N) {
      if (col < D::S) s[(size_t)g * D::S + col] = sm.Xs[row * D::XLD + col];
      else { const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V; v[(size_t)g * 3 * D::V + (col - D::S)] = sm.Va[(p * TM + row) * D::LDVA + c]; }
    }
  }
  // ---- per-node halves of the following edge phases ----------------------------------------------------------------------------
  if (has_next)      // next conv's first message linear, rows that multiply s_src (bias folded in)
    tile_linear_store<D::CPT_S, D::XLD>(sm.Xs, D::S, m.c(layer + 1, C_WSRC), m.c(layer + 1, C_BSRC), sm.wstage, Pnext, D::S,
                                        g0, bt.N, D::S);
  if (updater >= 0) {
    // EdgeUpdate first linear, rows that multiply s_src (EA, with bias) and s_dst (EB)
    tile_linear_store<2 * D::CPT_F, D::XLD>(sm.Xs, D::S, m.u(updater, U_EUPD_WN), m.u(updater, U_EUPD_BN), sm.wstage, EAB,
                                            2 * D::F, g0, bt.N, 2 * D::F);
    // ---- NodePositionUpdate: x += last vector of 3 GVPs (vector_field.py:813-842) ---------------------------------------------
    const GvpPtr w0 = gvp_ptr_upd(m, updater, U_POS0_WHCP), w1 = gvp_ptr_upd(m, updater, U_POS1_WHCP),
                 w2 = gvp_ptr_upd(m, updater, U_POS2_WHCP);
    gvp_tile<D::CPT_S, D::CPT_HC, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, sh, w0, BiasPre{w0.b});
    gvp_tile<D::CPT_S, D::CPT_HC, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, sh, w1, BiasPre{w1.b});
    const GvpShape sh2{D::V, D::V, D::CP, 1, D::S, D::S, false};
    gvp_tile<D::CPT_S, D::CPT_HC, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, sh2, w2, BiasPre{w2.b});
    if (tid < TM * 3) {
      const int row = tid / 3, p = tid - row * 3, g = g0 + row;
      if (g < bt.N) x[g * 3 + p] = __fadd_rn(x[g * 3 + p], sm.Va[(p * TM + row) * D::LDVA]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_dst_proj (use_dst_feats only): (s_d, v_d) = GVP_dst(s, v) per node; Q = s_d x Wdst   (gvp.py:472-473,533-537)
// ------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(NT, 1)
k_dst_proj(const ModelRT m, const BatchRT bt, int layer, const float* __restrict__ s, const float* __restrict__ v,
           float* __restrict__ Q, float* __restrict__ vd) {
  pdl_launch();
  pdl_wait();
  if constexpr (D::SD > 0) {
    extern __shared__ __align__(16) float smem_raw[];
    Smem<D> sm(smem_raw);
    const int tid = threadIdx.x;
    const int g0 = blockIdx.x * TM;
    for (int idx = tid; idx < TM * D::MW; idx += NT) {
      const int row = idx / D::MW, col = idx - row * D::MW, g = g0 + row;
      float val = 0.f;
      if (g < bt.N) val = col < D::S ? s[(size_t)g * D::S + col] : v[(size_t)g * 3 * D::V + (col - D::S)];
      if (col < D::S) sm.Xs[row * D::XLD + col] = val;
      else { const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V; sm.Va[(p * TM + row) * D::LDVA + c] = val; }
    }
    const GvpShape sh{D::V, D::V, 0, D::VD, D::S, D::SD, true};
    const GvpPtr w = gvp_ptr_conv(m, layer, C_DST_WHCP);
    gvp_tile<D::CPT_SD, (D::V + 31) / 32, D::XLD, D::LDVA, D::LDVB>(sm.Xs, sm.Va, sm.Vb, sm.G, sm.wstage, sh, w, BiasPre{w.b});
    for (int idx = tid; idx < TM * 3 * D::VD; idx += NT) {
      const int row = idx / (3 * D::VD), pc = idx - row * 3 * D::VD, g = g0 + row;
      const int p = pc / D::VD, c = pc - p * D::VD;
      if (g < bt.N) vd[((size_t)g * 3 + p) * D::VD + c] = sm.Va[(p * TM + row) * D::LDVA + c];
    }
    // Q = s_d x Wdst  (no bias: the bias travels with P)
    const int lane = tid & 31, warp = tid >> 5;
    float acc[1][RPW][D::CPT_S];
    tile_gemm<1, D::CPT_S>(sm.Xs, D::XLD, 0, pad4(D::SD), m.c(layer, C_WDST), sm.wstage, acc);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int g = g0 + warp * RPW + r;
      if (g < bt.N)
#pragma unroll
        for (int c = 0; c < D::CPT_S; ++c) Q[(size_t)g * D::S + ColMap<D::CPT_S>::col(lane, c)] = acc[0][r][c];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// k_edge_update: one tile = 64 directed edges (same tiling as k_conv_edge); ef updated in place
// ------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(NT, 1)
k_edge_update(const ModelRT m, const BatchRT bt, int updater, const float* __restrict__ x, const float* __restrict__ EAB,
              float* __restrict__ ef) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, mol = bt.etile_mol[tile];
  const int n = bt.mol_n[mol], nb = bt.mol_node[mol], ecount = n * (n - 1);
  const int le0 = (tile - bt.mol_etile[mol]) * TM;
  const size_t erow0 = (size_t)tile * TM;
  if (tid < TM) {
    const int le = le0 + tid;
    int s = -1, d = -1;
    float dist = 0.f;
    if (le < ecount) {
      int i, j;
      edge_src_dst(le, n, i, j);
      s = nb + i; d = nb + j;
      float dx, dy, dz;
      dist = pair_dist(x, s, d, dx, dy, dz);
    }
    sm.src[tid] = s; sm.dst[tid] = d; sm.dist[tid] = dist;
  }
  __syncthreads();
  {
    const float* mu = m.g(G_RBF_MU);
    const float sigma = m.rbf_dmax / (float)D::R;
    for (int idx = tid; idx < TM * (D::F / 4); idx += NT) {        // cols [0, F): ef
      const int row = idx / (D::F / 4), c4 = idx - row * (D::F / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sm.src[row] >= 0) val = *(reinterpret_cast<const float4*>(ef + (erow0 + row) * D::F) + c4);
      *reinterpret_cast<float4*>(sm.Xs + row * D::XLD + c4 * 4) = val;
    }
    for (int idx = tid; idx < TM * D::R; idx += NT) {              // cols [F, F+R): rbf(d)
      const int row = idx / D::R, k = idx - row * D::R;
      sm.Xs[row * D::XLD + D::F + k] = sm.src[row] >= 0 ? rbf_f(sm.dist[row], mu[k], sigma) : 0.f;
    }
  }
  constexpr int K1 = D::F + D::R;
  float acc[1][RPW][D::CPT_F];
  tile_gemm<1, D::CPT_F>(sm.Xs, D::XLD, 0, K1, m.u(updater, U_EUPD_WE), sm.wstage, acc);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r;
    const int s = sm.src[row], d = sm.dst[row];
#pragma unroll
    for (int c = 0; c < D::CPT_F; ++c) {
      const int col = ColMap<D::CPT_F>::col(lane, c);
      float pre = 0.f;
      if (s >= 0) pre = __fadd_rn(EAB[(size_t)s * 2 * D::F + col], EAB[(size_t)d * 2 * D::F + D::F + col]);
      sm.Xs[row * D::XLD + K1 + col] = silu_f(acc[0][r][c] + pre);
    }
  }
  tile_gemm<1, D::CPT_F>(sm.Xs + K1, D::XLD, 0, D::F, m.u(updater, U_EUPD_W2), sm.wstage, acc);
  {
    const float* b = m.u(updater, U_EUPD_B2);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < D::CPT_F; ++c) {
        const int col = ColMap<D::CPT_F>::col(lane, c);
        acc[0][r][c] = __fadd_rn(sm.Xs[(warp * RPW + r) * D::XLD + col], silu_f(acc[0][r][c] + b[col]));
      }
    rows_layernorm<D::CPT_F>(acc[0], m.u(updater, U_EUPD_LN_W), m.u(updater, U_EUPD_LN_B));
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r;
    if (sm.src[row] >= 0) {
      if constexpr (D::CPT_F >= 4) {
#pragma unroll
        for (int g = 0; g < D::CPT_F / 4; ++g)
          *reinterpret_cast<float4*>(ef + (erow0 + row) * D::F + g * 128 + lane * 4) =
              make_float4(acc[0][r][g * 4], acc[0][r][g * 4 + 1], acc[0][r][g * 4 + 2], acc[0][r][g * 4 + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < D::CPT_F; ++c) ef[(erow0 + row) * D::F + ColMap<D::CPT_F>::col(lane, c)] = acc[0][r][c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// output heads
// ------------------------------------------------------------------------------------------------------------------
// softmax over lanes [lo, hi) of a warp-distributed logit row (lane = column)
__device__ __forceinline__ float lane_softmax(float logit, int lane, int lo, int hi) {
  const bool in = lane >= lo && lane < hi;
  const float mx = warp_max(in ? logit : -INFINITY);
  const float e = in ? expf(logit - mx) : 0.f;
  const float sum = warp_sum(e);
  return __fdiv_rn(e, sum);
}

template <class D>
__global__ void __launch_bounds__(NT, 1)
k_node_head(const ModelRT m, const BatchRT bt, const float* __restrict__ s, float* __restrict__ pa, float* __restrict__ pc) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * TM;
  for (int idx = tid; idx < TM * D::S; idx += NT) {
    const int row = idx / D::S, col = idx - row * D::S, g = g0 + row;
    sm.Xs[row * D::XLD + col] = g < bt.N ? s[(size_t)g * D::S + col] : 0.f;
  }
  float acc[1][RPW][D::CPT_S];
  tile_gemm<1, D::CPT_S>(sm.Xs, D::XLD, 0, D::S, m.g(G_NHEAD0_W), sm.wstage, acc);
  {
    const float* b = m.g(G_NHEAD0_B);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < D::CPT_S; ++c) {
        const int col = ColMap<D::CPT_S>::col(lane, c);
        sm.Xs[(warp * RPW + r) * D::XLD + col] = silu_f(acc[0][r][c] + b[col]);
      }
  }
  float lg[1][RPW][1];
  tile_gemm<1, 1>(sm.Xs, D::XLD, 0, D::S, m.g(G_NHEAD2_W), sm.wstage, lg);     // A + C <= 32 logits: lane = column
  const int A = m.A, C = m.C;
  const float bias = m.g(G_NHEAD2_B)[lane];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int g = g0 + warp * RPW + r;
    const float logit = lg[0][r][0] + bias;
    const float sa = lane_softmax(logit, lane, 0, A);
    const float sc = lane_softmax(logit, lane, A, A + C);
    if (g < bt.N) {
      if (lane < A) pa[(size_t)g * A + lane] = sa;
      else if (lane < A + C) pc[(size_t)g * C + (lane - A)] = sc;
    }
  }
}

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_edge_head(const ModelRT m, const BatchRT bt, const float* __restrict__ ef, float* __restrict__ pe) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = EdgeSmem<D>::carve(smem_raw);
  constexpr int XE = EdgeSmem<D>::XE, WST = EdgeSmem<D>::WST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, mol = bt.utile_mol[tile];
  const int n = bt.mol_n[mol], ucount = n * (n - 1) / 2;
  const int lu0 = (tile - bt.mol_utile[mol]) * TM;
  const size_t ebase = (size_t)bt.mol_etile[mol] * TM;
  const int ub = bt.mol_u[mol];
  if (tid < TM) {
    const int lu = lu0 + tid;
    int p0 = -1, p1 = -1;
    if (lu < ucount) {
      int i, j;
      upper_ij(lu, n, i, j);
      p0 = edge_pos(i, j, n);
      p1 = edge_pos(j, i, n);
    }
    sm.src[tid] = p0; sm.dst[tid] = p1;
  }
  __syncthreads();
  for (int idx = tid; idx < TM * (D::F / 4); idx += NT) {          // ef[upper] + ef[lower]   (vector_field.py:342-344)
    const int row = idx / (D::F / 4), c4 = idx - row * (D::F / 4);
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sm.src[row] >= 0) {
      const float4 a = *(reinterpret_cast<const float4*>(ef + (ebase + sm.src[row]) * D::F) + c4);
      const float4 b = *(reinterpret_cast<const float4*>(ef + (ebase + sm.dst[row]) * D::F) + c4);
      val = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    }
    *reinterpret_cast<float4*>(sm.Xs + row * XE + c4 * 4) = val;
  }
  float acc[1][RPW][D::CPT_F];
  tile_gemm<1, D::CPT_F, RPW, WST>(sm.Xs, XE, 0, D::F, m.g(G_EHEAD0_W), sm.wstage, acc);
  {
    const float* b = m.g(G_EHEAD0_B);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int c = 0; c < D::CPT_F; ++c) {
        const int col = ColMap<D::CPT_F>::col(lane, c);
        sm.Xs[(warp * RPW + r) * XE + col] = silu_f(acc[0][r][c] + b[col]);
      }
  }
  float lg[1][RPW][1];
  tile_gemm<1, 1, RPW, WST>(sm.Xs, XE, 0, D::F, m.g(G_EHEAD2_W), sm.wstage, lg);
  const int EB = m.EB;
  const float bias = m.g(G_EHEAD2_B)[lane];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r;
    const float p = lane_softmax(lg[0][r][0] + bias, lane, 0, EB);
    if (sm.src[row] >= 0 && lane < EB) pe[(size_t)(ub + lu0 + row) * EB + lane] = p;
  }
}

// COM removal: x_hat = x - mean_mol(x)   (vector_field.py:347-350); one warp per molecule
__global__ void k_com(const BatchRT bt, const float* __restrict__ x, float* __restrict__ px, int remove_com) {
  pdl_launch();
  pdl_wait();
  const int mol = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (mol >= bt.B) return;
  const int n = bt.mol_n[mol], nb = bt.mol_node[mol];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (remove_com) {
    for (int i = lane; i < n; i += 32) { sx += x[(nb + i) * 3]; sy += x[(nb + i) * 3 + 1]; sz += x[(nb + i) * 3 + 2]; }
    sx = __fdiv_rn(warp_sum(sx), (float)n); sy = __fdiv_rn(warp_sum(sy), (float)n); sz = __fdiv_rn(warp_sum(sz), (float)n);
  }
  for (int i = lane; i < n; i += 32) {
    px[(nb + i) * 3 + 0] = __fsub_rn(x[(nb + i) * 3 + 0], sx);
    px[(nb + i) * 3 + 1] = __fsub_rn(x[(nb + i) * 3 + 1], sy);
    px[(nb + i) * 3 + 2] = __fsub_rn(x[(nb + i) * 3 + 2], sz);
  }
}

}  // namespace fm
