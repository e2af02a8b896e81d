// Vector-channel stages of the message GVPs for the wide tensor-core pipeline (egemm_tc.cuh): fp32 CUDA-core kernels on
// 64-edge tiles (same molecule-aligned tiling and segment-sum scheme as k_conv_edge).
//
//   k_vec_a : gather x / v of src, x_diff;  [Vh | Vcp] = V [Wh | Wcp], cross products, norms     (GVP 0, stage 1)
//   k_vec_b : V' = gate * (Vh_ext Wu)  (GVP g, stage 2)  then stage 1 of GVP g+1
//   k_vec_c : V' of GVP 2, then the segment-sum over in-edges of the vector AND scalar messages -> M / partL / partF
//
// Global intermediates per padded edge slot: VH [3][40] hidden vectors (Vh | cross), SH [40] their norms (zero padded),
// GT [32] gates, S [256] scalar activations (written by k_egemm_tc).
#pragma once
#include "kernels.cuh"
#include "mma3.cuh"

namespace fm {

constexpr int VHW = 40;   // row pitch of VH planes and SH
constexpr int WLD_HCP = 72;   // shared-memory row pitch of [Wh | Wcp] (64 packed columns + 8: conflict-free B fragments)
constexpr int WLD_U = 40;     // shared-memory row pitch of Wu (32 + 8)

// compact shared-memory plan for the stages that never touch the scalar tile: 2 CTAs per SM (latency hiding)
template <class D>
struct VecSmem {
  static constexpr int FLOATS = D::SM_VA + D::SM_VB + D::SM_G + WSTAGE_FLOATS / 2 + D::SM_MISC;
  static constexpr size_t BYTES = (size_t)FLOATS * 4;
  __device__ static Smem<D> carve(float* base) {
    Smem<D> sm(base);
    sm.Xs = nullptr;
    sm.Va = base;
    sm.Vb = sm.Va + D::SM_VA;
    sm.G = sm.Vb + D::SM_VB;
    sm.wstage = sm.G + D::SM_G;
    float* misc = sm.wstage + WSTAGE_FLOATS / 2;
    sm.src = reinterpret_cast<int*>(misc);
    sm.dst = sm.src + TM;
    sm.dist = misc + 2 * TM;
    sm.aux = reinterpret_cast<int*>(misc + 3 * TM);
    return sm;
  }
};

template <class D>
struct EdgeTile {
  int mol, n, nb, ecount, le0;
  size_t erow0;
  int tile;
  __device__ EdgeTile(const BatchRT& bt, int tile_) : tile(tile_) {
    mol = bt.etile_mol[tile];
    n = bt.mol_n[mol]; nb = bt.mol_node[mol]; ecount = n * (n - 1);
    le0 = (tile - bt.mol_etile[mol]) * TM;
    erow0 = (size_t)tile * TM;
  }
};

// stage 1 of a GVP on the tile in shared memory: Va[.., 0:v_in) -> Vb = [Vh | cross] (cols [0, h+cp)), stores VH and SH
template <class D, int CPT_HC>
__device__ __forceinline__ void vec_stage1(Smem<D>& sm, const int v_in, const int h, const float* __restrict__ whcp_sm, const size_t erow0,
                                           float* __restrict__ VH, float* __restrict__ SH) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hc = h + D::CP;
  {
    // [Vh | Vcp] = V x [Wh | Wcp]: 192 (edge, plane) rows x (h + 2cp <= 48) columns.  Warp w: 3 m16 tiles x 3 n8 tiles.
    __syncthreads();
    const int m0 = (warp >> 1) * 48, n0 = (warp & 1) * 24;
    float acc[3][3][4];
    warp_gemm_3xtf32<3, 3>(sm.Va, D::LDVA, m0, whcp_sm, WLD_HCP, n0, (v_in + 7) & ~7, acc);
    const int ncol = h + 2 * D::CP, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = m0 + 16 * mt + g + (i >> 1) * 8, col = n0 + 8 * nt + 2 * t + (i & 1);
          if (col < ncol) sm.Vb[row * D::LDVB + col] = acc[mt][nt][i];
        }
  }
  __syncthreads();
  for (int idx = tid; idx < TM * D::CP; idx += NT) {
    const int row = idx / D::CP, j = idx - row * D::CP;
    float* bx = sm.Vb + (0 * TM + row) * D::LDVB;
    float* by = sm.Vb + (1 * TM + row) * D::LDVB;
    float* bz = sm.Vb + (2 * TM + row) * D::LDVB;
    const int ca = h + j, cb = h + D::CP + j;
    const float ax = bx[ca], ay = by[ca], az = bz[ca], qx = bx[cb], qy = by[cb], qz = bz[cb];
    bx[ca] = __fsub_rn(__fmul_rn(ay, qz), __fmul_rn(az, qy));
    by[ca] = __fsub_rn(__fmul_rn(az, qx), __fmul_rn(ax, qz));
    bz[ca] = __fsub_rn(__fmul_rn(ax, qy), __fmul_rn(ay, qx));
  }
  __syncthreads();
  for (int idx = tid; idx < TM * (VHW / 4); idx += NT) {
    const int row = idx / (VHW / 4), c4 = idx - row * (VHW / 4);
    const bool okr = sm.src[row] >= 0;
    float va[4], vb_[4], vc[4], nn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c4 * 4 + k;
      const bool ok = okr && c < hc;
      va[k] = ok ? sm.Vb[(0 * TM + row) * D::LDVB + c] : 0.f;
      vb_[k] = ok ? sm.Vb[(1 * TM + row) * D::LDVB + c] : 0.f;
      vc[k] = ok ? sm.Vb[(2 * TM + row) * D::LDVB + c] : 0.f;
      nn[k] = ok ? norm_no_nan3(va[k], vb_[k], vc[k]) : 0.f;
    }
    float4* vh = reinterpret_cast<float4*>(VH + (erow0 + row) * 3 * VHW);
    vh[c4] = make_float4(va[0], va[1], va[2], va[3]);
    vh[VHW / 4 + c4] = make_float4(vb_[0], vb_[1], vb_[2], vb_[3]);
    vh[2 * (VHW / 4) + c4] = make_float4(vc[0], vc[1], vc[2], vc[3]);
    reinterpret_cast<float4*>(SH + (erow0 + row) * VHW)[c4] = make_float4(nn[0], nn[1], nn[2], nn[3]);
  }
}

// stage 2 of a GVP: Vb (loaded from VH) x Wu, gated by GT -> Va[.., 0:V)
template <class D>
__device__ __forceinline__ void vec_stage2(Smem<D>& sm, const int hc, const float* __restrict__ wu_sm, const size_t erow0,
                                           const float* __restrict__ VH, const float* __restrict__ GT) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    // the tile's VH block (64 x 120 floats) and GT block (64 x 32) are contiguous in memory: straight float4 streams, all
    // loads of a thread in flight together (padding slots hold finite garbage or zeros and are masked at the segment-sum)
    constexpr int NV = TM * 3 * VHW / 4, NG = TM * 32 / 4;
    const float4* vsrc = reinterpret_cast<const float4*>(VH + erow0 * 3 * VHW);
    const float4* gsrc = reinterpret_cast<const float4*>(GT + erow0 * 32);
    float4 vb[(NV + NT - 1) / NT], gb[(NG + NT - 1) / NT];
#pragma unroll
    for (int i = 0; i < (NV + NT - 1) / NT; ++i) { const int idx = tid + i * NT; if (idx < NV) vb[i] = vsrc[idx]; }
#pragma unroll
    for (int i = 0; i < (NG + NT - 1) / NT; ++i) { const int idx = tid + i * NT; if (idx < NG) gb[i] = gsrc[idx]; }
#pragma unroll
    for (int i = 0; i < (NV + NT - 1) / NT; ++i) {
      const int idx = tid + i * NT;
      if (idx < NV) {
        const int row = idx / (3 * VHW / 4), rem = idx - row * (3 * VHW / 4), p = rem / (VHW / 4), c4 = rem - p * (VHW / 4);
        const bool ok = sm.src[row] >= 0;
        *reinterpret_cast<float4*>(sm.Vb + (p * TM + row) * D::LDVB + c4 * 4) = ok ? vb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < (NG + NT - 1) / NT; ++i) {
      const int idx = tid + i * NT;
      if (idx < NG) *reinterpret_cast<float4*>(sm.G + idx * 4) = sm.src[idx >> 3] >= 0 ? gb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  {
    // Vu = Vh_ext x Wu: 192 rows x 32 columns, K = hc padded to 8.  Warp w: 3 m16 tiles x 2 n8 tiles.
    const int m0 = (warp >> 1) * 48, n0 = (warp & 1) * 16, g = lane >> 2, t = lane & 3;
    float acc[3][2][4];
    warp_gemm_3xtf32<3, 2>(sm.Vb, D::LDVB, m0, wu_sm, WLD_U, n0, (hc + 7) & ~7, acc);
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = m0 + 16 * mt + g + (i >> 1) * 8, col = n0 + 8 * nt + 2 * t + (i & 1);
          sm.Va[row * D::LDVA + col] = __fmul_rn(sm.G[(row & (TM - 1)) * 32 + col], acc[mt][nt][i]);
        }
  }
  __syncthreads();
}

// the CTA's weight matrices, once per CTA: src [K][np] (packer layout) -> dst [K padded to 8][ld] (padding rows zero)
__device__ __forceinline__ void load_resident(float* dst, const float* __restrict__ src, int K, int np, int ld) {
  const int kp = (K + 7) & ~7;
  for (int i = threadIdx.x; i < kp * (np / 4); i += NT) {
    const int k = i / (np / 4), c4 = i - k * (np / 4);
    if (k < K) cp_async16(dst + k * ld + c4 * 4, src + k * np + c4 * 4);
    else *reinterpret_cast<float4*>(dst + k * ld + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <class D>
__device__ __forceinline__ void tile_rows(Smem<D>& sm, const EdgeTile<D>& et) {
  if (threadIdx.x < TM) {
    const int le = et.le0 + threadIdx.x;
    int s = -1, d = -1;
    if (le < et.ecount) {
      int i, j;
      edge_src_dst(le, et.n, i, j);
      s = et.nb + i; d = et.nb + j;
    }
    sm.src[threadIdx.x] = s; sm.dst[threadIdx.x] = d;
  }
  __syncthreads();
}

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vec_a(const ModelRT m, const BatchRT bt, int layer, const float* __restrict__ x, const float* __restrict__ v,
        float* __restrict__ VH, float* __restrict__ SH) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x;
  float* w_hcp = sm.wstage;                                       // [pad4(VIN0)][64]
  load_resident(w_hcp, m.c(layer, C_MSG0_WHCP), pad4(D::VIN0), 32 * D::CPT_HC0, WLD_HCP);
  cp_async_commit();
  cp_async_wait<0>();
  for (int tile = blockIdx.x; tile < bt.n_edge_tiles; tile += gridDim.x) {
  const EdgeTile<D> et(bt, tile);
  __syncthreads();                                                // previous tile's readers of src/dst/Va are done
  tile_rows<D>(sm, et);
  if (tid < TM) {
    const int s = sm.src[tid], d = sm.dst[tid];
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (s >= 0) {
      float dx, dy, dz;
      const float dist = pair_dist(x, s, d, dx, dy, dz);
      ux = __fdiv_rn(dx, dist); uy = __fdiv_rn(dy, dist); uz = __fdiv_rn(dz, dist);
    }
    sm.Va[(0 * TM + tid) * D::LDVA] = ux;
    sm.Va[(1 * TM + tid) * D::LDVA] = uy;
    sm.Va[(2 * TM + tid) * D::LDVA] = uz;
  }
  {
    // v of the source node: 3 planes x V floats per edge, float4 gathers all in flight (v is L2 resident), then the K padding
    constexpr int Q4 = D::V / 4, NQ = 3 * TM * Q4, PER = (NQ + NT - 1) / NT;
    float4 buf[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = tid + i * NT;
      buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < NQ) {
        const int row = idx / (3 * Q4), rem = idx - row * (3 * Q4);          // rem = plane * Q4 + quad: contiguous in v[src]
        const int s = sm.src[row];
        if (s >= 0) buf[i] = __ldg(reinterpret_cast<const float4*>(v + (size_t)s * 3 * D::V) + rem);
      }
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = tid + i * NT;
      if (idx < NQ) {
        const int row = idx / (3 * Q4), rem = idx - row * (3 * Q4), p = rem / Q4, q = rem - p * Q4;
        float* d = sm.Va + (p * TM + row) * D::LDVA + 1 + q * 4;
        d[0] = buf[i].x; d[1] = buf[i].y; d[2] = buf[i].z; d[3] = buf[i].w;
      }
    }
    constexpr int PADW = D::LDVA - 1 - D::V;
    for (int idx = tid; idx < 3 * TM * PADW; idx += NT) {
      const int pr = idx / PADW, c = idx - pr * PADW;
      sm.Va[pr * D::LDVA + 1 + D::V + c] = 0.f;
    }
  }
  vec_stage1<D, D::CPT_HC0>(sm, D::VIN0, D::H0, w_hcp, et.erow0, VH, SH);
  }
}

// validity of the 64 rows of a node tile (rows = nodes; src >= 0 marks a live row for the vector stages)
template <class D>
__device__ __forceinline__ void node_tile_rows(Smem<D>& sm, int g0, int N) {
  if (threadIdx.x < TM) { sm.src[threadIdx.x] = g0 + (int)threadIdx.x < N ? 0 : -1; sm.dst[threadIdx.x] = -1; }
  __syncthreads();
}

// stage 2 of one GVP (Wu [hc_prev][32]) then stage 1 of the next (Whcp [V][64]); rows are padded edge slots, or nodes
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vec_b(const BatchRT bt, const float* __restrict__ wu, int hc_prev, const float* __restrict__ whcp, int node_rows,
        float* __restrict__ VH, float* __restrict__ SH, const float* __restrict__ GT) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  float* w_u = sm.wstage;                                         // [pad8(hc_prev)][32 (+8)]
  float* w_hcp = sm.wstage + 40 * WLD_U;                          // [V][64 (+8)]
  load_resident(w_u, wu, pad4(hc_prev), 32, WLD_U);
  load_resident(w_hcp, whcp, D::V, 32 * D::CPT_HC, WLD_HCP);
  cp_async_commit();
  cp_async_wait<0>();
  const int n_tiles = node_rows ? bt.n_node_tiles : bt.n_edge_tiles;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (node_rows) node_tile_rows<D>(sm, tile * TM, bt.N);
    else { const EdgeTile<D> et(bt, tile); tile_rows<D>(sm, et); }
    const size_t erow0 = (size_t)tile * TM;
    vec_stage2<D>(sm, hc_prev, w_u, erow0, VH, GT);
    vec_stage1<D, D::CPT_HC>(sm, D::V, D::V, w_hcp, erow0, VH, SH);
  }
}

// ---- node update as a pipeline around k_egemm_tc (node rows) ----------------------------------------------------------------------
// GVPLayerNorm (gvp.py:169-184), scalar half: one warp per row, lanes over columns; `f(col)` is the pre-norm value
template <class D, class F>
__device__ __forceinline__ void row_scalar_layernorm(float* __restrict__ out_row, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, F f) {
  const int lane = threadIdx.x & 31;
  float val[D::CPT_S];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < D::CPT_S; ++c) { val[c] = f(lane + 32 * c); s += val[c]; }
  const float mean = warp_sum(s) * (1.0f / D::S);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < D::CPT_S; ++c) { const float d = val[c] - mean; q = fmaf(d, d, q); }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D::S) + 1e-5f);
#pragma unroll
  for (int c = 0; c < D::CPT_S; ++c) {
    const int col = lane + 32 * c;
    out_row[col] = (val[c] - mean) * rstd * gamma[col] + beta[col];
  }
}
// vector half on the tile in Va: v / (sqrt(mean_c clamp(|v_c|^2, 1e-8) + eps) + eps); the result is also stored to v[g]
template <class D>
__device__ __forceinline__ void tile_vec_layernorm(float* __restrict__ Va, float* __restrict__ v, int g0, int N) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r, g = g0 + row;
    float a = 0.f, b = 0.f, c = 0.f, vq = 0.f;
    if (lane < D::V) {
      a = Va[(0 * TM + row) * D::LDVA + lane]; b = Va[(1 * TM + row) * D::LDVA + lane]; c = Va[(2 * TM + row) * D::LDVA + lane];
      vq = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)), 1e-8f);
    }
    const float vn = __fadd_rn(sqrtf(__fadd_rn(warp_sum(vq) * (1.0f / D::V), 1e-5f)), 1e-5f);
    if (lane < D::V) {
      a = __fdiv_rn(a, vn); b = __fdiv_rn(b, vn); c = __fdiv_rn(c, vn);
      Va[(0 * TM + row) * D::LDVA + lane] = a; Va[(1 * TM + row) * D::LDVA + lane] = b; Va[(2 * TM + row) * D::LDVA + lane] = c;
      if (g < N) {
        v[((size_t)g * 3 + 0) * D::V + lane] = a; v[((size_t)g * 3 + 1) * D::V + lane] = b; v[((size_t)g * 3 + 2) * D::V + lane] = c;
      }
    }
  }
}

// k_node_pre: (s, v) <- GVPLayerNorm((s, v) + aggregated message)  (gvp.py:509-513), then stage 1 of the first update GVP
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_node_pre(const ModelRT m, const BatchRT bt, int layer, int agg_rows, float* __restrict__ s, float* __restrict__ v,
           const float* __restrict__ M, const float* __restrict__ partF, const float* __restrict__ partL,
           float* __restrict__ VH, float* __restrict__ SH) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  float* w_hcp = sm.wstage;
  load_resident(w_hcp, m.c(layer, C_UPD0_WHCP), D::V, 32 * D::CPT_HC, WLD_HCP);
  cp_async_commit();
  const int g0 = blockIdx.x * TM;
  node_tile_rows<D>(sm, g0, bt.N);
  // per-row aggregation plan (see gather_message): first / last piece of the node's in-edge segment, and the normaliser
  if (tid < TM) {
    const int g = g0 + tid;
    int t0 = 0, t1 = 0;
    float div = 0.f;
    if (g < bt.N) {
      const int mol = bt.node_mol[g], n = bt.mol_n[mol], j = g - bt.mol_node[mol];
      const int first = j * (n - 1), last = first + n - 2, tb = bt.mol_etile[mol] * (TM / agg_rows);
      t0 = tb + first / agg_rows; t1 = tb + last / agg_rows;
      div = m.msg_norm > 0.f ? m.msg_norm : (m.msg_norm < 0.f ? (float)(n - 1) : 0.f);      // 'mean': over the in-edges
    }
    sm.dst[tid] = t0; sm.aux[tid] = t1; sm.dist[tid] = div;
  }
  __syncthreads();
  auto message = [&](int row, int g, int col) {
    const int t0 = sm.dst[row], t1 = sm.aux[row];
    float msg;
    if (t0 == t1) msg = M[(size_t)g * D::MW + col];
    else {
      msg = partL[(size_t)t0 * D::MW + col];
      for (int t = t0 + 1; t <= t1; ++t) msg = __fadd_rn(msg, partF[(size_t)t * D::MW + col]);
    }
    const float div = sm.dist[row];
    return div != 0.f ? __fdiv_rn(msg, div) : msg;
  };
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r, g = g0 + row;
    if (g >= bt.N) break;
    float* srow = s + (size_t)g * D::S;
    row_scalar_layernorm<D>(srow, m.c(layer, C_LN_MSG_W), m.c(layer, C_LN_MSG_B),
                            [&](int col) { return __fadd_rn(srow[col], message(row, g, col)); });
  }
#pragma unroll 4
  for (int idx = tid; idx < TM * 3 * D::V; idx += NT) {
    const int row = idx / (3 * D::V), pc = idx - row * 3 * D::V, g = g0 + row;
    float val = 0.f;
    if (g < bt.N) val = __fadd_rn(v[(size_t)g * 3 * D::V + pc], message(row, g, D::S + pc));
    sm.Va[((pc / D::V) * TM + row) * D::LDVA + (pc % D::V)] = val;
  }
  __syncthreads();
  tile_vec_layernorm<D>(sm.Va, v, g0, bt.N);
  cp_async_wait<0>();
  vec_stage1<D, D::CPT_HC>(sm, D::V, D::V, w_hcp, (size_t)g0, VH, SH);
}

// k_node_mid: stage 2 of the last update GVP, (s, v) <- GVPLayerNorm((s, v) + update)  (gvp.py:515-519), then stage 1 of the first
// NodePositionUpdate GVP when this conv is followed by a molecule update
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_node_mid(const ModelRT m, const BatchRT bt, int layer, int updater, float* __restrict__ s, float* __restrict__ v,
           const float* __restrict__ S3, float* __restrict__ VH, float* __restrict__ SH, const float* __restrict__ GT) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  float* w_u = sm.wstage;
  float* w_hcp = sm.wstage + 40 * WLD_U;
  load_resident(w_u, m.c(layer, C_UPD2_WU), pad4(D::V + D::CP), 32, WLD_U);
  if (updater >= 0) load_resident(w_hcp, m.u(updater, U_POS0_WHCP), D::V, 32 * D::CPT_HC, WLD_HCP);
  cp_async_commit();
  cp_async_wait<0>();
  const int g0 = blockIdx.x * TM;
  __syncthreads();
  node_tile_rows<D>(sm, g0, bt.N);
  vec_stage2<D>(sm, D::V + D::CP, w_u, (size_t)g0, VH, GT);
  for (int r = 0; r < RPW; ++r) {
    const int g = g0 + warp * RPW + r;
    if (g >= bt.N) break;
    float* srow = s + (size_t)g * D::S;
    const float* urow = S3 + (size_t)g * D::S;
    row_scalar_layernorm<D>(srow, m.c(layer, C_LN_UPD_W), m.c(layer, C_LN_UPD_B),
                            [&](int col) { return __fadd_rn(srow[col], urow[col]); });
  }
  for (int idx = tid; idx < TM * 3 * D::V; idx += NT) {
    const int row = idx / (3 * D::V), pc = idx - row * 3 * D::V, g = g0 + row;
    float* p = sm.Va + ((pc / D::V) * TM + row) * D::LDVA + (pc % D::V);
    *p = g < bt.N ? __fadd_rn(v[(size_t)g * 3 * D::V + pc], *p) : 0.f;
  }
  __syncthreads();
  tile_vec_layernorm<D>(sm.Va, v, g0, bt.N);
  if (updater >= 0) vec_stage1<D, D::CPT_HC>(sm, D::V, D::V, w_hcp, (size_t)g0, VH, SH);
}

// k_node_post: stage 2 of the last position GVP (one output vector); x += that vector   (vector_field.py:813-842)
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_node_post(const ModelRT m, const BatchRT bt, int updater, float* __restrict__ x, const float* __restrict__ VH,
            const float* __restrict__ GT) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x;
  float* w_u = sm.wstage;
  load_resident(w_u, m.u(updater, U_POS2_WU), pad4(D::V + D::CP), 32, WLD_U);
  cp_async_commit();
  cp_async_wait<0>();
  const int g0 = blockIdx.x * TM;
  __syncthreads();
  node_tile_rows<D>(sm, g0, bt.N);
  vec_stage2<D>(sm, D::V + D::CP, w_u, (size_t)g0, VH, GT);
  if (tid < TM * 3) {
    const int row = tid / 3, p = tid - row * 3, g = g0 + row;
    if (g < bt.N) x[g * 3 + p] = __fadd_rn(x[g * 3 + p], sm.Va[(p * TM + row) * D::LDVA]);
  }
}

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vec_c(const ModelRT m, const BatchRT bt, int layer, int first_col /* S when k_egemm_tc<EG_MSGA> reduced the scalar columns */,
        const float* __restrict__ VH, const float* __restrict__ GT, const float* __restrict__ Smsg, float* __restrict__ M, float* __restrict__ partF, float* __restrict__ partL) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x;
  float* w_u = sm.wstage;
  load_resident(w_u, m.c(layer, C_MSG2_WU), pad4(D::V + D::CP), 32, WLD_U);
  cp_async_commit();
  cp_async_wait<0>();
  for (int tile = blockIdx.x; tile < bt.n_edge_tiles; tile += gridDim.x) {
  const EdgeTile<D> et(bt, tile);
  __syncthreads();
  tile_rows<D>(sm, et);
  vec_stage2<D>(sm, D::V + D::CP, w_u, et.erow0, VH, GT);
  for (int col = first_col + tid; col < D::MW; col += NT) {
    float acc = 0.f;
    int seg_first = et.le0;
    for (int r0 = 0; r0 < TM; r0 += 16) {
      // scalar columns stream from HBM: 16 independent coalesced loads in flight per thread (padding slots exist in memory)
      float buf[16];
      if (col < D::S) {
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[k] = __ldg(Smsg + (et.erow0 + r0 + k) * D::S + col);
      } else {
        const int p = (col - D::S) / D::V, c = (col - D::S) - p * D::V;
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[k] = sm.Va[(p * TM + r0 + k) * D::LDVA + c];
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int row = r0 + k;
        const int d = sm.dst[row];
        if (d >= 0) {
          acc = __fadd_rn(acc, buf[k]);
          if (row == TM - 1 || sm.dst[row + 1] != d) {
            const int j = d - et.nb, le_last = et.le0 + row;
            const bool head = seg_first == j * (et.n - 1), tail = le_last == j * (et.n - 1) + (et.n - 2);
            if (head && tail) M[(size_t)d * D::MW + col] = acc;
            else if (head) partL[(size_t)tile * D::MW + col] = acc;
            else partF[(size_t)tile * D::MW + col] = acc;
            acc = 0.f;
            seg_first = le_last + 1;
          }
        }
      }
    }
  }
  }
}

}  // namespace fm
