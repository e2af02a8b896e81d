// Vector-channel stages of the GVPs for the wide tensor-core pipeline (egemm_tc.cuh): CUDA-core / warp-MMA kernels on
// 64-row tiles (rows = padded edge slots with the molecule-aligned tiling of k_conv_edge, or nodes).
//
//   k_vec_a : gather x / v of src, x_diff;  [Vh | Vcp] = V [Wh | Wcp], cross products, norms     (message GVP 0, stage 1)
//   k_vec_b : V' = gate * (Vh_ext Wu)  (GVP g, stage 2)  then stage 1 of GVP g+1                   (edge or node rows)
//   k_vec_c : V' of message GVP 2, then the segment-sum over in-edges of the vector messages (and, without EG_MSGA, of the
//             scalar messages) -> M / partL / partF
//   k_node_pre / k_node_mid / k_node_post : the node update around k_egemm_tc on node rows (see api.cu:node_wide)
//
// Global intermediates per row: VH [3][40] hidden vectors (Vh | cross), SH [40] their norms (zero padded), GT [32] gates,
// S [256] scalar activations (written by k_egemm_tc).
//
// Work decomposition: a CTA (8 warps) owns a 64-row tile, but each WARP PAIR owns 16 rows of it end to end -- their 48
// (row, plane) vectors are consecutive "flat rows" (row * 3 + plane) of the shared-memory operands, i.e. three m16 MMA tiles.
// Both GEMMs, the cross products and the norms of those 16 rows only ever need the pair's own data, so the stages synchronise
// with 64-thread named barriers and the four pairs of a CTA drift apart, hiding each other's global-memory latency.
#pragma once
#include "kernels.cuh"
#include "mma3.cuh"

namespace fm {

constexpr int VHW = 40;       // row pitch of VH planes and SH
constexpr int WLD_HCP = 72;   // shared-memory row pitch of [Wh | Wcp] (64 packed columns + 8: conflict-free B fragments)
constexpr int WLD_U = 40;     // shared-memory row pitch of Wu (32 + 8)
constexpr int PE = 16;        // rows owned by one warp pair
constexpr int PT = 64;        // threads of a warp pair

// compact shared-memory plan for the stages that never touch the scalar tile: 2 CTAs per SM (latency hiding)
template <class D>
struct VecSmem {
  static constexpr int FLOATS = D::SM_VA + D::SM_VB + D::SM_G + WSTAGE_FLOATS / 2 + D::SM_MISC;
  static constexpr size_t BYTES = (size_t)FLOATS * 4;
  __device__ static Smem<D> carve(float* base) {
    Smem<D> sm(base);
    sm.Xs = nullptr;
    sm.Va = base;                      // [64 rows * 3 planes][LDVA]   flat row = row * 3 + plane
    sm.Vb = sm.Va + D::SM_VA;          // [64 rows * 3 planes][LDVB]
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

// 64-thread named barrier of warp pair `pair` (ids 1..4; immediate ids so that the kernel reserves 5 barriers, not all 16)
__device__ __forceinline__ void pair_sync(int pair) {
  switch (pair) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

template <class D>
struct EdgeTile {
  int mol, n, nb, ecount, le0, nvalid;
  size_t erow0;
  int tile;
  __device__ EdgeTile(const BatchRT& bt, int tile_) : tile(tile_) {
    mol = bt.etile_mol[tile];
    n = bt.mol_n[mol]; nb = bt.mol_node[mol]; ecount = n * (n - 1);
    le0 = (tile - bt.mol_etile[mol]) * TM;
    nvalid = min(TM, ecount - le0);                    // live rows are a prefix of the tile
    erow0 = (size_t)tile * TM;
  }
};

// The CTA's weight matrices, once per CTA, as fp16 (hi, lo) operand words for warp_gemm_h16x3 (mma3.cuh):
// src [K][np] fp32 (packer layout) -> hi / lo [pad8(K) / 2][ld] packed (k, k + 1) half2 words of w * 2^s, padding rows zero.  The
// power-of-two scale puts max |w| 2^s into [2^13, 2^14) (the `lo` parts of every weight that matters stay normal fp16 numbers) and
// is undone exactly on the accumulators (`inv`).  Whole CTA; contains barriers.  `red`: >= 8 floats of idle shared memory.
struct WH16 {
  const uint32_t* hi;
  const uint32_t* lo;
  float inv;
};
__device__ __forceinline__ WH16 load_resident_h16(float* dst, const float* __restrict__ src, int K, int np, int ld, float* red) {
  const int tid = threadIdx.x;
  float mx = 0.f;
  for (int i = tid; i < K * np; i += NT) mx = fmaxf(mx, fabsf(__ldg(src + i)));
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = 0.f;
#pragma unroll
  for (int w = 0; w < NWARP; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();                                                   // `red` may be reused by the caller
  const int e = mx >= 1e-30f ? (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127 : 13;   // floor(log2(max |w|)); all-zero matrix: scale 1
  const float scale = __uint_as_float((uint32_t)(127 + 13 - e) << 23), inv = __uint_as_float((uint32_t)(127 - 13 + e) << 23);
  const int kp2 = ((K + 7) & ~7) >> 1;
  uint32_t* hi = reinterpret_cast<uint32_t*>(dst);
  uint32_t* lo = hi + kp2 * ld;
  for (int i = tid; i < kp2 * np; i += NT) {
    const int r = i / np, n = i - r * np, k = 2 * r;
    const float w0 = k < K ? __ldg(src + k * np + n) * scale : 0.f, w1 = k + 1 < K ? __ldg(src + (k + 1) * np + n) * scale : 0.f;
    uint32_t h2, l2;
    tc::split_h16x2(w0, w1, h2, l2);
    hi[r * ld + n] = h2;
    lo[r * ld + n] = l2;
  }
  return WH16{hi, lo, inv};
}

// HBM -> L2 prefetch of the 16-row VH / GT blocks a warp pair will read for its NEXT tile (one bulk-prefetch instruction each, no
// registers held): the loads at the top of the next tile then see L2 latency instead of HBM latency (ncu r01o: ~20 % of the
// vector-stage kernels' samples sat on those loads and the barrier behind them).
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_pair_inputs(const float* __restrict__ VH, const float* __restrict__ GT, size_t erow0_next) {
  if ((threadIdx.x & (PT - 1)) == 0) {
    const size_t r = erow0_next + (size_t)(threadIdx.x >> 6) * PE;
    prefetch_l2(VH + r * 3 * VHW, PE * 3 * VHW * 4);
    prefetch_l2(GT + r * 32, PE * 32 * 4);
  }
}

// stage 1 of a GVP for the pair's 16 rows: Va[.., 0:v_in) -> Vb = [Vh | cross] (cols [0, h+cp)), stores VH and SH
template <class D>
__device__ __forceinline__ void vec_stage1(Smem<D>& sm, const int v_in, const int h, const WH16& whcp_sm, const size_t erow0,
                                           const int nvalid, float* __restrict__ VH, float* __restrict__ SH) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pair = warp >> 1, pt = tid & (PT - 1);
  const int e0 = pair * PE, r0 = e0 * 3, hc = h + D::CP;
  pair_sync(pair);                                       // the pair's Va rows are complete
  {
    // [Vh | Vcp] = V x [Wh | Wcp]: 48 flat rows x (h + 2cp <= 48) columns.  Each warp: 3 m16 tiles x 3 n8 tiles.
    const int n0 = (warp & 1) * 24;
    float acc[3][3][4];
    warp_gemm_h16x3<3, 3>(sm.Va, D::LDVA, r0, whcp_sm.hi, whcp_sm.lo, WLD_HCP, n0, (v_in + 7) & ~7, acc);
    const float inv = whcp_sm.inv;
    const int ncol = h + 2 * D::CP, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                   // accumulator pairs (c0, c1) / (c2, c3): two adjacent columns
          const int row = r0 + 16 * mt + g + hh * 8, col = n0 + 8 * nt + 2 * t;
          if (col < ncol)                                    // col + 1 == ncol: a padding column of the weights, its accumulator is 0
            *reinterpret_cast<float2*>(sm.Vb + row * D::LDVB + col) = make_float2(acc[mt][nt][2 * hh] * inv, acc[mt][nt][2 * hh + 1] * inv);
        }
  }
  pair_sync(pair);
  for (int idx = pt; idx < PE * D::CP; idx += PT) {     // torch.linalg.cross(Vcp[:cp], Vcp[cp:]) appended after Vh
    const int e = e0 + idx / D::CP, j = idx % D::CP;
    float* bx = sm.Vb + (e * 3 + 0) * D::LDVB;
    float* by = bx + D::LDVB;
    float* bz = by + D::LDVB;
    const int ca = h + j, cb = h + D::CP + j;
    const float ax = bx[ca], ay = by[ca], az = bz[ca], qx = bx[cb], qy = by[cb], qz = bz[cb];
    bx[ca] = __fsub_rn(__fmul_rn(ay, qz), __fmul_rn(az, qy));
    by[ca] = __fsub_rn(__fmul_rn(az, qx), __fmul_rn(ax, qz));
    bz[ca] = __fsub_rn(__fmul_rn(ax, qy), __fmul_rn(ay, qx));
  }
  pair_sync(pair);
  for (int idx = pt; idx < PE * (VHW / 4); idx += PT) {
    const int e = e0 + idx / (VHW / 4), c4 = idx % (VHW / 4);
    const bool okr = e < nvalid;
    const float* b0 = sm.Vb + (e * 3) * D::LDVB + c4 * 4;
    const float4 pa = *reinterpret_cast<const float4*>(b0), pb = *reinterpret_cast<const float4*>(b0 + D::LDVB),
                 pc = *reinterpret_cast<const float4*>(b0 + 2 * D::LDVB);
    const float ra[4] = {pa.x, pa.y, pa.z, pa.w}, rb[4] = {pb.x, pb.y, pb.z, pb.w}, rc[4] = {pc.x, pc.y, pc.z, pc.w};
    float va[4], vb_[4], vc[4], nn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = okr && c4 * 4 + k < hc;
      va[k] = ok ? ra[k] : 0.f;
      vb_[k] = ok ? rb[k] : 0.f;
      vc[k] = ok ? rc[k] : 0.f;
      // gvp.py:14-21 sqrt(clamp(x^2 + y^2 + z^2, 1e-8)), square root within 1 ulp (sqrt_pos)
      const float q = __fadd_rn(__fadd_rn(__fmul_rn(va[k], va[k]), __fmul_rn(vb_[k], vb_[k])), __fmul_rn(vc[k], vc[k]));
      nn[k] = ok ? sqrt_pos(fmaxf(q, 1e-8f)) : 0.f;
    }
    float4* vh = reinterpret_cast<float4*>(VH + (erow0 + e) * 3 * VHW);
    vh[c4] = make_float4(va[0], va[1], va[2], va[3]);
    vh[VHW / 4 + c4] = make_float4(vb_[0], vb_[1], vb_[2], vb_[3]);
    vh[2 * (VHW / 4) + c4] = make_float4(vc[0], vc[1], vc[2], vc[3]);
    reinterpret_cast<float4*>(SH + (erow0 + e) * VHW)[c4] = make_float4(nn[0], nn[1], nn[2], nn[3]);
  }
}

// stage 2 of a GVP for the pair's 16 rows: Vb (loaded from VH) x Wu, gated by GT -> Va[.., 0:V).  No trailing barrier:
// vec_stage1 opens with the pair barrier, other consumers synchronise themselves.
template <class D>
__device__ __forceinline__ void vec_stage2(Smem<D>& sm, const int hc, const WH16& wu_sm, const size_t erow0,
                                           const int nvalid, const float* __restrict__ VH, const float* __restrict__ GT) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pair = warp >> 1, pt = tid & (PT - 1);
  const int e0 = pair * PE, r0 = e0 * 3;
  {
    // the pair's VH block (16 x 120 floats = 48 flat rows x 40) and GT block (16 x 32) are contiguous in memory: straight float4
    // streams, all loads of a thread in flight together; dead rows are zeroed
    constexpr int NV = PE * 3 * VHW / 4, NG = PE * 32 / 4, PV = (NV + PT - 1) / PT, PG = (NG + PT - 1) / PT;
    const float4* vsrc = reinterpret_cast<const float4*>(VH + (erow0 + e0) * 3 * VHW);
    const float4* gsrc = reinterpret_cast<const float4*>(GT + (erow0 + e0) * 32);
    float4 vb[PV], gb[PG];
#pragma unroll
    for (int i = 0; i < PV; ++i) { const int idx = pt + i * PT; if (idx < NV) vb[i] = vsrc[idx]; }
#pragma unroll
    for (int i = 0; i < PG; ++i) { const int idx = pt + i * PT; if (idx < NG) gb[i] = gsrc[idx]; }
    pair_sync(pair);                                     // the pair's previous readers of Vb / G / Va are done
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      const int idx = pt + i * PT;
      if (idx < NV) {
        const int rr = idx / (VHW / 4), c4 = idx - rr * (VHW / 4);
        *reinterpret_cast<float4*>(sm.Vb + (r0 + rr) * D::LDVB + c4 * 4) = e0 + rr / 3 < nvalid ? vb[i] : zero;
      }
    }
#pragma unroll
    for (int i = 0; i < PG; ++i) {
      const int idx = pt + i * PT;
      if (idx < NG) *reinterpret_cast<float4*>(sm.G + e0 * 32 + idx * 4) = e0 + (idx >> 3) < nvalid ? gb[i] : zero;
    }
  }
  pair_sync(pair);
  {
    // Vu = Vh_ext x Wu: 48 flat rows x 32 columns, K = hc padded to 8.  Each warp: 3 m16 tiles x 2 n8 tiles.
    const int n0 = (warp & 1) * 16, g = lane >> 2, t = lane & 3;
    float acc[3][2][4];
    warp_gemm_h16x3<3, 2>(sm.Vb, D::LDVB, r0, wu_sm.hi, wu_sm.lo, WLD_U, n0, (hc + 7) & ~7, acc);
    const float inv = wu_sm.inv;
    float2 gg[3][2][2];                                      // gates first: all 12 shared-memory loads in flight together
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = r0 + 16 * mt + g + hh * 8, col = n0 + 8 * nt + 2 * t;
          gg[mt][nt][hh] = *reinterpret_cast<const float2*>(sm.G + (row / 3) * 32 + col);
        }
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = r0 + 16 * mt + g + hh * 8, col = n0 + 8 * nt + 2 * t;
          *reinterpret_cast<float2*>(sm.Va + row * D::LDVA + col) =
              make_float2(__fmul_rn(gg[mt][nt][hh].x, acc[mt][nt][2 * hh] * inv), __fmul_rn(gg[mt][nt][hh].y, acc[mt][nt][2 * hh + 1] * inv));
        }
  }
}

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vec_a(const ModelRT m, const BatchRT bt, int layer, const float* __restrict__ x, const float* __restrict__ v,
        float* __restrict__ VH, float* __restrict__ SH) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x, pair = tid >> 6, pt = tid & (PT - 1), e0 = pair * PE;
  // [pad8(VIN0)][64 (+8)] words
  const WH16 w_hcp = load_resident_h16(sm.wstage, m.c(layer, C_MSG0_WHCP), pad4(D::VIN0), 32 * D::CPT_HC0, WLD_HCP, sm.dist);
  __syncthreads();
  for (int tile = blockIdx.x; tile < bt.n_edge_tiles; tile += gridDim.x) {
    const EdgeTile<D> et(bt, tile);
    if (pt < PE) {                                                // the pair's 16 edges: endpoints and the unit vector x_diff
      const int e = e0 + pt, le = et.le0 + e;
      int s = -1;
      float ux = 0.f, uy = 0.f, uz = 0.f;
      if (le < et.ecount) {
        int i, j;
        edge_src_dst(le, et.n, i, j);
        s = et.nb + i;
        float dx, dy, dz;
        const float dist = pair_dist(x, s, et.nb + j, dx, dy, dz);
        ux = __fdiv_rn(dx, dist); uy = __fdiv_rn(dy, dist); uz = __fdiv_rn(dz, dist);
      }
      sm.src[e] = s;
      sm.Va[(e * 3 + 0) * D::LDVA] = ux;
      sm.Va[(e * 3 + 1) * D::LDVA] = uy;
      sm.Va[(e * 3 + 2) * D::LDVA] = uz;
    }
    pair_sync(pair);
    {
      // v of the source node: 3 planes x V floats per edge, float4 gathers all in flight (v is L2 resident), then the K padding
      constexpr int Q4 = D::V / 4, NQ = PE * 3 * Q4, PER = (NQ + PT - 1) / PT;
      float4 buf[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = pt + i * PT;
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < NQ) {
          const int el = idx / (3 * Q4), rem = idx - el * (3 * Q4);          // rem = plane * Q4 + quad: contiguous in v[src]
          const int s = sm.src[e0 + el];
          if (s >= 0) buf[i] = __ldg(reinterpret_cast<const float4*>(v + (size_t)s * 3 * D::V) + rem);
        }
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = pt + i * PT;
        if (idx < NQ) {
          const int el = idx / (3 * Q4), rem = idx - el * (3 * Q4), p = rem / Q4, q = rem - p * Q4;
          float* d = sm.Va + ((e0 + el) * 3 + p) * D::LDVA + 1 + q * 4;
          d[0] = buf[i].x; d[1] = buf[i].y; d[2] = buf[i].z; d[3] = buf[i].w;
        }
      }
      constexpr int PADW = D::LDVA - 1 - D::V;
      for (int idx = pt; idx < PE * 3 * PADW; idx += PT) {
        const int fr = idx / PADW, c = idx - fr * PADW;
        sm.Va[(e0 * 3 + fr) * D::LDVA + 1 + D::V + c] = 0.f;
      }
    }
    vec_stage1<D>(sm, D::VIN0, D::H0, w_hcp, et.erow0, et.nvalid, VH, SH);
  }
}

// stage 2 of one GVP (Wu [hc_prev][32]) then stage 1 of the next (Whcp [V][64]); rows are padded edge slots, or nodes
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vec_b(const BatchRT bt, const float* __restrict__ wu, int hc_prev, const float* __restrict__ whcp, int node_rows,
        float* __restrict__ VH, float* __restrict__ SH, const float* __restrict__ GT) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  // [pad8(hc_prev)][32 (+8)] words, then [V][64 (+8)] words
  const WH16 w_u = load_resident_h16(sm.wstage, wu, pad4(hc_prev), 32, WLD_U, sm.dist);
  const WH16 w_hcp = load_resident_h16(sm.wstage + 40 * WLD_U, whcp, D::V, 32 * D::CPT_HC, WLD_HCP, sm.dist);
  __syncthreads();
  const int n_tiles = node_rows ? bt.n_node_tiles : bt.n_edge_tiles;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t erow0 = (size_t)tile * TM;
    const int nvalid = node_rows ? min(TM, bt.N - tile * TM) : EdgeTile<D>(bt, tile).nvalid;
    if (tile + (int)gridDim.x < n_tiles) prefetch_pair_inputs(VH, GT, (size_t)(tile + gridDim.x) * TM);
    vec_stage2<D>(sm, hc_prev, w_u, erow0, nvalid, VH, GT);
    vec_stage1<D>(sm, D::V, D::V, w_hcp, erow0, nvalid, VH, SH);
  }
}

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vec_c(const ModelRT m, const BatchRT bt, int layer, int first_col /* S when k_egemm_tc<EG_MSGA> reduced the scalar columns */,
        const float* __restrict__ VH, const float* __restrict__ GT, const float* __restrict__ Smsg, float* __restrict__ M,
        float* __restrict__ partF, float* __restrict__ partL) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x;
  const WH16 w_u = load_resident_h16(sm.wstage, m.c(layer, C_MSG2_WU), pad4(D::V + D::CP), 32, WLD_U, sm.dist);
  for (int tile = blockIdx.x; tile < bt.n_edge_tiles; tile += gridDim.x) {
    const EdgeTile<D> et(bt, tile);
    __syncthreads();                                              // previous tile's segment-sum readers of Va / dst are done
    if (tid < TM) {
      const int le = et.le0 + tid;
      sm.dst[tid] = le < et.ecount ? et.nb + le / (et.n - 1) : -1;
    }
    if (tile + (int)gridDim.x < bt.n_edge_tiles) prefetch_pair_inputs(VH, GT, (size_t)(tile + gridDim.x) * TM);
    vec_stage2<D>(sm, D::V + D::CP, w_u, et.erow0, et.nvalid, VH, GT);
    __syncthreads();
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
          for (int k = 0; k < 16; ++k) buf[k] = sm.Va[((r0 + k) * 3 + p) * D::LDVA + c];
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
// vector half on the tile in Va: v / (sqrt(mean_c clamp(|v_c|^2, 1e-8) + eps) + eps); the result is also stored to v[g].
// Warp w normalises rows 8w .. 8w+7, i.e. a warp pair covers exactly its own 16 rows.
template <class D>
__device__ __forceinline__ void tile_vec_layernorm(float* __restrict__ Va, float* __restrict__ v, int g0, int N) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < RPW; ++r) {
    const int row = warp * RPW + r, g = g0 + row;
    float* va = Va + (row * 3) * D::LDVA + lane;
    float a = 0.f, b = 0.f, c = 0.f, vq = 0.f;
    if (lane < D::V) {
      a = va[0]; b = va[D::LDVA]; c = va[2 * D::LDVA];
      vq = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)), 1e-8f);
    }
    const float vn = __fadd_rn(sqrtf(__fadd_rn(warp_sum(vq) * (1.0f / D::V), 1e-5f)), 1e-5f);
    if (lane < D::V) {
      a = __fdiv_rn(a, vn); b = __fdiv_rn(b, vn); c = __fdiv_rn(c, vn);
      va[0] = a; va[D::LDVA] = b; va[2 * D::LDVA] = c;
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
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const WH16 w_hcp = load_resident_h16(sm.wstage, m.c(layer, C_UPD0_WHCP), D::V, 32 * D::CPT_HC, WLD_HCP, sm.dist);
  const int g0 = blockIdx.x * TM;
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
  // The aggregated message of a node is M[g] when its in-edge segment lies in one aggregation tile, else partL[t0] + partF[t0 + 1]
  // (+ partF[t0 + 2] ... for molecules with more than ~64 atoms), summed in tile order.  The first two pieces are fetched
  // branch-free (a missing second piece is -0.0f: x + (-0.0f) == x bit for bit), so that all loads of a group of rows are in
  // flight together; the rare further pieces are added by a fix-up loop.  (The row-by-row version with a data-dependent loop per
  // element exposed one L2 round trip per piece and column: 235 us for 24 k nodes, profiles/r01q.)
  auto piece0 = [&](int t0, int t1, int g, int col) { return t0 == t1 ? M[(size_t)g * D::MW + col] : partL[(size_t)t0 * D::MW + col]; };
  auto piece1 = [&](int t0, int t1, int col) { return t1 > t0 ? partF[(size_t)(t0 + 1) * D::MW + col] : -0.0f; };
  // third piece, branch-free as well: with 32-row aggregation pieces (gate-fused message linears) a GEOM-sized node (45 in-edges)
  // spans two or three pieces, and the data-dependent loop below was k_node_pre's critical path (101 -> 182 us, profiles/r02f)
  auto piece2 = [&](int t0, int t1, int col) { return t1 > t0 + 1 ? partF[(size_t)(t0 + 2) * D::MW + col] : -0.0f; };
  auto finish = [&](int row, int t0, int t1, int col, float msg) {
    for (int t = t0 + 3; t <= t1; ++t) msg = __fadd_rn(msg, partF[(size_t)t * D::MW + col]);
    const float div = sm.dist[row];
    return div != 0.f ? __fdiv_rn(msg, div) : msg;
  };
  constexpr int RG = 2;                                     // rows per group: 2 x 8 columns x 3 loads in flight per lane
  const int lane = tid & 31;
  for (int r0 = 0; r0 < RPW; r0 += RG) {
    float sv[RG][D::CPT_S], a0[RG][D::CPT_S], a1[RG][D::CPT_S], a2[RG][D::CPT_S];
#pragma unroll
    for (int rr = 0; rr < RG; ++rr) {
      const int row = warp * RPW + r0 + rr, g = min(g0 + row, bt.N - 1);      // dead rows read a live row and are not stored
      const int t0 = sm.dst[row], t1 = sm.aux[row];
#pragma unroll
      for (int c = 0; c < D::CPT_S; ++c) {
        const int col = lane + 32 * c;
        sv[rr][c] = s[(size_t)g * D::S + col];
        a0[rr][c] = piece0(t0, t1, g, col);
        a1[rr][c] = piece1(t0, t1, col);
        a2[rr][c] = piece2(t0, t1, col);
      }
    }
#pragma unroll
    for (int rr = 0; rr < RG; ++rr) {
      const int row = warp * RPW + r0 + rr, g = g0 + row;
      if (g >= bt.N) break;
      const int t0 = sm.dst[row], t1 = sm.aux[row];
      float* srow = s + (size_t)g * D::S;
      row_scalar_layernorm<D>(srow, m.c(layer, C_LN_MSG_W), m.c(layer, C_LN_MSG_B), [&](int col) {
        const int c = col >> 5;
        return __fadd_rn(sv[rr][c], finish(row, t0, t1, col, __fadd_rn(__fadd_rn(a0[rr][c], a1[rr][c]), a2[rr][c])));
      });
    }
  }
  for (int i0 = tid; i0 < TM * 3 * D::V; i0 += 4 * NT) {
    float vv[4], b0[4], b1[4], b2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                            // 4 x 3 loads in flight
      const int idx = i0 + k * NT, row = idx / (3 * D::V), pc = idx - row * 3 * D::V, g = min(g0 + row, bt.N - 1);
      const int t0 = sm.dst[row], t1 = sm.aux[row];
      vv[k] = v[(size_t)g * 3 * D::V + pc];
      b0[k] = piece0(t0, t1, g, D::S + pc);
      b1[k] = piece1(t0, t1, D::S + pc);
      b2[k] = piece2(t0, t1, D::S + pc);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = i0 + k * NT, row = idx / (3 * D::V), pc = idx - row * 3 * D::V, g = g0 + row;
      float val = 0.f;
      if (g < bt.N) val = __fadd_rn(vv[k], finish(row, sm.dst[row], sm.aux[row], D::S + pc, __fadd_rn(__fadd_rn(b0[k], b1[k]), b2[k])));
      sm.Va[(row * 3 + pc / D::V) * D::LDVA + (pc % D::V)] = val;
    }
  }
  __syncthreads();
  tile_vec_layernorm<D>(sm.Va, v, g0, bt.N);
  __syncthreads();
  vec_stage1<D>(sm, D::V, D::V, w_hcp, (size_t)g0, min(TM, bt.N - g0), VH, SH);
}

// k_node_mid: stage 2 of the last update GVP, (s, v) <- GVPLayerNorm((s, v) + update)  (gvp.py:515-519), then stage 1 of the first
// NodePositionUpdate GVP when this conv is followed by a molecule update
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_node_mid(const ModelRT m, const BatchRT bt, int layer, int updater, float* __restrict__ s, float* __restrict__ v,
           const float* __restrict__ S3, float* __restrict__ VH, float* __restrict__ SH, const float* __restrict__ GT) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const WH16 w_u = load_resident_h16(sm.wstage, m.c(layer, C_UPD2_WU), pad4(D::V + D::CP), 32, WLD_U, sm.dist);
  WH16 w_hcp{nullptr, nullptr, 1.f};
  if (updater >= 0) w_hcp = load_resident_h16(sm.wstage + 40 * WLD_U, m.u(updater, U_POS0_WHCP), D::V, 32 * D::CPT_HC, WLD_HCP, sm.dist);
  const int g0 = blockIdx.x * TM, nvalid = min(TM, bt.N - g0);
  __syncthreads();
  vec_stage2<D>(sm, D::V + D::CP, w_u, (size_t)g0, nvalid, VH, GT);
  for (int r = 0; r < RPW; ++r) {
    const int g = g0 + warp * RPW + r;
    if (g >= bt.N) break;
    float* srow = s + (size_t)g * D::S;
    const float* urow = S3 + (size_t)g * D::S;
    row_scalar_layernorm<D>(srow, m.c(layer, C_LN_UPD_W), m.c(layer, C_LN_UPD_B),
                            [&](int col) { return __fadd_rn(srow[col], urow[col]); });
  }
  __syncthreads();
  for (int idx = tid; idx < TM * 3 * D::V; idx += NT) {
    const int row = idx / (3 * D::V), pc = idx - row * 3 * D::V, g = g0 + row;
    float* p = sm.Va + (row * 3 + pc / D::V) * D::LDVA + (pc % D::V);
    *p = g < bt.N ? __fadd_rn(v[(size_t)g * 3 * D::V + pc], *p) : 0.f;
  }
  __syncthreads();
  tile_vec_layernorm<D>(sm.Va, v, g0, bt.N);
  __syncthreads();
  if (updater >= 0) vec_stage1<D>(sm, D::V, D::V, w_hcp, (size_t)g0, nvalid, VH, SH);
}

// k_node_post: stage 2 of the last position GVP (one output vector); x += that vector   (vector_field.py:813-842)
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_node_post(const ModelRT m, const BatchRT bt, int updater, float* __restrict__ x, const float* __restrict__ VH,
            const float* __restrict__ GT) {
  pdl_launch();
  pdl_wait();
  extern __shared__ __align__(16) float smem_raw[];
  Smem<D> sm = VecSmem<D>::carve(smem_raw);
  const int tid = threadIdx.x;
  const WH16 w_u = load_resident_h16(sm.wstage, m.u(updater, U_POS2_WU), pad4(D::V + D::CP), 32, WLD_U, sm.dist);
  const int g0 = blockIdx.x * TM;
  __syncthreads();
  vec_stage2<D>(sm, D::V + D::CP, w_u, (size_t)g0, min(TM, bt.N - g0), VH, GT);
  __syncthreads();
  if (tid < TM * 3) {
    const int row = tid / 3, p = tid - row * 3, g = g0 + row;
    if (g < bt.N) x[g * 3 + p] = __fadd_rn(x[g * 3 + p], sm.Va[tid * D::LDVA]);      // flat row = row * 3 + p = tid
  }
}

}  // namespace fm
