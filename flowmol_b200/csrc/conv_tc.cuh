// k_conv_edge_tc -- the hot kernel on the 5th-generation tensor cores (tcgen05 / TMEM / bulk-TMA weight stream).
//
// Same contract as k_conv_edge (kernels.cuh): one tile of consecutive in-edges (dst-major) of one molecule, gather of
// src/dst node data, x_diff / rbf recomputed in-kernel, three chained message GVPs, deterministic segment-sum per dst.
// What moves to the tensor cores: the three 197/292 -> 256 scalar linears and the three 256 -> 32 gate linears
// (95 % of the kernel's MACs), as error-compensated 3xTF32 (tc.cuh) with fp32 accumulation in TMEM.
//
// Orientation ("features on M"): D^T[f][e] = sum_k W[f][k] * X[e][k]
//   UMMA A operand = weight tile  [128 features][32 k]  (pre-swizzled hi / lo images packed by weights.py:tc_units,
//                                                        streamed L2 -> smem by cp.async.bulk into a 4-slot ring)
//   UMMA B operand = activations  [32 edges][32 k] per k-slab, hi and lo, written by the producing stage of this CTA
//   accumulator    = TMEM lanes = features, columns = edges: the epilogue thread of feature f sees all 32 edges of its
//                    feature, so bias + SiLU + the hand-over to the next layer's B operand (one k-slab per warp) and the
//                    final segment-sum over edges are thread-local -- no shared-memory transpose, no atomics.
// The small vector-channel GEMMs (33x41, 37x32 per plane) stay on the fp32 CUDA-core tile_gemm path.
//
// Warp roles inside tc_gemm: warp 0 lane 0 = weight producer (bulk copies, mbarrier expect_tx), warp 1 lane 0 = MMA
// issuer (tcgen05.mma + tcgen05.commit to free ring slots / publish the accumulator); all 8 warps run the epilogues.
#pragma once
#include "kernels.cuh"
#include "tc.cuh"

namespace fm {

constexpr int TCT = 32;             // edges per tile = UMMA N
constexpr int TC_RING = 4;          // weight ring slots
constexpr int TC_UNIT = 16384;      // bytes per slot: one [128][32] fp32 operand tile
constexpr int TC_XSLAB = TCT * 128; // bytes per activation k-slab (hi or lo)
constexpr int TC_RPW = TCT / NWARP; // rows per warp on the CUDA-core vector path

template <class D>
struct TcPlan {
  static constexpr int NSLAB = (D::K1 + 31) / 32;                    // 10 at flowmol3 (K = 292)
  static constexpr int LDVA = D::LDVA, LDVB = D::LDVB;
  static constexpr int OFF_XHI = 0;
  static constexpr int OFF_XLO = OFF_XHI + NSLAB * TC_XSLAB;
  static constexpr int OFF_RING = OFF_XLO + NSLAB * TC_XSLAB;        // multiple of 1024
  static constexpr int OFF_VA = OFF_RING + TC_RING * TC_UNIT;
  static constexpr int OFF_VB = OFF_VA + 3 * TCT * LDVA * 4;
  static constexpr int OFF_G = OFF_VB + 3 * TCT * LDVB * 4;
  static constexpr int OFF_WST = OFF_G + TCT * 32 * 4;
  static constexpr int OFF_MISC = OFF_WST + 2 * KC * 64 * 4;         // vector GEMMs have <= 64 padded columns
  static constexpr int OFF_BAR = OFF_MISC + 4 * TCT * 4;
  static constexpr int BYTES = OFF_BAR + (2 * TC_RING + 2) * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES + 1024;                 // + slack for the manual 1024-byte alignment
  static_assert(OFF_RING % 1024 == 0 && OFF_XLO % 1024 == 0, "UMMA operand tiles need 1024-byte alignment");
};

struct TcGemm {
  const float* units;   // consecutive units in consumption order (weights.py:tc_units)
  int n_slab, n_mt, unit_bytes, last_ksteps;
};

struct TcCtx {
  uint8_t *xhi, *xlo, *ring;
  uint64_t *full, *empty, *accbar;
  uint32_t tmem;        // TMEM base (lane 0, column 0 of the allocation)
  uint32_t g;           // global unit counter (ring slot / phase bookkeeping), identical in every thread
  uint32_t acc_phase;
  int dbg;              // timing experiments only: 1 = no weight copies, 2 = no MMA issue, 4 = no vector stages, 8 = no x_store
};

// C^T[f][e] (TMEM columns [col0 + 32*m, +32) for m-tile m) = sum over k-slabs of W-units x X-slabs, 3xTF32.
// Precondition: the X slabs were written by this CTA's threads, each followed by tc::fence_proxy_async().
__device__ __forceinline__ void tc_gemm(const TcGemm w, TcCtx& cx, uint32_t col0) {
  tc::tc_fence_before();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nunits = (uint32_t)(w.n_slab * w.n_mt * 2);
  if (warp == 0 && lane == 0) {                       // ---- weight producer ------------------------------------------
    for (uint32_t u = 0; u < nunits; ++u) {
      const uint32_t gi = cx.g + u, slot = gi % TC_RING, use = gi / TC_RING;
      if (use > 0) tc::mbar_wait(&cx.empty[slot], (use - 1) & 1);
      if (cx.dbg & 1) { tc::mbar_arrive_expect_tx(&cx.full[slot], 0u); continue; }
      tc::mbar_arrive_expect_tx(&cx.full[slot], (uint32_t)w.unit_bytes);
      tc::bulk_g2s(cx.ring + slot * TC_UNIT, reinterpret_cast<const uint8_t*>(w.units) + (size_t)u * w.unit_bytes,
                   (uint32_t)w.unit_bytes, &cx.full[slot]);
    }
  } else if (warp == 1 && lane == 0) {                // ---- MMA issuer ------------------------------------------------
    tc::tc_fence_after();
    const uint32_t idesc = tc::idesc_tf32(128, TCT);
    uint32_t u = 0;
    for (int s = 0; s < w.n_slab; ++s) {
      const int ksteps = (s == w.n_slab - 1) ? w.last_ksteps : 4;
      const uint32_t xh = tc::smem_u32(cx.xhi + s * TC_XSLAB), xl = tc::smem_u32(cx.xlo + s * TC_XSLAB);
      for (int m = 0; m < w.n_mt; ++m) {
        const uint32_t d = cx.tmem + col0 + (uint32_t)(m * TCT);
        {   // hi weights: x_hi*w_hi + x_lo*w_hi
          const uint32_t gi = cx.g + u, slot = gi % TC_RING, use = gi / TC_RING;
          tc::mbar_wait(&cx.full[slot], use & 1);
          tc::tc_fence_after();
          const uint32_t wb = tc::smem_u32(cx.ring + slot * TC_UNIT);
          for (int j = 0; j < ksteps && !(cx.dbg & 2); ++j) {
            const uint64_t dw = tc::desc_sw128(wb + 32 * j);
            tc::umma_tf32(d, dw, tc::desc_sw128(xl + 32 * j), idesc, (s > 0 || j > 0) ? 1u : 0u);
            tc::umma_tf32(d, dw, tc::desc_sw128(xh + 32 * j), idesc, 1u);
          }
          tc::umma_commit(&cx.empty[slot]);
          ++u;
        }
        {   // lo weights: x_hi*w_lo
          const uint32_t gi = cx.g + u, slot = gi % TC_RING, use = gi / TC_RING;
          tc::mbar_wait(&cx.full[slot], use & 1);
          tc::tc_fence_after();
          const uint32_t wb = tc::smem_u32(cx.ring + slot * TC_UNIT);
          for (int j = 0; j < ksteps && !(cx.dbg & 2); ++j)
            tc::umma_tf32(d, tc::desc_sw128(wb + 32 * j), tc::desc_sw128(xh + 32 * j), idesc, 1u);
          tc::umma_commit(&cx.empty[slot]);
          ++u;
        }
      }
    }
    tc::umma_commit(cx.accbar);
  }
  cx.g += nunits;
  tc::mbar_wait(cx.accbar, cx.acc_phase & 1);
  ++cx.acc_phase;
  tc::tc_fence_after();
}

// write one activation value (edge row e, feature / k index kx) into the hi / lo B-operand slabs
__device__ __forceinline__ void x_store(const TcCtx& cx, int e, int kx, float v) {
  float hi, lo;
  tc::split_tf32(v, hi, lo);
  const uint32_t off = (uint32_t)(kx >> 5) * TC_XSLAB + tc::sw128_off(e, kx & 31);
  *reinterpret_cast<float*>(cx.xhi + off) = hi;
  *reinterpret_cast<float*>(cx.xlo + off) = lo;
}

// deterministic segment-sum bookkeeping shared by the scalar (register) and vector (shared-memory) message columns
struct SegCtx {
  const int* dst;       // [TCT] global dst node of each row, -1 = padding row
  int nb, n, le0, tile;
  float *M, *partF, *partL;
  int MW;
};
__device__ __forceinline__ void seg_flush(const SegCtx& sg, float acc, int col, int d, int seg_first, int le_last) {
  const int j = d - sg.nb;
  const bool head = seg_first == j * (sg.n - 1), tail = le_last == j * (sg.n - 1) + (sg.n - 2);
  if (head && tail) sg.M[(size_t)d * sg.MW + col] = acc;
  else if (head) sg.partL[(size_t)sg.tile * sg.MW + col] = acc;
  else sg.partF[(size_t)sg.tile * sg.MW + col] = acc;
}

template <class D>
__global__ void __launch_bounds__(NT, 1)
k_conv_edge_tc(const ModelRT m, const BatchRT bt, int layer, const float* __restrict__ x, const float* __restrict__ v,
               const float* __restrict__ ef, const float* __restrict__ P, float* __restrict__ M, float* __restrict__ partF,
               float* __restrict__ partL, int dbg) {
  pdl_launch();
  pdl_wait();
  using PL = TcPlan<D>;
  static_assert(D::S == 256 && D::V == 32 && D::SD == 0, "k_conv_edge_tc is specialised for the flowmol3 dimensions");
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  float* Va = reinterpret_cast<float*>(base + PL::OFF_VA);
  float* Vb = reinterpret_cast<float*>(base + PL::OFF_VB);
  float* G = reinterpret_cast<float*>(base + PL::OFF_G);
  float* wstage = reinterpret_cast<float*>(base + PL::OFF_WST);
  int* s_src = reinterpret_cast<int*>(base + PL::OFF_MISC);
  int* s_dst = s_src + TCT;
  float* s_dist = reinterpret_cast<float*>(s_dst + TCT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + PL::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_RING + 1);
  TcCtx cx;
  cx.xhi = base + PL::OFF_XHI; cx.xlo = base + PL::OFF_XLO; cx.ring = base + PL::OFF_RING;
  cx.full = bars; cx.empty = bars + TC_RING; cx.accbar = bars + 2 * TC_RING;
  cx.g = 0; cx.acc_phase = 0; cx.dbg = dbg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 32-row tiles are the halves of the 64-slot storage tiles of the batch descriptor
  const int tile = blockIdx.x, tile64 = tile >> 1, mol = bt.etile_mol[tile64];
  const int n = bt.mol_n[mol], nb = bt.mol_node[mol], ecount = n * (n - 1);
  const int le0 = (tile64 - bt.mol_etile[mol]) * TM + (tile & 1) * TCT;
  if (le0 >= ecount) return;                                        // second half of a short last tile: nothing to do
  const size_t erow0 = (size_t)tile64 * TM + (tile & 1) * TCT;
  // ---- one-time setup: barriers, TMEM, zeroed activation slabs ------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < TC_RING; ++i) { tc::mbar_init(&cx.full[i], 1); tc::mbar_init(&cx.empty[i], 1); }
    tc::mbar_init(cx.accbar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 128);
  for (int i = tid; i < 2 * PL::NSLAB * TC_XSLAB / 16; i += NT) reinterpret_cast<float4*>(cx.xhi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // ---- gather + geometric features ---------------------------------------------------------------------------------------------
  if (tid < TCT) {
    const int le = le0 + tid;
    int s = -1, d = -1;
    float dist = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
    if (le < ecount) {
      int i, j;
      edge_src_dst(le, n, i, j);
      s = nb + i; d = nb + j;
      float dx, dy, dz;
      dist = pair_dist(x, s, d, dx, dy, dz);
      ux = __fdiv_rn(dx, dist); uy = __fdiv_rn(dy, dist); uz = __fdiv_rn(dz, dist);
    }
    s_src[tid] = s; s_dst[tid] = d; s_dist[tid] = dist;
    Va[(0 * TCT + tid) * PL::LDVA] = ux;
    Va[(1 * TCT + tid) * PL::LDVA] = uy;
    Va[(2 * TCT + tid) * PL::LDVA] = uz;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cx.tmem = *tmem_slot;
  {
    const float* mu = m.g(G_RBF_MU);
    const float sigma = m.rbf_dmax / (float)D::R;
    for (int idx = tid; idx < TCT * D::R; idx += NT) {                 // k [0, R): rbf(d)            -> slab 0
      const int row = idx / D::R, k = idx - row * D::R;
      if (s_src[row] >= 0) x_store(cx, row, k, rbf_f(s_dist[row], mu[k], sigma));
    }
    for (int idx = tid; idx < TCT * (D::F / 4); idx += NT) {           // k [R, R+F): edge features   -> slabs 1..4
      const int row = idx / (D::F / 4), c4 = idx - row * (D::F / 4);
      if (s_src[row] >= 0) {
        const float4 val = __ldg(reinterpret_cast<const float4*>(ef + (erow0 + row) * D::F) + c4);
        const int kx = D::R + c4 * 4;
        float4 hi, lo;
        tc::split_tf32(val.x, hi.x, lo.x); tc::split_tf32(val.y, hi.y, lo.y);
        tc::split_tf32(val.z, hi.z, lo.z); tc::split_tf32(val.w, hi.w, lo.w);
        const uint32_t off = (uint32_t)(kx >> 5) * TC_XSLAB + tc::sw128_off(row, kx & 31);
        *reinterpret_cast<float4*>(cx.xhi + off) = hi;
        *reinterpret_cast<float4*>(cx.xlo + off) = lo;
      }
    }
    constexpr int VW = PL::LDVA - 1;                                    // vector cols [1, LDVA): v_src | zero pad
    for (int idx = tid; idx < 3 * TCT * VW; idx += NT) {
      const int pr = idx / VW, c = idx - pr * VW;
      const int p = pr / TCT, row = pr - p * TCT;
      float val = 0.f;
      const int s = s_src[row];
      if (s >= 0 && c < D::V) val = v[((size_t)s * 3 + p) * D::V + c];
      Va[pr * PL::LDVA + 1 + c] = val;
    }
  }
  SegCtx sg{s_dst, nb, n, le0, tile, M, partF, partL, D::MW};
  // ---- three message GVPs ----------------------------------------------------------------------------------------------------------
  for (int gi = 0; gi < 3; ++gi) {
    const int v_in = gi == 0 ? D::VIN0 : D::V, h = gi == 0 ? D::H0 : D::V, s_in = gi == 0 ? D::R + D::F : D::S;
    const int hc = h + D::CP;
    const GvpPtr w = gvp_ptr_conv(m, layer, gi == 0 ? C_MSG0_WHCP : (gi == 1 ? C_MSG1_WHCP : C_MSG2_WHCP));
    // -- vector stage 1 (CUDA cores): [Vh | Vcp] = V x [Wh | Wcp], cross products, norms -> k-slabs of the scalar operand
    if (!(dbg & 4)) {
      float acc[3][TC_RPW][D::CPT_HC0];
      tile_gemm<3, D::CPT_HC0, TC_RPW, 2 * KC * 64>(Va, PL::LDVA, TCT * PL::LDVA, pad4(v_in), w.whcp, wstage, acc);
      const int ncol = h + 2 * D::CP;
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int r = 0; r < TC_RPW; ++r)
#pragma unroll
          for (int c = 0; c < D::CPT_HC0; ++c) {
            const int col = ColMap<D::CPT_HC0>::col(lane, c);
            if (col < ncol) Vb[(p * TCT + warp * TC_RPW + r) * PL::LDVB + col] = acc[p][r][c];
          }
    }
    __syncthreads();
    for (int idx = tid; idx < TCT * D::CP; idx += NT) {
      const int row = idx / D::CP, j = idx - row * D::CP;
      float* bx = Vb + (0 * TCT + row) * PL::LDVB;
      float* by = Vb + (1 * TCT + row) * PL::LDVB;
      float* bz = Vb + (2 * TCT + row) * PL::LDVB;
      const int ca = h + j, cb = h + D::CP + j;
      const float ax = bx[ca], ay = by[ca], az = bz[ca], qx = bx[cb], qy = by[cb], qz = bz[cb];
      bx[ca] = __fsub_rn(__fmul_rn(ay, qz), __fmul_rn(az, qy));
      by[ca] = __fsub_rn(__fmul_rn(az, qx), __fmul_rn(ax, qz));
      bz[ca] = __fsub_rn(__fmul_rn(ax, qy), __fmul_rn(ay, qx));
    }
    __syncthreads();
    {
      const int hcp = pad4(hc);
      for (int idx = tid; idx < TCT * hcp; idx += NT) {
        const int row = idx / hcp, c = idx - row * hcp;
        if (c < hc) {
          const float a = Vb[(0 * TCT + row) * PL::LDVB + c], b = Vb[(1 * TCT + row) * PL::LDVB + c], cc = Vb[(2 * TCT + row) * PL::LDVB + c];
          x_store(cx, row, s_in + c, s_src[row] >= 0 ? norm_no_nan3(a, b, cc) : 0.f);
        } else {
          Vb[(0 * TCT + row) * PL::LDVB + c] = 0.f; Vb[(1 * TCT + row) * PL::LDVB + c] = 0.f; Vb[(2 * TCT + row) * PL::LDVB + c] = 0.f;
        }
      }
    }
    tc::fence_proxy_async();
    // -- scalar path on the tensor cores: s' = SiLU(W [s | sh] + pre) ------------------------------------------------------------------
    const int K = s_in + hc;
    TcGemm gm{m.c(layer, (gi == 0 ? C_MSG0_TCW : (gi == 1 ? C_MSG1_TCW : C_MSG2_TCW))), (K + 31) / 32, 2, TC_UNIT,
              ((K - 1) % 32) / 8 + 1};
    tc_gemm(gm, cx, 0);
    {
      const int mt = warp >> 2, q = warp & 3, f = mt * 128 + q * 32 + lane;
      float acc[32];
      tc::tmem_ld32(cx.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * TCT), acc);
      tc::tmem_ld_wait();
      const float bias = gi == 0 ? 0.f : w.b[f];
      float run = 0.f;
      int seg_first = le0;
#pragma unroll
      for (int e = 0; e < TCT; ++e) {
        const int s = s_src[e];
        float pre = bias;
        if (gi == 0) pre = s >= 0 ? P[(size_t)s * D::S + f] : 0.f;     // per-node half of the first linear (bias folded in)
        const float val = silu_f(acc[e] + pre);
        if (!(dbg & 8)) x_store(cx, e, f, val);                          // next layer's / the gate GEMM's operand, k = f
        if (gi == 2 && s >= 0) {                                         // segment-sum of the scalar message, thread-local
          const int d = s_dst[e];
          run = __fadd_rn(run, val);
          if (e == TCT - 1 || s_dst[e + 1] != d) {
            seg_flush(sg, run, f, d, seg_first, le0 + e);
            run = 0.f;
            seg_first = le0 + e + 1;
          }
        }
      }
    }
    tc::fence_proxy_async();
    // -- gates on the tensor cores: g = sigmoid(Wg s' + bg) ------------------------------------------------------------------------------
    TcGemm gg{m.c(layer, (gi == 0 ? C_MSG0_TCG : (gi == 1 ? C_MSG1_TCG : C_MSG2_TCG))), D::S / 32, 1, 32 * 128, 4};
    tc_gemm(gg, cx, 64);
    if (warp == 0) {
      float acc[32];
      tc::tmem_ld32(cx.tmem + 64u, acc);
      tc::tmem_ld_wait();
      const float bg = w.bg[lane];
#pragma unroll
      for (int e = 0; e < TCT; ++e) G[e * 32 + lane] = sigmoid_f(acc[e] + bg);
    }
    // -- vector stage 2 (CUDA cores): V' = gate * (Vh_ext x Wu) --------------------------------------------------------------------------
    if (!(dbg & 4)) {
      float acc[3][TC_RPW][1];
      tile_gemm<3, 1, TC_RPW, 2 * KC * 64>(Vb, PL::LDVB, TCT * PL::LDVB, pad4(hc), w.wu, wstage, acc);    // entry barrier publishes G
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int r = 0; r < TC_RPW; ++r) {
          const int row = warp * TC_RPW + r;
          Va[(p * TCT + row) * PL::LDVA + lane] = __fmul_rn(G[row * 32 + lane], acc[p][r][0]);
        }
    }
    __syncthreads();
  }
  // ---- segment-sum of the vector message (columns S .. S+3V) -----------------------------------------------------------------------------
  if (tid < 3 * D::V) {
    const int p = tid / D::V, c = tid - p * D::V;
    float run = 0.f;
    int seg_first = le0;
    for (int e = 0; e < TCT; ++e) {
      const int d = s_dst[e];
      if (d < 0) break;
      run = __fadd_rn(run, Va[(p * TCT + e) * PL::LDVA + c]);
      if (e == TCT - 1 || s_dst[e + 1] != d) {
        seg_flush(sg, run, D::S + tid, d, seg_first, le0 + e);
        run = 0.f;
        seg_first = le0 + e + 1;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(cx.tmem, 128);
}

}  // namespace fm
