// k_egemm_tc -- per-edge dense layers on the 5th-generation tensor cores, 256 edges per CTA.
//
//   OUT[e][f] = act( sum_k W[f][k] * IN[e][k] + pre(e, f) )            error-compensated 3xTF32, fp32 accumulation in TMEM
//
// The fused single-CTA chain (conv_tc.cuh) is limited by what fits next to its operands in 227 KB: 32-edge tiles leave
// tcgen05.mma at its ~64-cycle-per-instruction floor (N = 32 costs the same as N = 128, measured: profiles/r01b) and stream
// 1.9 MB of weight images per 32 edges.  Here the three scalar linears and the three gate linears of a message pass run as
// wide tiles instead: N = 2 x 128 edges per CTA (full-rate MMAs, weight images amortised over 256 edges), activations
// round-trip through HBM between the stages (1 KB per edge) and the small vector-channel stages run in their own
// CUDA-core kernels (vec_stages.cuh).
//
// Orientation "features on M": UMMA A = weight unit [128 features][32 k] (bulk-TMA ring, pre-swizzled hi / lo images),
// UMMA B = activation k-slab [128 edges][32 k] x 2 halves, converted fp32 -> (hi, lo) SW128 images by the loader warps,
// D in TMEM: lane = feature, column = edge => epilogue stores are coalesced over features, bias is per thread.
//
// Warp roles (10 warps): 0 = weight producer, 1 = MMA issuer, 2..9 = activation loaders (one edge row per thread), then
// the same 8 warps run the epilogue (TMEM lane quarter = warp % 4, edge half = (warp - 2) / 4).
#pragma once
#include "conv_tc.cuh"

namespace fm {

// NH = 128-edge halves per CTA.  NH = 2: 256 edges, 1 CTA / SM (weight images amortised over 256 edges).
// NH = 1: 128 edges, 6 warps, 2 CTAs / SM -- the MMAs of one CTA overlap the prologue / epilogue of the other.
template <int NH>
struct EgPlan {
  static constexpr int T = 128 * NH;
  static constexpr int THREADS = 64 + 128 * NH;
  static constexpr int RING_BYTES = (NH == 2 ? 6 : 3) * TC_UNIT;   // weight ring
  static constexpr int MAX_SLOTS = 12;                             // barriers reserved (small gate units use more slots)
  static constexpr int XSTAGE = 32768 * NH;               // hi halves | lo halves, 16 KB each
  static constexpr int OFF_X = 0;
  static constexpr int OFF_RING = 2 * XSTAGE;
  static constexpr int OFF_ROW = OFF_RING + RING_BYTES;       // int src[T]; int16 (dst - src)[T]
  static constexpr int OFF_BAR = OFF_ROW + T * 6;
  static_assert(NH == 2 || OFF_BAR + (2 * MAX_SLOTS + 5) * 8 + 16 <= 115712, "2 CTAs per SM need <= 113 KB each");
  static_assert(NH == 1 || OFF_BAR + (2 * MAX_SLOTS + 5) * 8 + 16 <= 232448, "1 CTA per SM: 227 KB");
  static constexpr int BYTES = OFF_BAR + (2 * MAX_SLOTS + 5) * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES;             // the dynamic shared-memory base is 1024-byte aligned (declared so)
};

enum EgMode : int { EG_MSG0 = 0, EG_MSG = 1, EG_GATE = 2, EG_EU1 = 3, EG_EU2 = 4, EG_LIN = 5, EG_MSGA = 6 };
// EG_MSGA: EG_MSG whose epilogue also reduces the scalar messages over the in-edges of every destination node (edges are
// dst-major, so a thread walks its feature's 128 edge values in order): writes M / partL / partF exactly like k_vec_c.
// EG_LIN: plain linear out = W in + b over S inputs / S outputs (per-node halves of the next edge phases, node rows).
enum EgFlags : int { EGF_NODE_ROWS = 1,      // rows are nodes (EP = node count) instead of padded edge slots
                     EGF_IDENTITY = 2 };     // EG_GATE without the sigmoid (vector gate of the last position GVP)
// EG_EU1 / EG_EU2: the two linears of EdgeUpdate (flowmol/models/vector_field.py:844-880): h = SiLU(We [ef | rbf(d)] + EA[src] +
// EB[dst]);  ef <- LayerNorm(ef + SiLU(W2 h + b2)).  NH = 1 only.

struct EgArgs {
  const float* units;        // weight images (weights.py:tc_units)
  const float* bias;         // MSG: to_feats_out bias; GATE: gate bias (padded to 32)
  const float* in_s;         // MSG / GATE: [EPA][S] activations of the previous stage; MSG0: edge features ef [EP][F]
  const float* in_sh;        // MSG0 / MSG: [EPA][40] vector norms of this GVP
  const float* P;            // MSG0: per-node pre-activations [N][S]
  const float* x;            // MSG0: positions [N][3]
  float* out;                // MSG0 / MSG: [EPA][S];  GATE: [EPA][32];  EU1: h [EPA][F];  EU2: ef [EP][F] (in place)
  const float* ln_w;         // EU2: LayerNorm gamma / beta
  const float* ln_b;
  long long EP;              // padded edge slots (multiple of 64)
  long long* trace;          // optional clock64 stamps of one CTA (timeline experiments)
  int trace_cta;
  int flags;                 // EgFlags
  int dbg;                   // timing experiments: 1 no weight copies, 2 no MMA issue, 4 loaders skip global reads, 8 no epilogue math/stores
  float* M;                  // MSGA: aggregated messages [N][MW] and the per-64-slot-tile partial sums (see k_conv_edge)
  float* partF;
  float* partL;
  int* status;               // PREC 1: bit 0 is set when an activation leaves the fp16 operand range (|x| >= ACT_LIMIT_H16)
  const float* in_img;       // k_egemm_p<.., EGI_IN>: fp16 (hi, lo) operand images of the leading k-slabs (egemm_p.cuh)
  float* out_img;            // k_egemm_p<.., EGI_OUT>: where the output's operand images go
  const float* g_units;      // k_egemm_g: weight images of the GVP's gate linear (the EG_GATE units), its bias and the gate rows [EPA][32]
  const float* g_bias;
  float* g_out;
  const float* sh_img;       // k_egemm_g<.., SH_IMG>: the vector norms as the last k-slab's operand images (written by k_vecr_b), one 32 KB block per tile
};

// EG_MSGA: one finished segment sum of feature f.  Deliberately not inlined: the call sits behind a rarely taken branch in a
// 32-way unrolled loop, and inlining the three-way store there tripled the kernel's code size (instruction-cache misses).
__device__ __noinline__ void eg_store_segment(float* __restrict__ M, float* __restrict__ partF, float* __restrict__ partL, int mw,
                                              int d, long long t64, unsigned head_tail, int f, float run) {
  if (head_tail == 3u) M[(size_t)d * mw + f] = run;
  else if (head_tail & 1u) partL[(size_t)t64 * mw + f] = run;
  else partF[(size_t)t64 * mw + f] = run;
}

// PREC 0: error-compensated 3xTF32 (32 k values per 128-byte operand row).  PREC 1: scaled fp16 hi/lo ("fp16x3": 64 k values per
// row, kind::f16 MMAs at twice the TF32 rate, half the weight-image bytes; same 22 significand bits, see tc.cuh).  The byte
// geometry of the operand tiles, the weight ring, the stages and the barrier protocol are identical in both modes; the loaders
// still fetch 32-float chunks of K (two per fp16 slab).
template <class D, int MODE, int NH, int PREC>
__global__ void __launch_bounds__(EgPlan<NH>::THREADS, NH == 1 ? 2 : 1)
k_egemm_tc(const ModelRT m, const BatchRT bt, const EgArgs a) {
  pdl_launch();
  pdl_wait();
  using PL = EgPlan<NH>;
  constexpr int EG_T = PL::T, EG_XSTAGE = PL::XSTAGE;
  constexpr int LO_OFF = NH * 16384;                          // offset of the lo images inside a stage
  constexpr int S = D::S;
  constexpr bool IS_EU = MODE == EG_EU1 || MODE == EG_EU2;
  constexpr bool IS_MSG = MODE == EG_MSG || MODE == EG_MSGA;
  constexpr int K = MODE == EG_MSG0 ? D::KE0 : (IS_MSG ? D::K1 : (MODE == EG_EU1 ? D::F + D::R : (MODE == EG_EU2 ? D::F : S)));
  constexpr int KS = PREC ? 64 : 32;                          // k values per operand row (one slab)
  constexpr int NSLAB = (K + KS - 1) / KS;
  constexpr int NCH = (K + 31) / 32;                          // 32-float chunks of K the loaders fetch (== NSLAB for PREC 0)
  constexpr int LAST_KSTEPS = ((K - 1) % KS) / (KS / 4) + 1;
  constexpr int NMT = MODE == EG_GATE ? 1 : (IS_EU ? D::F / 128 : S / 128);
  constexpr int OW = MODE == EG_GATE ? 32 : (IS_EU ? D::F : S);          // output row width
  static_assert(!IS_EU || (NH == 1 && D::F == 128), "edge-update modes: one 128-edge half, F = 128");
  constexpr int UNIT_BYTES = MODE == EG_GATE ? 32 * 128 : TC_UNIT;
  // slots of UNIT_BYTES; the MMA always reads a full 128-row (16 KB) operand tile from the slot base (rows beyond a 32-row gate
  // unit are stale data feeding accumulator rows nobody reads), so the last slot must still have 16 KB behind it
  constexpr int EG_RING_FIT = (PL::RING_BYTES - TC_UNIT) / UNIT_BYTES + 1;
  constexpr int EG_RING = EG_RING_FIT < PL::MAX_SLOTS ? EG_RING_FIT : PL::MAX_SLOTS;
  constexpr int SH_W = 40;                                     // row pitch of the norm buffer
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* base = smem_dyn;
  uint8_t* xst = base + PL::OFF_X;
  uint8_t* ring = base + PL::OFF_RING;
  int* r_src = reinterpret_cast<int*>(base + PL::OFF_ROW);
  short* r_dd = reinterpret_cast<short*>(r_src + EG_T);        // dst - src (same molecule, |.| < 2000)
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + PL::OFF_BAR);
  uint64_t *w_full = bars, *w_empty = bars + PL::MAX_SLOTS, *x_full = bars + 2 * PL::MAX_SLOTS, *x_empty = x_full + 2, *acc_full = x_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long slot0 = (long long)blockIdx.x * EG_T;
  if (tid == 0) {
    for (int i = 0; i < EG_RING; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    tc::mbar_init(&x_full[0], 4 * NH); tc::mbar_init(&x_full[1], 4 * NH);
    tc::mbar_init(&x_empty[0], 1); tc::mbar_init(&x_empty[1], 1);
    tc::mbar_init(acc_full, 1);
    tc::fence_mbar_init();
  }
  constexpr uint32_t TMEM_COLS = NH == 2 ? 512 : 256;
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  // per-row bookkeeping: validity, and for the first message linear the (src, dst) pair and its distance
  float my_dist = 0.f;                        // MSG0: distance of this loader thread's own edge row
  if (tid >= 64) {
    const int r = tid - 64;
    const long long slot = slot0 + r;
    int s = -1, dd = 0;
    float dist = 0.f;
    if (a.flags & EGF_NODE_ROWS) {
      if (slot < a.EP) s = 0;
    } else if (slot < a.EP) {
      const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
      const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
      if (le < n * (n - 1)) {
        s = 0;
        if (MODE == EG_MSGA) {
          const int j = le / (n - 1), rem = le - j * (n - 1);
          s = bt.mol_node[mol] + j;                                   // destination node
          const bool tail = rem == n - 2;                            // the node's last in-edge
          // bit 0: the node's segment ends here inside this 64-slot tile; bit 1: ... with the node's last in-edge;
          // bit 2: that segment began with the node's first in-edge (it lies in the same 64-slot tile)
          dd = (((r & 63) == 63 || tail) ? 1 : 0) | (tail ? 2 : 0) | (rem <= (r & 63) ? 4 : 0);
        }
        if (MODE == EG_MSG0 || MODE == EG_EU1) {
          int i, j;
          edge_src_dst(le, n, i, j);
          const int nb = bt.mol_node[mol];
          s = nb + i;
          dd = j - i;
          float dx, dy, dz;
          dist = pair_dist(a.x, s, nb + j, dx, dy, dz);
        }
      }
    }
    r_src[r] = s;
    if (MODE == EG_MSGA) {      // one 32-bit mask per 32-row chunk (= per loader warp) instead of per-row flags
      const unsigned em = __ballot_sync(0xffffffffu, dd & 1), tm = __ballot_sync(0xffffffffu, dd & 2), hm = __ballot_sync(0xffffffffu, dd & 4);
      if (lane == 0) {
        unsigned* masks = reinterpret_cast<unsigned*>(r_dd);
        masks[(warp - 2) * 3 + 0] = em; masks[(warp - 2) * 3 + 1] = tm; masks[(warp - 2) * 3 + 2] = hm;
      }
    } else {
      r_dd[r] = (short)dd;
    }
    my_dist = dist;
  }
  if (a.trace && blockIdx.x == a.trace_cta && tid == 64) a.trace[0] = clock64();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- weight producer -------------------------------------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t NU = NSLAB * NMT * 2;
      for (uint32_t u = 0; u < NU; ++u) {
        const uint32_t sl = u % EG_RING, use = u / EG_RING;
        if (use > 0) tc::mbar_wait(&w_empty[sl], (use - 1) & 1);
        if (a.dbg & 1) { tc::mbar_arrive_expect_tx(&w_full[sl], 0u); continue; }
        tc::mbar_arrive_expect_tx(&w_full[sl], UNIT_BYTES);
        tc::bulk_g2s(ring + sl * UNIT_BYTES, reinterpret_cast<const uint8_t*>(a.units) + (size_t)u * UNIT_BYTES, UNIT_BYTES, &w_full[sl]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = PREC ? tc::idesc_f16(128, 128) : tc::idesc_tf32(128, 128);
      auto umma = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if (PREC) tc::umma_f16(d, da, db, id, acc); else tc::umma_tf32(d, da, db, id, acc);
      };
      uint32_t u = 0;
      for (int j = 0; j < NSLAB; ++j) {
        const int st = j & 1, ksteps = (j == NSLAB - 1) ? LAST_KSTEPS : 4;
        tc::mbar_wait(&x_full[st], (j >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t xb = tc::smem_u32(xst + st * EG_XSTAGE);
        for (int mt = 0; mt < NMT; ++mt) {
          {
            const uint32_t sl = u % EG_RING, use = u / EG_RING;
            tc::mbar_wait(&w_full[sl], use & 1);
            tc::tc_fence_after();
            const uint32_t wb = tc::smem_u32(ring + sl * UNIT_BYTES);
            for (int h = 0; h < NH; ++h) {
              const uint32_t d = tmem + (uint32_t)((mt * NH + h) * 128);
              for (int ks = 0; ks < ksteps && !(a.dbg & 2); ++ks) {
                const uint64_t dw = tc::desc_sw128(wb + 32 * ks);
                umma(d, dw, tc::desc_sw128(xb + LO_OFF + h * 16384 + 32 * ks), idesc, (j > 0 || ks > 0) ? 1u : 0u);
                umma(d, dw, tc::desc_sw128(xb + h * 16384 + 32 * ks), idesc, 1u);
              }
            }
            tc::umma_commit(&w_empty[sl]);
            ++u;
          }
          {
            const uint32_t sl = u % EG_RING, use = u / EG_RING;
            tc::mbar_wait(&w_full[sl], use & 1);
            tc::tc_fence_after();
            const uint32_t wb = tc::smem_u32(ring + sl * UNIT_BYTES);
            for (int h = 0; h < NH; ++h) {
              const uint32_t d = tmem + (uint32_t)((mt * NH + h) * 128);
              for (int ks = 0; ks < ksteps && !(a.dbg & 2); ++ks)
                umma(d, tc::desc_sw128(wb + 32 * ks), tc::desc_sw128(xb + h * 16384 + 32 * ks), idesc, 1u);
            }
            tc::umma_commit(&w_empty[sl]);
            ++u;
          }
        }
        tc::umma_commit(&x_empty[st]);
        if (a.trace && blockIdx.x == a.trace_cta) a.trace[16 + j] = clock64();
      }
      tc::umma_commit(acc_full);
    }
  } else {
    // ---- activation loaders: one edge row per thread, fp32 -> (hi, lo) SW128 images ------------------------------------------------
    const float inv_sigma = (float)D::R / m.rbf_dmax;
    // One warp owns 32 consecutive edge rows.  A load instruction covers 4 rows x 128 B (lanes 8g..8g+7 read the 8 16-byte
    // chunks of row 4i + g): full 128-byte lines per request.  The fetch of slab j+1 is issued before slab j is converted, so
    // the HBM / L2 round trip overlaps the wait for the stage and the MMAs of the previous slab.
    const int wrow0 = (warp - 2) * 32, lg = lane >> 3, ch = lane & 7;
    const float4 mu4 = *reinterpret_cast<const float4*>(m.g(G_RBF_MU) + ch * 4);      // this lane's four rbf centres
    auto fetch = [&](int j, float4 (&buf)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rw = wrow0 + 4 * i + lg;                     // row inside the tile
        const long long sl_ = slot0 + rw;
        const bool ok = r_src[rw] >= 0 && !(a.dbg & 4);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == EG_MSG0 || MODE == EG_EU1) {
          const float dd = __shfl_sync(0xffffffffu, my_dist, 4 * i + lg);      // the row's distance lives in lane (row % 32)
          // k order of the first message linear: rbf | ef | norms in the 3xTF32 images, ef | rbf | norms in the fp16 images
          // (weights.py: the edge features are whole k-slabs there, k_egemm_p fetches them as operand images)
          const bool rbf_slab = (MODE == EG_MSG0 && !PREC) ? j == 0 : j == D::F / 32;
          const int jf = PREC ? j : j - 1;                       // MSG0: ef chunk index
          if (rbf_slab) {
            if (ok) val = make_float4(rbf_fast(dd, mu4.x, inv_sigma), rbf_fast(dd, mu4.y, inv_sigma), rbf_fast(dd, mu4.z, inv_sigma),
                                      rbf_fast(dd, mu4.w, inv_sigma));
          } else if (MODE == EG_MSG0) {
            if (ok) {
              if (j <= D::F / 32) val = __ldg(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * D::F + jf * 32) + ch);
              else {
                const int k0 = (j - 1 - D::F / 32) * 32 + ch * 4;
                if (k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
              }
            }
          } else {
            if (ok) val = *(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * D::F + j * 32) + ch);
          }
        } else if (IS_MSG) {
          if (ok) {
            if (j < S / 32) val = *(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * S + j * 32) + ch);
            else {
              const int k0 = (j - S / 32) * 32 + ch * 4;
              if (k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
            }
          }
        } else if (MODE == EG_EU2) {
          if (ok) val = *(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * D::F + j * 32) + ch);
        } else {
          if (ok) val = *(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * S + j * 32) + ch);
        }
        buf[i] = val;
      }
    };
    // Prefetch distance: the narrow modes (one 128-feature tile: few MMAs per slab) are bound by the bytes in flight per SM,
    // not by the tensor pipe -- keep two slabs of loads outstanding there; the wide message linears hide one slab behind MMAs.
    constexpr bool DEEP = NMT == 1;
    float4 cur[8], nxt[8], nx2[8];
    float amax = 0.f;                                      // PREC 1: largest |activation| this thread converted
    fetch(0, cur);
    if (NCH > 1) fetch(1, nxt);
    if (a.trace && blockIdx.x == a.trace_cta && tid == 64) a.trace[1] = clock64();
    for (int j = 0; j < NCH; ++j) {
      // chunk j fills the whole operand row of slab j (PREC 0) or half `hf` of the row of slab j / 2 (PREC 1)
      const int sb = PREC ? (j >> 1) : j, hf = PREC ? (j & 1) : 0, st = sb & 1;
      if constexpr (DEEP) { if (j + 2 < NCH) fetch(j + 2, nx2); }
      if (sb >= 2 && hf == 0) tc::mbar_wait(&x_empty[st], ((sb >> 1) - 1) & 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rw = wrow0 + 4 * i + lg, hh = rw >> 7, rr_ = rw & 127;
        uint8_t* hi = xst + st * EG_XSTAGE + hh * 16384;
        const float4 val = cur[i];
        if constexpr (PREC) {
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(val.x), fabsf(val.y)), fmaxf(fabsf(val.z), fabsf(val.w))));
          uint2 vh, vl;
          tc::split_h16x2(val.x * tc::ACT_SCALE_H16, val.y * tc::ACT_SCALE_H16, vh.x, vl.x);
          tc::split_h16x2(val.z * tc::ACT_SCALE_H16, val.w * tc::ACT_SCALE_H16, vh.y, vl.y);
          const uint32_t off = tc::sw128_off_h(rr_, hf * 32 + ch * 4);
          *reinterpret_cast<uint2*>(hi + off) = vh;
          *reinterpret_cast<uint2*>(hi + LO_OFF + off) = vl;
        } else {
          float4 vh, vl;
          tc::split_tf32(val.x, vh.x, vl.x); tc::split_tf32(val.y, vh.y, vl.y);
          tc::split_tf32(val.z, vh.z, vl.z); tc::split_tf32(val.w, vh.w, vl.w);
          const uint32_t off = tc::sw128_off(rr_, ch * 4);
          *reinterpret_cast<float4*>(hi + off) = vh;
          *reinterpret_cast<float4*>(hi + LO_OFF + off) = vl;
        }
      }
      if (!PREC || hf == 1 || j == NCH - 1) {              // the slab's operand rows are complete
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&x_full[st])) : "memory");
        }
      }
      if (a.trace && blockIdx.x == a.trace_cta && tid == 64) a.trace[2 + (j < 14 ? j : 13)] = clock64();
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      if constexpr (DEEP) {
#pragma unroll
        for (int i = 0; i < 8; ++i) nxt[i] = nx2[i];
      } else {
        if (j + 2 < NCH) fetch(j + 2, nxt);
      }
    }
    if constexpr (PREC) {
      if (!(amax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);     // also catches NaN
    }
    // ---- epilogue: TMEM -> registers -> bias / gathered pre-activation -> activation -> coalesced global stores ------------------------
    const int q = warp & 3, eh = (warp - 2) >> 2;
    if ((MODE != EG_GATE || q == 0) && !(a.dbg & 8)) {
#pragma unroll 1
      for (int mt = 0; mt < NMT; ++mt) {
        const int f = mt * 128 + q * 32 + lane;
        float run = 0.f;                                       // MSGA: running segment sum of feature f
        const float bias = (MODE == EG_MSG0 || MODE == EG_EU1) ? 0.f : a.bias[f];
        // PREC 1: undo the power-of-two operand scales (exact); the factor sits right behind the weight units (weights.py:tc_units_h16)
        const float unscale = PREC ? a.units[(size_t)NSLAB * NMT * 2 * (UNIT_BYTES / 4)] : 1.0f;
        float* red = reinterpret_cast<float*>(xst);            // EU2: cross-warp LayerNorm partials (the stages are idle now)
        float pre[32], pnext[32];
        auto gather = [&](int c, float (&dst_)[32]) {          // per-edge pre-activations / residuals of chunk c (L2 / HBM gathers)
          if (MODE == EG_MSG0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst_[i] = __ldg(a.P + (size_t)max(r_src[eh * 128 + c * 32 + i], 0) * S + f);
          } else if (MODE == EG_EU1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int e = c * 32 + i, sn = max(r_src[e], 0);
              dst_[i] = __fadd_rn(__ldg(a.P + (size_t)sn * 2 * D::F + f), __ldg(a.P + (size_t)(sn + r_dd[e]) * 2 * D::F + D::F + f));
            }
          } else if (MODE == EG_EU2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst_[i] = a.in_sh[(size_t)(slot0 + c * 32 + i) * D::F + f];     // residual: current ef
          }
        };
        constexpr bool GATHERS = MODE == EG_MSG0 || MODE == EG_EU1 || MODE == EG_EU2;
        if (GATHERS) gather(0, pre);                          // in flight while the last MMAs drain
        if (mt == 0) {
          tc::mbar_wait(acc_full, 0);
          tc::tc_fence_after();
          if (a.trace && blockIdx.x == a.trace_cta && tid == 64) a.trace[30] = clock64();
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float acc[32];
          tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((mt * NH + eh) * 128 + c * 32), acc);
          if (GATHERS && c + 1 < 4) gather(c + 1, pnext);     // in flight while this chunk is finished
          tc::tmem_ld_wait();
          float* op = a.out + (size_t)(slot0 + eh * 128 + c * 32) * OW + f;
          unsigned em = 0, tm = 0, hm = 0;
          if (MODE == EG_MSGA) {
            const unsigned* masks = reinterpret_cast<const unsigned*>(r_dd) + (eh * 4 + c) * 3;
            em = masks[0]; tm = masks[1]; hm = masks[2];
          }
          if (MODE != EG_EU2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {      // padding rows are computed and stored too (their slots exist)
              const float z = (PREC ? acc[i] * unscale : acc[i]) + ((MODE == EG_MSG0 || MODE == EG_EU1) ? pre[i] : bias);
              if (MODE == EG_LIN) { op[(size_t)i * OW] = z; continue; }
              const float sg = sigmoid_fast(z);
              const float o = MODE == EG_GATE ? ((a.flags & EGF_IDENTITY) ? z : sg) : z * sg;
              op[(size_t)i * OW] = o;
              if (MODE == EG_MSGA) acc[i] = o;
            }
            if (MODE == EG_MSGA) {
              // dst-major edges: the sum over a node's in-edges is a running sum along this thread's row of the accumulator
              // (a second pass, so the activations above keep their instruction-level parallelism).  Slots after a molecule's
              // last edge add garbage that is dropped at the next 64-slot boundary.
              if ((c & 1) == 0) run = 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                run = __fadd_rn(run, acc[i]);
                if ((em >> i) & 1u) {                        // uniform over the CTA's epilogue threads
                  const int row = eh * 128 + c * 32 + i;
                  eg_store_segment(a.M, a.partF, a.partL, D::MW, r_src[row], (slot0 + row) >> 6,
                                   ((hm >> i) & 1u) | (((tm >> i) & 1u) << 1), f, run);
                  run = 0.f;
                }
              }
            }
          } else {
            // y = ef + SiLU(W2 h + b2);  LayerNorm over the 128 features of every edge = over the lanes of the 4 epilogue warps
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float z = (PREC ? acc[i] * unscale : acc[i]) + bias;
              acc[i] = __fadd_rn(pre[i], z * sigmoid_fast(z));
            }
            {
              float t[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) t[i] = acc[i];
              red[q * 32 + lane] = warp_transpose_sum(t);      // lane i <- sum over the warp's 32 features of edge i
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) pre[i] = (red[i] + red[32 + i] + red[64 + i] + red[96 + i]) * (1.0f / 128.0f);   // mean
            {
              float t[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) { const float dlt = acc[i] - pre[i]; t[i] = dlt * dlt; }
              red[128 + q * 32 + lane] = warp_transpose_sum(t);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float gam = a.ln_w[f], bet = a.ln_b[f];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float var = (red[128 + i] + red[160 + i] + red[192 + i] + red[224 + i]) * (1.0f / 128.0f);
              op[(size_t)i * OW] = (acc[i] - pre[i]) * rsqrtf(var + 1e-5f) * gam + bet;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");     // red[] is rewritten by the next chunk
          }
          if (GATHERS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) pre[i] = pnext[i];
          }
        }
      }
    }
  }
  if (a.trace && blockIdx.x == a.trace_cta && tid == 64) a.trace[31] = clock64();
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace fm
