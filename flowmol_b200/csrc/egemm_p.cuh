// k_egemm_p -- persistent, role-specialised version of k_egemm_tc (same modes, same arguments, same results bit for bit).
//
// What the profiles said.  k_egemm_tc (profiles/r01e): tensor pipe 31 % active, DRAM 36 % of peak; a CTA there is one 128-edge
// tile whose phases run one after the other with the same four warps loading and draining.  First persistent version (r01g-r01k):
// once consecutive linears hand their activations over as ready-made operand images (below), every mode of the kernel is bound by
// its EPILOGUE -- four warps, one per scheduler, ~20 instructions per output element -- while eight loader warps mostly spin.
// Hence this split: ONE CTA per SM walks over tiles (tile = blockIdx.x + i * gridDim.x) with
//
//   warp 0      weight producer   bulk-TMA ring of operand images, runs ahead across tile boundaries
//   warps 1-2   MMA issuers       tcgen05.mma into one of TWO TMEM accumulator buffers (2 x 256 columns); issuer w owns the 128-feature
//                                 tile w (its own accumulator columns and its own weight units)
//   warps 3-6   activation loaders k-slabs that exist as operand images: one 32 KB bulk-TMA copy per slab (issued by warp 3);
//                                 everything else (rbf of the edge length, vector norms, fp32 node rows): fp32 -> fp16 (hi, lo)
//                                 SW128 images, 32 rows per warp, 32-wide chunks, loads prefetched one chunk ahead across tiles
//   warps 7-14  epilogue          TMEM -> registers -> activation -> global; two warps per TMEM lane quarter (each takes 64 of
//                                 the tile's 128 rows), draining buffer b while the MMAs fill buffer b ^ 1
// The per-row bookkeeping of a tile (source / destination node of every edge row, segment flags) is computed by the loaders,
// which are one to three tiles ahead of the epilogue, and handed over through a 4-deep ring with its own mbarriers -- computed in
// the epilogue it was a chain of dependent global loads at the top of every tile with all eight epilogue warps parked behind it.
//
// All ring positions / mbarrier phases are running counters that every role advances identically per tile, so nothing is
// re-initialised between tiles.  PREC 1 (fp16x3 operands) only.
#pragma once
#include "egemm_tc.cuh"

namespace fm {

struct EgpPlan {
  static constexpr int T = 128;                             // rows (edges / nodes) per tile
  static constexpr int NST = 4;                             // activation stages
  static constexpr int XSTAGE = 32768;                      // hi image 16 KB | lo image 16 KB
  static constexpr int RING_BYTES = 4 * TC_UNIT;            // weight ring
  static constexpr int MAX_SLOTS = 12;
  static constexpr int NLW = 4, NEW = 8;                    // loader / epilogue warps
  static constexpr int THREADS = (3 + NLW + NEW) * 32;
  static constexpr int W_LOAD0 = 3, W_EPI0 = 3 + NLW;       // first loader / epilogue warp
  static_assert(W_EPI0 + NEW == THREADS / 32 && NLW * 32 == T, "one loader warp per 32 rows");
  static constexpr int OFF_X = 0;
  static constexpr int OFF_RING = NST * XSTAGE;
  static constexpr int NROWBUF = 4;                          // ring of per-tile row bookkeeping (loaders run up to 3 tiles ahead of the epilogue)
  static constexpr int OFF_ROW = OFF_RING + RING_BYTES;     // NROWBUF x { int src[T]; short dd[T] }: written by the loaders, read by the epilogue
  static constexpr int OFF_RED = OFF_ROW + NROWBUF * T * 6; // EU2 LayerNorm partials (2 row halves x 256 floats)
  static constexpr int OFF_BAR = OFF_RED + 2048;
  static constexpr int NBAR = 2 * MAX_SLOTS + 2 * NST + 4 + NROWBUF;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES;
  static_assert(BYTES <= 232448, "227 KB of shared memory per CTA");
};

// IMG: "operand images" between consecutive tensor-core linears.  Bit 1 (EGI_OUT): the epilogue stores its output rows (also)
// already split into the (hi, lo) fp16 SW128 operand images the next linear's MMAs read -- per 128-row tile and 64-wide k-slab
// one 32 KB block [hi image 16 KB | lo image 16 KB], the same 4 bytes per element -- at `out_img`.  Bit 0 (EGI_IN): `in_img`
// holds such images for the leading k-slabs of this linear and each of them is ONE 32 KB bulk-TMA copy into the activation stage
// (up to NST stages = 128 KB in flight per SM, no registers, no conversion instructions) instead of a loader round trip
// LDG -> split -> STS (profiles/r01g: ~40 % of the kernel's instructions, the GATE linear stuck at 2.4 TB/s).  The split is the
// same function of the same fp32 value on either side of HBM, so the results are bit-identical to the fp32 hand-over.
enum EgImg : int { EGI_IN = 1, EGI_OUT = 2 };

// (Measured and dropped, profiles/r02w: a cp.async.bulk.prefetch.L2 of the operand images two tiles ahead of every CTA made every
// mode SLOWER -- GATE 214 -> 264 us, MSG 687 -> 741 us: the stream is already deep enough for the DRAM controllers and the
// prefetched lines compete with the output stream for L2.)

// CL: thread-block cluster size (launch attribute).  What ncu says about the message linears (profiles/r01s, r02): 6.6 KB of L2 -> SM
// traffic per cycle, the measured fabric cap of the chip, and more than half of it is every CTA re-streaming the same 320 KB of
// weight images for every 128-edge tile.  With CL > 1 the CTAs of a cluster walk the weight stream together: unit u is fetched
// from L2 ONCE by CTA u % CL as a multicast bulk copy that lands in the ring slot of all CL CTAs (each CTA's w_full barrier
// expects the bytes itself), a slot is refilled when the MMAs of ALL CL CTAs have released it (tcgen05.commit multicast onto the
// peers' w_empty barriers, arrival count CL).  The MMAs themselves stay cta_group::1; tiles, activations, accumulators and
// epilogues are private to each CTA.  A CTA with one tile less than its peers still takes part in the last round of the weight
// protocol (waits for the units, releases them, issues its share of the copies).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

template <class D, int MODE, int IMG = 0, int CL = 1>
__global__ void __launch_bounds__(EgpPlan::THREADS, 1)
k_egemm_p(const ModelRT m, const BatchRT bt, const EgArgs a, const int n_tiles) {
  pdl_launch();
  pdl_wait();
  using PL = EgpPlan;
  constexpr int S = D::S;
  constexpr bool IMG_IN = (IMG & EGI_IN) != 0, IMG_OUT = (IMG & EGI_OUT) != 0;
  static_assert(!IMG_IN || (MODE != EG_LIN), "image input: MSG0 / EU1 (edge features), MSG / MSGA / GATE / EU2 (previous linear)");
  static_assert(!IMG_OUT || MODE == EG_MSG0 || MODE == EG_MSG || MODE == EG_MSGA || MODE == EG_EU1 || MODE == EG_EU2,
                "image output: activations of a next linear (EU2: the new edge features, next to their fp32 rows)");
  constexpr bool IS_EU = MODE == EG_EU1 || MODE == EG_EU2;
  constexpr bool IS_MSG = MODE == EG_MSG || MODE == EG_MSGA;
  constexpr int K = MODE == EG_MSG0 ? D::KE0 : (IS_MSG ? D::K1 : (MODE == EG_EU1 ? D::F + D::R : (MODE == EG_EU2 ? D::F : S)));
  constexpr int NSLAB = (K + 63) / 64;
  constexpr int NCH = (K + 31) / 32;                        // 32-float chunks of K
  constexpr int LAST_KSTEPS = ((K - 1) % 64) / 16 + 1;
  constexpr int NMT = MODE == EG_GATE ? 1 : (IS_EU ? D::F / 128 : S / 128);
  constexpr int OW = MODE == EG_GATE ? 32 : (IS_EU ? D::F : S);
  static_assert(!IS_EU || D::F == 128, "edge-update modes: F = 128");
  static_assert(D::F % 64 == 0 && S % 64 == 0, "image slabs are 64 k values");
  constexpr int UNIT_BYTES = MODE == EG_GATE ? 32 * 128 : TC_UNIT;
  constexpr int RING_FIT = (PL::RING_BYTES - TC_UNIT) / UNIT_BYTES + 1;      // the MMA reads 16 KB from every slot base
  constexpr int RING = RING_FIT < PL::MAX_SLOTS ? RING_FIT : PL::MAX_SLOTS;
  constexpr uint32_t NU = NSLAB * NMT * 2;                                     // weight units per tile
  constexpr int NST = PL::NST;
  // leading k-slabs that arrive as operand images: the edge features (MSG0, EU1: k order ef | rbf | ...) or the whole previous linear
  constexpr int NIMG = IMG_IN ? ((MODE == EG_MSG0 || IS_EU) ? D::F / 64 : S / 64) : 0;
  static_assert(NIMG <= NSLAB, "image slabs are a prefix of K");
  constexpr int FIRST_CH = 2 * NIMG;                        // first chunk the loader warps convert
  static_assert(!(MODE == EG_MSG0 || MODE == EG_EU1 || MODE == EG_MSGA) || FIRST_CH < NCH,
                "modes whose epilogue needs the row bookkeeping must have a converted chunk: the loaders publish it from rowinfo()");
  constexpr int SH_W = 40;
  constexpr int LO_OFF = 16384;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* xst = smem_dyn + PL::OFF_X;
  uint8_t* ring = smem_dyn + PL::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + PL::OFF_BAR);
  uint64_t *w_full = bars, *w_empty = bars + PL::MAX_SLOTS, *x_full = bars + 2 * PL::MAX_SLOTS, *x_empty = x_full + NST;
  uint64_t *acc_full = x_empty + NST, *acc_empty = acc_full + 2, *rows_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rows_full + PL::NROWBUF);
  constexpr bool NEED_ROWS = MODE == EG_MSG0 || MODE == EG_EU1 || MODE == EG_MSGA;      // the epilogue needs per-row node indices
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // rounds of the weight protocol: the tile count of the cluster's first CTA (the largest of the cluster; n_my <= n_cl <= n_my + 1)
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const int cfirst = (int)blockIdx.x - (int)crank;
  const int n_cl = CL > 1 ? ((cfirst < n_tiles) ? (n_tiles - 1 - cfirst) / (int)gridDim.x + 1 : 0) : n_my;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], CL); }
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&x_full[i], PL::NLW); tc::mbar_init(&x_empty[i], NMT); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], NMT); tc::mbar_init(&acc_empty[i], PL::NEW); }
    for (int i = 0; i < PL::NROWBUF; ++i) tc::mbar_init(&rows_full[i], PL::NLW);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // the peers' barriers exist before anything is multicast at them
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- weight producer ---------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t u = 0;
      for (int it = 0; it < n_cl; ++it) {
        for (uint32_t k = 0; k < NU; ++k, ++u) {
          const uint32_t sl = u % RING, use = u / RING;
          if (use > 0) tc::mbar_wait(&w_empty[sl], (use - 1) & 1);            // CL > 1: released by every CTA of the cluster
          tc::mbar_arrive_expect_tx(&w_full[sl], UNIT_BYTES);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.units) + (size_t)k * UNIT_BYTES;
          if (CL == 1) tc::bulk_g2s(ring + sl * UNIT_BYTES, src, UNIT_BYTES, &w_full[sl]);
          else if (u % CL == crank) tc::bulk_g2s_multicast(ring + sl * UNIT_BYTES, src, UNIT_BYTES, &w_full[sl], CMASK);
        }
      }
    }
  } else if (warp <= 2) {
    // ---- MMA issuers: warp 1 + mt issues the MMAs of feature tile mt (the whole warp stays converged, one elected lane issues) -----
    const int mt = warp - 1;
    if (mt < NMT) {
      const bool leader = tc::elect_one();
      const uint32_t idesc = tc::idesc_f16(128, 128);
      // this issuer's weight units are every NMT-th pair of the stream: ring slot / phase advance by 2 * NMT units per slab
      uint32_t sl = 2 * mt, ph = 0, g = 0;
      const uint32_t ring_lo = tc::smem_u32(ring) >> 4, x_lo = tc::smem_u32(xst) >> 4;
      for (int it = 0; it < n_my; ++it) {
        const int b = it & 1;
        if (it >= 2) { tc::mbar_wait(&acc_empty[b], ((it >> 1) - 1) & 1); tc::tc_fence_after(); }
        const uint32_t d = tmem + (uint32_t)(b * 256 + mt * 128);
        for (int j = 0; j < NSLAB; ++j, ++g) {
          const uint32_t st = g % NST, ksteps = (j == NSLAB - 1) ? LAST_KSTEPS : 4;
          tc::mbar_wait(&x_full[st], (g / NST) & 1);
          const uint32_t xh = x_lo + st * (PL::XSTAGE >> 4), xl = xh + (LO_OFF >> 4);
          tc::mbar_wait(&w_full[sl], ph);
          tc::tc_fence_after();
          const uint32_t wh = ring_lo + sl * (UNIT_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                const uint64_t dw = tc::desc_sw128_lo(wh + 2 * ks);
                tc::umma_f16(d, dw, tc::desc_sw128_lo(xl + 2 * ks), idesc, (j > 0 || ks > 0) ? 1u : 0u);
                tc::umma_f16(d, dw, tc::desc_sw128_lo(xh + 2 * ks), idesc, 1u);
              }
            }
            if (CL == 1) tc::umma_commit(&w_empty[sl]); else tc::umma_commit_multicast(&w_empty[sl], CMASK);
          }
          tc::mbar_wait(&w_full[sl + 1], ph);
          tc::tc_fence_after();
          const uint32_t wl = wh + (UNIT_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
              if (ks < ksteps) tc::umma_f16(d, tc::desc_sw128_lo(wl + 2 * ks), tc::desc_sw128_lo(xh + 2 * ks), idesc, 1u);
            if (CL == 1) tc::umma_commit(&w_empty[sl + 1]); else tc::umma_commit_multicast(&w_empty[sl + 1], CMASK);
            tc::umma_commit(&x_empty[st]);
          }
          sl += 2 * NMT;
          if (sl >= RING) { sl -= RING; ph ^= 1; }
        }
        if (leader) tc::umma_commit(&acc_full[b]);
      }
      if (CL > 1 && n_cl > n_my) {
        // one tile less than the cluster's first CTA: take the last round's weight units and release them untouched
        for (int j = 0; j < NSLAB; ++j) {
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            tc::mbar_wait(&w_full[sl + hl], ph);
            tc::tc_fence_after();
            if (leader) tc::umma_commit_multicast(&w_empty[sl + hl], CMASK);
          }
          sl += 2 * NMT;
          if (sl >= RING) { sl -= RING; ph ^= 1; }
        }
      }
    }
  } else if (warp < PL::W_EPI0) {
    // ---- activation loaders ------------------------------------------------------------------------------------------------------------------
    // Warp w owns rows [32 w, 32 w + 32) of the tile.  Row bookkeeping (validity, edge length) lives in registers: lane l holds row
    // wrow0 + l, the fetch reads it with a shuffle.  A load instruction covers 4 rows x 128 B (lanes 8g..8g+7 read the 8 16-byte
    // pieces of row 4i + g of a 32-float chunk).
    const int wrow0 = (warp - PL::W_LOAD0) * 32, lg = lane >> 3, ch = lane & 7;
    const float inv_sigma = (float)D::R / m.rbf_dmax;
    const float4 mu4 = *reinterpret_cast<const float4*>(m.g(G_RBF_MU) + ch * 4);      // this lane's four rbf centres
    int r_ok = 0;              // row valid (of the tile the NEXT fetch belongs to)
    float r_dist = 0.f;
    long long f_slot0 = 0;
    auto rowinfo = [&](int it) {
      f_slot0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * PL::T;
      const int r = wrow0 + lane;                               // this lane's row of the tile
      const long long slot = f_slot0 + r;
      int ok = 0, s = -1, dd = 0;                               // s / dd: the epilogue's bookkeeping (NEED_ROWS modes)
      float dist = 0.f;
      if (a.flags & EGF_NODE_ROWS) {
        ok = slot < a.EP;
      } else if (slot < a.EP) {
        const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
        const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
        if (le < n * (n - 1)) {
          ok = 1;
          if (MODE == EG_MSG0 || MODE == EG_EU1) {
            int i, j;
            edge_src_dst(le, n, i, j);
            const int nb = bt.mol_node[mol];
            float dx, dy, dz;
            dist = pair_dist(a.x, nb + i, nb + j, dx, dy, dz);
            s = nb + i;                                           // source node; dd = dst - src (same molecule, |.| < 2000)
            dd = j - i;
          } else if (MODE == EG_MSGA) {
            const int j = le / (n - 1), rem = le - j * (n - 1);
            s = bt.mol_node[mol] + j;                             // destination node
            const bool tail = rem == n - 2;                       // the node's last in-edge
            // bit 0: the node's segment ends here inside this 64-slot tile; bit 1: ... with the node's last in-edge;
            // bit 2: that segment began with the node's first in-edge (it lies in the same 64-slot tile)
            dd = (((r & 63) == 63 || tail) ? 1 : 0) | (tail ? 2 : 0) | (rem <= (r & 63) ? 4 : 0);
          }
        }
      }
      r_ok = ok;
      r_dist = dist;
      if (NEED_ROWS) {                                            // publish the tile's bookkeeping to the epilogue warps
        int* r_src = reinterpret_cast<int*>(smem_dyn + PL::OFF_ROW + (it % PL::NROWBUF) * (PL::T * 6));
        short* r_dd = reinterpret_cast<short*>(r_src + PL::T);
        r_src[r] = s;
        if (MODE == EG_MSGA) {      // one 32-bit mask per 32-row chunk (= per loader warp) instead of per-row flags
          const unsigned em = __ballot_sync(0xffffffffu, dd & 1), tm = __ballot_sync(0xffffffffu, dd & 2), hm = __ballot_sync(0xffffffffu, dd & 4);
          if (lane == 0) {
            unsigned* masks = reinterpret_cast<unsigned*>(r_dd);
            masks[(r >> 5) * 3 + 0] = em; masks[(r >> 5) * 3 + 1] = tm; masks[(r >> 5) * 3 + 2] = hm;
          }
        } else {
          r_dd[r] = (short)dd;
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&rows_full[it % PL::NROWBUF])) : "memory");
      }
    };
    // chunk j = k values [32 j, 32 j + 32) of this linear's input row; j is a compile-time constant at every call site (the slab /
    // chunk loops are fully unrolled), so the source selection below folds to one path per call.
    // k order: MSG0  ef(F) | rbf(32) | norms;  EU1  ef(F) | rbf(32);  MSG / MSGA  s'(S) | norms;  EU2  h(F);  GATE / LIN  s(S)
    auto fetch = [&](const int j, float4 (&buf)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + lg;                              // row inside the warp's 32
        const long long sl_ = f_slot0 + wrow0 + rl;
        const bool ok = __shfl_sync(0xffffffffu, r_ok, rl) != 0;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == EG_MSG0 || MODE == EG_EU1) {
          const float dd = __shfl_sync(0xffffffffu, r_dist, rl);
          if (j == D::F / 32) {
            if (ok) val = make_float4(rbf_fast(dd, mu4.x, inv_sigma), rbf_fast(dd, mu4.y, inv_sigma), rbf_fast(dd, mu4.z, inv_sigma),
                                      rbf_fast(dd, mu4.w, inv_sigma));
          } else if (j < D::F / 32) {
            if (ok) val = __ldg(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * D::F + j * 32) + ch);
          } else {                                               // MSG0 only
            const int k0 = (j - 1 - D::F / 32) * 32 + ch * 4;
            if (ok && k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
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
    float4 cur[8], nxt[8];
    float amax = 0.f;
    if (n_my > 0 && FIRST_CH < NCH) { rowinfo(0); fetch(FIRST_CH, cur); }
    uint32_t g = 0;                                            // running slab counter (stage / phase bookkeeping)
    for (int it = 0; it < n_my; ++it) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s, ++g) {
        const uint32_t st = g % NST, use = g / NST;
        if (s < NIMG) {
          // operand-image slab: one 32 KB bulk copy (hi | lo) straight into the stage, issued by the first loader warp.  Every
          // loader warp still waits for the stage and arrives, so the barrier's arrival count is the same for both kinds of slab
          // and no warp can run a whole phase ahead of the others.
          if (use > 0) tc::mbar_wait(&x_empty[st], (use - 1) & 1);
          if (lane == 0) {
            if (warp == PL::W_LOAD0) {
              const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
              tc::mbar_arrive_expect_tx(&x_full[st], PL::XSTAGE);
              tc::bulk_g2s(xst + st * PL::XSTAGE, reinterpret_cast<const uint8_t*>(a.in_img) + ((size_t)tile * NIMG + s) * PL::XSTAGE,
                           PL::XSTAGE, &x_full[st]);
            } else {
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&x_full[st])) : "memory");
            }
          }
          __syncwarp();
          continue;
        }
        uint8_t* hi = xst + st * PL::XSTAGE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * s + h;
          if (j >= NCH) break;                                 // K spans an odd number of chunks: the MMAs stop before this half
          // prefetch the next chunk: same tile, or the first converted chunk of the next tile
          if (j + 1 < NCH) {
            fetch(j + 1, nxt);
          } else if (it + 1 < n_my) {
            rowinfo(it + 1);
            fetch(FIRST_CH, nxt);
          }
          if (h == 0 && use > 0) tc::mbar_wait(&x_empty[st], (use - 1) & 1);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr_ = wrow0 + 4 * i + lg;
            const float4 val = cur[i];
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(val.x), fabsf(val.y)), fmaxf(fabsf(val.z), fabsf(val.w))));
            uint2 vh, vl;
            tc::split_h16x2(val.x * tc::ACT_SCALE_H16, val.y * tc::ACT_SCALE_H16, vh.x, vl.x);
            tc::split_h16x2(val.z * tc::ACT_SCALE_H16, val.w * tc::ACT_SCALE_H16, vh.y, vl.y);
            const uint32_t off = tc::sw128_off_h(rr_, h * 32 + ch * 4);
            *reinterpret_cast<uint2*>(hi + off) = vh;
            *reinterpret_cast<uint2*>(hi + LO_OFF + off) = vl;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&x_full[st])) : "memory");
      }
    }
    if (!(amax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  } else {
    // ---- epilogue: TMEM -> registers -> bias / gathered pre-activation -> activation -> coalesced global stores ---------------------------
    // warp -> TMEM lane quarter q (a warp may only touch lanes 32 (warp % 4) ...) and row half hf: rows [64 hf, 64 hf + 64) = chunks 2 hf, 2 hf + 1
    const int q = warp & 3, hf = (warp - PL::W_EPI0) >> 2;
    float* red = reinterpret_cast<float*>(smem_dyn + PL::OFF_RED) + hf * 256;
    const float unscale = a.units[(size_t)NU * (UNIT_BYTES / 4)];
    float omax = 0.f;                                                  // IMG_OUT: largest |activation| this thread split into fp16 (hi, lo)
    constexpr bool GATHERS = MODE == EG_MSG0 || MODE == EG_EU1 || MODE == EG_EU2;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      const long long slot0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * PL::T;
      // row bookkeeping of this tile, published by the loaders (they are ahead: normally no wait at all).  Ring of NROWBUF: a
      // loader can only be three tiles ahead of the slowest epilogue warp (stages -> accumulator buffers), see rowinfo().
      int* r_src = reinterpret_cast<int*>(smem_dyn + PL::OFF_ROW + (it % PL::NROWBUF) * (PL::T * 6));
      short* r_dd = reinterpret_cast<short*>(r_src + PL::T);
      if (NEED_ROWS) tc::mbar_wait(&rows_full[it % PL::NROWBUF], (it / PL::NROWBUF) & 1);
      const bool active = MODE != EG_GATE || q == 0;
#pragma unroll 1
      for (int mt = 0; mt < NMT; ++mt) {
        const int f = mt * 128 + q * 32 + lane;
        float run = 0.f;
        const float bias = (MODE == EG_MSG0 || MODE == EG_EU1 || !active) ? 0.f : a.bias[f];
        float pre[32], pnext[32];
        auto gather = [&](int c, float (&dst_)[32]) {
          if (MODE == EG_MSG0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst_[i] = __ldg(a.P + (size_t)max(r_src[c * 32 + i], 0) * S + f);
          } else if (MODE == EG_EU1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int e = c * 32 + i, sn = max(r_src[e], 0);
              dst_[i] = __fadd_rn(__ldg(a.P + (size_t)sn * 2 * D::F + f), __ldg(a.P + (size_t)(sn + r_dd[e]) * 2 * D::F + D::F + f));
            }
          } else if (MODE == EG_EU2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst_[i] = a.in_sh[(size_t)(slot0 + c * 32 + i) * D::F + f];
          }
        };
        if (GATHERS) gather(2 * hf, pre);
        if (mt == 0) {
          tc::mbar_wait(&acc_full[b], (it >> 1) & 1);
          tc::tc_fence_after();
        }
        if (active) {
#pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            const int c = 2 * hf + ci;
            float acc[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + mt * 128 + c * 32), acc);
            if (GATHERS && ci == 0) gather(c + 1, pnext);
            tc::tmem_ld_wait();
            float* op = a.out + (size_t)(slot0 + c * 32) * OW + f;
            unsigned em = 0, tm = 0, hm = 0;
            if (MODE == EG_MSGA) {
              const unsigned* masks = reinterpret_cast<const unsigned*>(r_dd) + c * 3;
              em = masks[0]; tm = masks[1]; hm = masks[2];
            }
            if (MODE != EG_EU2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float z = acc[i] * unscale + ((MODE == EG_MSG0 || MODE == EG_EU1) ? pre[i] : bias);
                if (MODE == EG_LIN) { op[(size_t)i * OW] = z; continue; }
                const float sg = sigmoid_fast(z);
                const float o = MODE == EG_GATE ? ((a.flags & EGF_IDENTITY) ? z : sg) : z * sg;
                if (!IMG_OUT) op[(size_t)i * OW] = o;
                if (MODE == EG_MSGA || IMG_OUT) acc[i] = o;
              }
              if (MODE == EG_MSGA) {
                if (ci == 0) run = 0.f;
                const unsigned em_lo = em & 0x7fffffffu;
                if (__popc(em_lo) <= 1) {
                  // Common case (in-edge segments of n - 1 >= 32 rows): at most one segment ends before the chunk's last row.
                  // Branch-free running sum -- a conditional branch per row cost ~11 k cycles per tile with one epilogue warp per
                  // scheduler (ncu r01k: MSGA 864 us vs MSG 533 us); same additions in the same order.
                  float stash = 0.f;
#pragma unroll
                  for (int i = 0; i < 31; ++i) {
                    run = __fadd_rn(run, acc[i]);
                    const bool end = (em_lo >> i) & 1u;
                    stash = end ? run : stash;
                    run = end ? 0.f : run;
                  }
                  run = __fadd_rn(run, acc[31]);
                  if (em_lo) {
                    const int i0 = __ffs(em_lo) - 1, row = c * 32 + i0;
                    eg_store_segment(a.M, a.partF, a.partL, D::MW, r_src[row], (slot0 + row) >> 6,
                                     ((hm >> i0) & 1u) | (((tm >> i0) & 1u) << 1), f, stash);
                  }
                  if (em >> 31) {
                    const int row = c * 32 + 31;
                    eg_store_segment(a.M, a.partF, a.partL, D::MW, r_src[row], (slot0 + row) >> 6, (hm >> 31) | ((tm >> 31) << 1), f, run);
                    run = 0.f;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    run = __fadd_rn(run, acc[i]);
                    if ((em >> i) & 1u) {
                      const int row = c * 32 + i;
                      eg_store_segment(a.M, a.partF, a.partL, D::MW, r_src[row], (slot0 + row) >> 6,
                                       ((hm >> i) & 1u) | (((tm >> i) & 1u) << 1), f, run);
                      run = 0.f;
                    }
                  }
                }
              }
            } else {
              // y = ef + SiLU(W2 h + b2);  LayerNorm over the 128 features of every edge = over the lanes of the 4 warps of this row half
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float z = acc[i] * unscale + bias;
                acc[i] = __fadd_rn(pre[i], z * sigmoid_fast(z));
              }
              {
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = acc[i];
                red[q * 32 + lane] = warp_transpose_sum(t);
              }
              if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
#pragma unroll
              for (int i = 0; i < 32; ++i) pre[i] = (red[i] + red[32 + i] + red[64 + i] + red[96 + i]) * (1.0f / 128.0f);
              {
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { const float dlt = acc[i] - pre[i]; t[i] = dlt * dlt; }
                red[128 + q * 32 + lane] = warp_transpose_sum(t);
              }
              if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
              const float gam = a.ln_w[f], bet = a.ln_b[f];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float var = (red[128 + i] + red[160 + i] + red[192 + i] + red[224 + i]) * (1.0f / 128.0f);
                const float y = (acc[i] - pre[i]) * rsqrtf(var + 1e-5f) * gam + bet;
                op[(size_t)i * OW] = y;
                if (IMG_OUT) acc[i] = y;
              }
              if (hf == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
            }
            if constexpr (IMG_OUT) {
              // Operand images of the next linear.  Feature f is k = f % 64 of slab f / 64; row r of a slab image is 128 bytes of
              // 64 fp16 whose 16-byte pieces are XOR-swizzled by r % 8.  Lanes 2m / 2m+1 hold neighbouring k: they swap one
              // packed (hi, lo) word per row pair so that the even lane stores the 32-bit (k, k+1) words of row 2t and the odd
              // lane those of row 2t + 1 -- a warp store covers two rows x 64 contiguous bytes (four full sectors).
              const int odd = lane & 1, kk = (q & 1) * 32 + (lane & ~1);
              uint8_t* ob = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)(slot0 >> 7) * (OW / 64) + (size_t)(mt * 2 + (q >> 1))) * PL::XSTAGE +
                            (size_t)((c * 32 + odd) * 128 + (kk & 7) * 2);
              const uint32_t c0x = (uint32_t)((kk >> 3) ^ odd);
              const uint32_t sel_keep = odd ? 0x7632u : 0x5410u, sel_send = odd ? 0x5410u : 0x7632u;
              const uint32_t sel_hi = odd ? 0x1054u : 0x5410u, sel_lo = odd ? 0x3276u : 0x7632u;
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                const float v0 = acc[2 * t], v1 = acc[2 * t + 1];
                omax = fmaxf(omax, fmaxf(fabsf(v0), fabsf(v1)));
                uint32_t h2, l2;                              // (hi(v0), hi(v1)), (lo(v0), lo(v1))
                tc::split_h16x2(v0, v1, h2, l2);
                const uint32_t keep = __byte_perm(h2, l2, sel_keep);      // (hi, lo) of the row this lane stores
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, __byte_perm(h2, l2, sel_send), 1);
                uint8_t* dst_ = ob + (2 * t) * 128 + ((c0x ^ (uint32_t)((2 * t) & 7)) << 4);
                *reinterpret_cast<uint32_t*>(dst_) = __byte_perm(keep, recv, sel_hi);
                *reinterpret_cast<uint32_t*>(dst_ + LO_OFF) = __byte_perm(keep, recv, sel_lo);
              }
            }
            if (GATHERS) {
#pragma unroll
              for (int i = 0; i < 32; ++i) pre[i] = pnext[i];
            }
          }
        }
      }
      // this warp has read everything it needs from accumulator buffer b
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&acc_empty[b])) : "memory");
    }
    if (IMG_OUT && !(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);      // also catches NaN
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no peer copy / commit may still be aimed at this CTA's shared memory when it exits
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// fp32 edge features -> their operand images (the entry of the image chain: k_edge_init writes fp32 rows only).  Dead rows (the
// padding slots of a molecule's block and everything beyond EP) become zeros in BOTH buffers, so that the linears -- which compute
// every row of a tile -- never see uninitialised data there.
template <class D>
__global__ void __launch_bounds__(256) k_ef_image(const BatchRT bt, float* __restrict__ ef, float* __restrict__ img, long long EP, int n_tiles) {
  pdl_launch();
  pdl_wait();
  static_assert(D::F % 64 == 0, "image slabs are 64 k values");
  constexpr int F4 = D::F / 4;
  __shared__ int ok_row[128];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 128) {
      const long long slot = (long long)tile * 128 + threadIdx.x;
      int ok = 0;
      if (slot < EP) {
        const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
        const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
        ok = le < n * (n - 1);
      }
      ok_row[threadIdx.x] = ok;
    }
    __syncthreads();
    uint8_t* ib = reinterpret_cast<uint8_t*>(img) + (size_t)tile * (D::F / 64) * EgpPlan::XSTAGE;
    for (int idx = threadIdx.x; idx < 128 * F4; idx += 256) {
      const int row = idx / F4, k = (idx - row * F4) * 4;
      float4* src = reinterpret_cast<float4*>(ef + ((size_t)tile * 128 + row) * D::F + k);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok_row[row]) v = *src; else *src = v;
      uint2 vh, vl;
      tc::split_h16x2(v.x, v.y, vh.x, vl.x);
      tc::split_h16x2(v.z, v.w, vh.y, vl.y);
      uint8_t* d = ib + (size_t)(k >> 6) * EgpPlan::XSTAGE + tc::sw128_off_h(row, k & 63);
      *reinterpret_cast<uint2*>(d) = vh;
      *reinterpret_cast<uint2*>(d + 16384) = vl;
    }
  }
}

}  // namespace fm
