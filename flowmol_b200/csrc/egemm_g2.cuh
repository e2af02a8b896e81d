// k_egemm_g2 -- k_egemm_g (a message linear of GVP 1 / 2 with its gate linear) on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// What the knock-out timings said about k_egemm_g (profiles/r02i): with the epilogue switched off the MMAs + loads alone take 470 us,
// the MMAs alone 420 us -- far above the tensor floor of the launch (~250 us).  Every M = 128, N = 128 MMA reads 4 KB of A and 4 KB of
// B from shared memory in its 64 cycles = 128 B / clk, the whole shared-memory bandwidth of the SM, and the bulk copies that refill
// the same memory (480 KB per tile: 320 KB of weight images + 160 KB of activations) compete for it.  A CTA pair shares the weights:
//     D[256 edges][256 features] += X[256 edges][16 k] . W[256 features][16 k]^T        (cta_group::2, M = 256, N = 256)
// CTA r of the pair supplies the 128 edge rows of ITS tile as A and the weight unit of features [128 r, 128 r + 128) as its half of
// B; the accumulators land in each CTA's own tensor memory exactly as before.  Per CTA and k-step the tensor core now reads 8 KB in
// 128 cycles (64 B / clk), each CTA streams only HALF of the weight images (160 KB per tile, so the 64 KB ring is two k-slabs deep
// instead of one) and the L2 -> SM traffic of the weights halves.  Everything tcgen05 in a kernel must use one cta_group, so the gate
// MMAs (A operand in tensor memory) become M = 256, N = 32 with each CTA holding 16 of the 32 gate rows of Wg.
//
// Protocol on top of k_egemm_g's: only the pair's first CTA (the leader) issues MMAs.  Its operand barriers (x_full, w_full, wg_full)
// expect one more arrival: the other CTA's idle issuer warp relays the completion of ITS barrier of the same name (remote
// mbarrier.arrive through the cluster address).  tcgen05.commit.cta_group::2 with a multicast mask releases the barriers of the same
// name in both CTAs (x_empty, w_empty, acc_full, gate_full).  The epilogue warps of BOTH CTAs report to the leader's a_ready /
// acc_empty.  Same MMAs on the same operands in the same order per accumulator column as k_egemm_g: bit-identical results.
#pragma once
#include "egemm_e.cuh"

namespace fm {

namespace tc {
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs this thread issued -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrival on the barrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait with cluster-scope acquire (arrivals come from the other CTA of the pair)
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if (++spins > 400000u) __trap();
  }
}
}  // namespace tc

template <class D, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EggPlan::THREADS, 1)
k_egemm_g2(const ModelRT m, const BatchRT bt, const EgArgs a, const int n_tiles) {
  pdl_launch();
  pdl_wait();
  using PL = EggPlan;
  static_assert(MODE == EG_MSG || MODE == EG_MSGA, "gate-fused message linears of GVP 1 (image out) and GVP 2 (segment sum)");
  static_assert(D::S == 256, "256 output features = one N = 256 MMA, eight 32-feature chunks");
  constexpr int S = D::S;
  constexpr int K = D::K1;
  constexpr int NSLAB = (K + 63) / 64;
  constexpr int NCH = (K + 31) / 32;
  constexpr int LAST_KSTEPS = ((K - 1) % 64) / 16 + 1;
  constexpr int NST = PL::NST;
  constexpr int NIMG = S / 64;
  constexpr int FIRST_CH = 2 * NIMG;
  static_assert(FIRST_CH < NCH, "the loaders publish the row bookkeeping from their first converted chunk");
  constexpr int SH_W = 40;
  constexpr int LO_OFF = 16384;
  constexpr bool AGG = MODE == EG_MSGA;
  constexpr uint32_t GCOL = 224;                             // gate accumulator = the parked chunk's accumulator columns
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* xst = smem_dyn + PL::OFF_X;
  uint8_t* ring = smem_dyn + PL::OFF_RING;
  uint8_t* wg = smem_dyn + PL::OFF_WG;
  uint8_t* park = smem_dyn + PL::OFF_PARK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + PL::OFF_BAR);
  uint64_t *w_full = bars, *w_empty = bars + 4, *x_full = bars + 8, *x_empty = x_full + NST;
  uint64_t *acc_full = x_empty + NST, *acc_empty = acc_full + 2, *rows_full = acc_empty + 2;
  uint64_t *a_ready = rows_full + PL::NROWBUF, *gate_full = a_ready + 2, *wg_full = gate_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wg_full + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tile = blockIdx.x + it * gridDim.x: the pair's CTAs (consecutive blockIdx.x, even grid, even tile count) take neighbouring tiles
  // and always the same number of them
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t crank = cluster_ctarank();
  const bool leader_cta = crank == 0;
  const uint32_t extra = leader_cta ? 1u : 0u;              // the leader's operand barriers also wait for the other CTA's relay
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&w_full[i], 1 + extra); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&x_full[i], PL::NLW + extra); tc::mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], PL::NEW);                 // the four warps that read the gate accumulator, of both CTAs (leader's copy)
      tc::mbar_init(&a_ready[i], 2 * PL::NEW);               // the eight epilogue warps of both CTAs (leader's copy)
      tc::mbar_init(&gate_full[i], 1);
    }
    for (int i = 0; i < PL::NROWBUF; ++i) tc::mbar_init(&rows_full[i], PL::NLW);
    tc::mbar_init(wg_full, 1 + extra);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc2(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // both CTAs' barriers and tensor memory exist before anything crosses over
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- weight producer (+ the resident gate weights, once) --------------------------------------------------------------------------
    if (lane == 0) {
      if (n_my > 0) {
        // gate weights: this CTA's 16 of the 32 gate rows of every unit (two 8-row swizzle atoms = 2 KB), stored compactly
        tc::mbar_arrive_expect_tx(wg_full, PL::WG_BYTES / 2);
        for (int u = 0; u < 8; ++u)
          tc::bulk_g2s(wg + u * 2048, reinterpret_cast<const uint8_t*>(a.g_units) + (size_t)u * 4096 + crank * 2048, 2048, wg_full);
      }
      // this CTA's half of B: the hi and lo unit of features [128 crank, 128 crank + 128) of every k-slab (units 4 j + 2 crank, + 1 of the
      // stream); four ring slots = two k-slabs in flight
      uint32_t u = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < NSLAB; ++j) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.units) + (size_t)(4 * j + 2 * crank) * TC_UNIT;
#pragma unroll
          for (int hl = 0; hl < 2; ++hl, ++u) {
            const uint32_t sl = u & 3, use = u >> 2;
            if (use > 0) tc::mbar_wait(&w_empty[sl], (use - 1) & 1);
            if (a.dbg & 1) { tc::mbar_arrive_expect_tx(&w_full[sl], 0u); continue; }
            tc::mbar_arrive_expect_tx(&w_full[sl], TC_UNIT);
            tc::bulk_g2s(ring + sl * TC_UNIT, src + hl * TC_UNIT, TC_UNIT, &w_full[sl]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- main MMA issuer (leader CTA) / operand relay (the other CTA) ------------------------------------------------------------------------
    if (leader_cta) {
      const bool leader = tc::elect_one();
      const uint32_t idesc = tc::idesc_f16(256, 256);
      const uint32_t w_base = tc::smem_u32(ring) >> 4, x_lo = tc::smem_u32(xst) >> 4;
      uint32_t g = 0, u = 0;
      for (int it = 0; it < n_my; ++it) {
        const int b = it & 1;
        if (it >= 2) { tc::mbar_wait_cl(&acc_empty[b], ((it >> 1) - 1) & 1); tc::tc_fence_after(); }
        const uint32_t d = tmem + (uint32_t)(b * 256);
        for (int j = 0; j < NSLAB; ++j, ++g) {
          const uint32_t st = g % NST, ksteps = (j == NSLAB - 1) ? LAST_KSTEPS : 4;
          tc::mbar_wait_cl(&x_full[st], (g / NST) & 1);
          const uint32_t xh = x_lo + st * (PL::XSTAGE >> 4), xl = xh + (LO_OFF >> 4);
          {
            const uint32_t sl = u & 3;
            tc::mbar_wait_cl(&w_full[sl], (u >> 2) & 1);
            tc::tc_fence_after();
            const uint32_t wh = w_base + sl * (TC_UNIT >> 4);
            if (leader) {
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks) {
                if (ks < ksteps && !(a.dbg & 2)) {
                  const uint64_t dw = tc::desc_sw128_lo(wh + 2 * ks);
                  tc::umma2_f16(d, tc::desc_sw128_lo(xl + 2 * ks), dw, idesc, (j > 0 || ks > 0) ? 1u : 0u);
                  tc::umma2_f16(d, tc::desc_sw128_lo(xh + 2 * ks), dw, idesc, 1u);
                }
              }
              tc::umma2_commit(&w_empty[sl]);
            }
            ++u;
          }
          {
            const uint32_t sl = u & 3;
            tc::mbar_wait_cl(&w_full[sl], (u >> 2) & 1);
            tc::tc_fence_after();
            const uint32_t wl = w_base + sl * (TC_UNIT >> 4);
            if (leader) {
#pragma unroll
              for (uint32_t ks = 0; ks < 4; ++ks)
                if (ks < ksteps && !(a.dbg & 2)) tc::umma2_f16(d, tc::desc_sw128_lo(xh + 2 * ks), tc::desc_sw128_lo(wl + 2 * ks), idesc, 1u);
              tc::umma2_commit(&w_empty[sl]);
              tc::umma2_commit(&x_empty[st]);
            }
            ++u;
          }
        }
        if (leader) tc::umma2_commit(&acc_full[b]);
      }
    } else if (lane == 0) {
      // relay: when an operand has landed in THIS CTA's shared memory, tell the leader's barrier of the same name
      if (n_my > 0) { tc::mbar_wait(wg_full, 0); tc::mbar_arrive_cta(wg_full, 0); }
      uint32_t g = 0, u = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < NSLAB; ++j, ++g) {
          const uint32_t st = g % NST;
          tc::mbar_wait(&x_full[st], (g / NST) & 1);
          tc::mbar_arrive_cta(&x_full[st], 0);
#pragma unroll
          for (int hl = 0; hl < 2; ++hl, ++u) {
            tc::mbar_wait(&w_full[u & 3], (u >> 2) & 1);
            tc::mbar_arrive_cta(&w_full[u & 3], 0);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ---- gate MMA issuer (leader CTA): G = S' Wg^T for both tiles once the epilogue warps of BOTH CTAs have put s' in place -----------------
    if (leader_cta) {
      const bool leader = tc::elect_one();
      const uint32_t idesc = tc::idesc_f16(256, 32);
      const uint32_t wg_lo = tc::smem_u32(wg) >> 4, pk = tc::smem_u32(park) >> 4;
      if (n_my > 0) tc::mbar_wait_cl(wg_full, 0);
      for (int it = 0; it < n_my; ++it) {
        const int b = it & 1;
        tc::mbar_wait_cl(&a_ready[b], (it >> 1) & 1);
        tc::tc_fence_after();
        if (leader) {
          const uint32_t cb = tmem + (uint32_t)(b * 256), dg = cb + GCOL;
          uint32_t first = 0;
#pragma unroll
          for (uint32_t c = 0; c < 8; ++c) {
            if (a.dbg & 16) break;
            const uint32_t sl = c >> 1;                                         // k-slab of the gate weights
#pragma unroll
            for (uint32_t k2 = 0; k2 < 2; ++k2) {
              const uint32_t kb = (c & 1) * 2 + k2;                             // k-step inside the slab
              const uint64_t bh = tc::desc_sw128_lo(wg_lo + (2 * sl) * (2048 >> 4) + 2 * kb);
              const uint64_t bl = tc::desc_sw128_lo(wg_lo + (2 * sl + 1) * (2048 >> 4) + 2 * kb);
              if (c == 7) {                                                     // parked chunk: ordinary shared-memory A operand
                tc::umma2_f16(dg, tc::desc_sw128_lo(pk + (LO_OFF >> 4) + 2 * kb), bh, idesc, first);
                tc::umma2_f16(dg, tc::desc_sw128_lo(pk + 2 * kb), bh, idesc, 1u);
                tc::umma2_f16(dg, tc::desc_sw128_lo(pk + 2 * kb), bl, idesc, 1u);
              } else {
                const uint32_t ah = cb + 32 * c + 8 * k2, al = ah + 16;
                tc::umma2_f16_ts(dg, al, bh, idesc, first);
                tc::umma2_f16_ts(dg, ah, bh, idesc, 1u);
                tc::umma2_f16_ts(dg, ah, bl, idesc, 1u);
              }
              first = 1u;
            }
          }
          tc::umma2_commit(&gate_full[b]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= PL::W_LOAD0 && warp < PL::W_EPI0) {
    // ---- activation loaders: 4 image slabs of s' by bulk TMA, the norms converted; MSGA: destination bookkeeping for the epilogue ------
    const int wrow0 = (warp - PL::W_LOAD0) * 32, lg = lane >> 3, ch = lane & 7;
    int r_ok = 0;
    long long f_slot0 = 0;
    auto rowinfo = [&](int it) {
      f_slot0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * PL::T;
      const int r = wrow0 + lane;
      const long long slot = f_slot0 + r;
      int ok = 0, info = -1;
      if (slot < a.EP) {
        const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
        const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
        if (le < n * (n - 1)) {
          ok = 1;
          if (AGG) {
            const int j = le / (n - 1), rem = le - j * (n - 1);
            info = ((bt.mol_node[mol] + j) << 2) | (rem == n - 2 ? 2 : 0) | (rem == 0 ? 1 : 0);   // dst node | last in-edge | first in-edge
          }
        }
      }
      r_ok = ok;
      if (AGG) {
        int* r_row = reinterpret_cast<int*>(smem_dyn + PL::OFF_ROW) + (it % PL::NROWBUF) * PL::T;
        r_row[r] = info;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&rows_full[it % PL::NROWBUF])) : "memory");
      }
    };
    auto fetch = [&](const int j, float4 (&buf)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + lg;
        const long long sl_ = f_slot0 + wrow0 + rl;
        const bool ok = __shfl_sync(0xffffffffu, r_ok, rl) != 0;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k0 = (j - S / 32) * 32 + ch * 4;
        if (ok && k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
        buf[i] = val;
      }
    };
    float4 cur[8], nxt[8];
    float amax = 0.f;
    if (n_my > 0) { rowinfo(0); fetch(FIRST_CH, cur); }
    uint32_t g = 0;
    for (int it = 0; it < n_my; ++it) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s, ++g) {
        const uint32_t st = g % NST, use = g / NST;
        if (s < NIMG) {
          if (use > 0) tc::mbar_wait(&x_empty[st], (use - 1) & 1);
          if (lane == 0) {
            if (warp == PL::W_LOAD0) {
              const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
              if (a.dbg & 4) tc::mbar_arrive_expect_tx(&x_full[st], 0u);
              else {
              tc::mbar_arrive_expect_tx(&x_full[st], PL::XSTAGE);
              tc::bulk_g2s(xst + st * PL::XSTAGE, reinterpret_cast<const uint8_t*>(a.in_img) + ((size_t)tile * NIMG + s) * PL::XSTAGE,
                           PL::XSTAGE, &x_full[st]);
              }
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
          if (j >= NCH) break;
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
  } else if (warp >= PL::W_EPI0) {
    // ---- epilogue -----------------------------------------------------------------------------------------------------------------------------
    const int q = warp & 3, hf = (warp - PL::W_EPI0) >> 2, row = q * 32 + lane;
    const float unscale = a.units[(size_t)(NSLAB * 4) * (TC_UNIT / 4)];
    const float g_unscale = a.g_units[8 * 1024];
    const uint32_t x7 = (uint32_t)(row & 7);
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float omax = 0.f;
    // gate rows of tile t: bias + sigmoid on the gate accumulator, 128 contiguous bytes per edge; releases the accumulator buffer
    auto gate_epilogue = [&](int t) {
      const int bt_ = t & 1;
      tc::mbar_wait(&gate_full[bt_], (t >> 1) & 1);
      tc::tc_fence_after();
      float gacc[32];
      tc::tmem_ld32(tmem + lane_addr + (uint32_t)(bt_ * 256) + GCOL, gacc);
      tc::tmem_ld_wait();
      const long long slot = (((long long)blockIdx.x + (long long)t * gridDim.x) * PL::T) + row;
      float* gp = a.g_out + (size_t)slot * 32;
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {
        if (a.dbg & (8 | 32)) break;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = sigmoid_fast(gacc[8 * i8 + e] * g_unscale + __ldg(a.g_bias + 8 * i8 + e));
        st_global_256(gp + 8 * i8, make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])),
                      make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cta(&acc_empty[bt_], 0);
    };
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      // Gate rows of the PREVIOUS tile first: its accumulator buffer is the one the main MMAs of tile it + 1 need, and they can only
      // start once these four warps have read G out of it.  (Done one chunk into this tile -- i.e. after acc_full of THIS tile -- the
      // main issuer could never run ahead: MMAs alone took 420 us per launch against a tensor floor of ~260 us, profiles/r02i.)
      if (hf == 0 && it > 0) gate_epilogue(it - 1);
      // MSGA: this lane's row: destination node and whether it is the node's first / last in-edge; segments of the warp's 32 rows
      int info = -1;
      unsigned seg_ends = 0;
      if (AGG) {
        tc::mbar_wait(&rows_full[it % PL::NROWBUF], (it / PL::NROWBUF) & 1);
        info = (reinterpret_cast<const int*>(smem_dyn + PL::OFF_ROW) + (it % PL::NROWBUF) * PL::T)[row];
        const unsigned vmask = __ballot_sync(0xffffffffu, info >= 0);
        const bool next_valid = lane < 31 && ((vmask >> (lane + 1)) & 1u);
        seg_ends = __ballot_sync(0xffffffffu, info >= 0 && ((info & 2) || !next_valid));
      }
      tc::mbar_wait(&acc_full[b], (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t cb = tmem + lane_addr + (uint32_t)(b * 256);
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const int c = 2 * s + hf;                                // hf 0: chunks 0 2 4 6;  hf 1: 1 3 5 7 (the parked chunk last)
        float acc[32];
        tc::tmem_ld32(cb + (uint32_t)(c * 32), acc);
        tc::tmem_ld_wait();
        if (a.dbg & 8) continue;
        uint32_t h2[16], l2[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bq = __ldg(reinterpret_cast<const float4*>(a.bias + c * 32) + i4);
          const float ad[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z = acc[4 * i4 + e] * unscale + ad[e];
            const float o = z * sigmoid_fast(z);
            acc[4 * i4 + e] = o;
            omax = fmaxf(omax, fabsf(o));
          }
          tc::split_h16x2(acc[4 * i4], acc[4 * i4 + 1], h2[2 * i4], l2[2 * i4]);
          tc::split_h16x2(acc[4 * i4 + 2], acc[4 * i4 + 3], h2[2 * i4 + 1], l2[2 * i4 + 1]);
        }
        if (c == 7) {
          // park: row `row` of operand slab 3, pieces 4..7 (k = 32..63), 16 bytes each at position p ^ (row % 8).  The gate MMAs of the
          // previous tile read this buffer: they were committed to gate_full long ago (a chunk of this tile lies in between)
          if (it > 0) tc::mbar_wait(&gate_full[(it - 1) & 1], ((it - 1) >> 1) & 1);
          uint8_t* pr = park + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t pos = ((uint32_t)(4 + j) ^ x7) << 4;
            *reinterpret_cast<uint4*>(pr + pos) = make_uint4(h2[4 * j], h2[4 * j + 1], h2[4 * j + 2], h2[4 * j + 3]);
            *reinterpret_cast<uint4*>(pr + LO_OFF + pos) = make_uint4(l2[4 * j], l2[4 * j + 1], l2[4 * j + 2], l2[4 * j + 3]);
          }
          tc::fence_proxy_async();
        } else {
          tmem_st16(cb + (uint32_t)(c * 32), h2);                 // in place: 16 columns of (hi, hi) pairs, 16 columns of (lo, lo) pairs
          tmem_st16(cb + (uint32_t)(c * 32 + 16), l2);
        }
        if (MODE == EG_MSG && !(a.dbg & 32)) {
          uint8_t* ob = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)tile * (S / 64) + s) * PL::XSTAGE + (size_t)row * 128;
#pragma unroll
          for (int pr_ = 0; pr_ < 2; ++pr_) {
            const uint32_t p0 = (uint32_t)(hf * 4 + 2 * pr_), pos = (p0 ^ x7) & ~1u;
            const bool swap = (x7 & 1u) != 0;
            const uint4 ha = make_uint4(h2[8 * pr_], h2[8 * pr_ + 1], h2[8 * pr_ + 2], h2[8 * pr_ + 3]);
            const uint4 hb = make_uint4(h2[8 * pr_ + 4], h2[8 * pr_ + 5], h2[8 * pr_ + 6], h2[8 * pr_ + 7]);
            const uint4 la = make_uint4(l2[8 * pr_], l2[8 * pr_ + 1], l2[8 * pr_ + 2], l2[8 * pr_ + 3]);
            const uint4 lb = make_uint4(l2[8 * pr_ + 4], l2[8 * pr_ + 5], l2[8 * pr_ + 6], l2[8 * pr_ + 7]);
            st_global_256(ob + pos * 16, swap ? hb : ha, swap ? ha : hb);
            st_global_256(ob + LO_OFF + pos * 16, swap ? lb : la, swap ? la : lb);
          }
        }
        if (AGG && !(a.dbg & 64)) {
          // scalar messages summed over the in-edges of every destination (gvp.py:491): per segment of this warp's 32 rows one masked
          // transposing reduction -- lane i ends up with feature 32 c + i summed over the segment's rows -- stored as a 32-row piece
          const long long t32 = (tile * PL::T + q * 32) >> 5;
          unsigned rem_mask = seg_ends;
          int lo = 0;
          while (rem_mask) {
            const int hi_ = __ffs(rem_mask) - 1;
            rem_mask &= rem_mask - 1;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (lane >= lo && lane <= hi_) ? acc[i] : 0.f;
            const float tot = warp_transpose_sum(v);
            const int i_lo = __shfl_sync(0xffffffffu, info, lo), i_hi = __shfl_sync(0xffffffffu, info, hi_);
            const bool head = (i_lo & 1) != 0, tail = (i_hi & 2) != 0;
            float* dstp = (head && tail) ? a.M + (size_t)(i_hi >> 2) * D::MW : (head ? a.partL + (size_t)t32 * D::MW : a.partF + (size_t)t32 * D::MW);
            dstp[c * 32 + lane] = tot;
            lo = hi_ + 1;
          }
        }
      }
      tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cta(&a_ready[b], 0);
    }
    if (hf == 0 && n_my > 0) gate_epilogue(n_my - 1);
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // nothing of the pair is still aimed at this CTA's barriers / tensor memory
  if (warp == 1) tc::tmem_dealloc2(tmem, 512);
}


}  // namespace fm
