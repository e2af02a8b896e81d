// k_egemm_h -- k_egemm_g (message linear of GVP 1 / 2 + its gate linear) for the all-image case, with TWELVE epilogue warps.
//
// Once the vector norms arrive as operand images too (k_vecr_b writes the last k-slab, option sh_img) the four loader warps of
// k_egemm_g have nothing left to convert: every k-slab of a tile is one bulk copy.  The kernel is bound by its epilogue (knock-out
// timings, profiles/r03c: MSG epilogue alone 477 us = 284 us of arithmetic + 193 us of image stores, MSGA 514 us = 361 + 150 us of
// segment sums, against 390 us for the MMAs + loads), so the three freed warps join it:
//   warp 0      lane 0: weight producer;  lane 1: activation producer (five bulk copies per tile) -- two single-thread roles in
//               divergent branches of one warp (independent thread scheduling; both only wait on mbarriers and issue bulk copies)
//   warp 1      main MMA issuer (as k_egemm_g)
//   warp 2      gate MMA issuer; MSGA: the whole warp also computes the destination bookkeeping of the tile two ahead
//   warps 3-14  epilogue: TMEM lane quarter q = warp % 4 -- three warps per quarter, all three on the quarter's own scheduler;
//               warp (q, hf) takes chunks hf, hf + 3, hf + 6 (3 / 3 / 2 chunks: the hf = 2 warps also write the gate rows)
// Same MMAs, same epilogue arithmetic per element, same 32-row aggregation pieces as k_egemm_g: bit-identical results.
#pragma once
#include "egemm_e.cuh"

namespace fm {

struct EghPlan {
  static constexpr int T = 128;
  static constexpr int NST = 3;
  static constexpr int XSTAGE = 32768;
  static constexpr int RING_BYTES = 4 * TC_UNIT;
  static constexpr int WG_BYTES = 8 * 4096;
  static constexpr int PARK_BYTES = 32768;
  static constexpr int NEW = 12;
  static constexpr int THREADS = (3 + NEW) * 32;
  static constexpr int W_EPI0 = 3;
  static constexpr int NROWBUF = 3;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_RING = NST * XSTAGE;
  static constexpr int OFF_WG = OFF_RING + RING_BYTES;
  static constexpr int OFF_PARK = OFF_WG + WG_BYTES;
  static constexpr int OFF_ROW = OFF_PARK + PARK_BYTES;
  static constexpr int OFF_BAR = OFF_ROW + NROWBUF * T * 4;
  static constexpr int NBAR = 8 + 2 * NST + 4 + NROWBUF + 2 + 2 + 1;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES;
  static_assert(BYTES <= 232448, "227 KB of shared memory per CTA");
  static_assert(OFF_PARK % 1024 == 0 && OFF_WG % 1024 == 0, "operand tiles are 1024-byte aligned");
};

// 16 lanes x 128 bits, four repeats: thread T receives rows T / 4 and T / 4 + 8 of the 16 lanes, column 4 n + T % 4 of repeat n
// (registers 2 n, 2 n + 1) -- measured mapping, tools/gpu_tmem_shapes.py
__device__ __forceinline__ void tmem_ld_16x128b_x4(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// PERM = 1 (EG_MSG, option eg_perm): the weight units / bias / gate units arrive with the output features permuted inside every
// 32-feature chunk (weights.py:feature_perm).  The epilogue arithmetic is unchanged (it works on accumulator columns), the packed
// fp16 (hi | lo) words it writes back into tensor memory are the gate GEMM's A operand as before -- and ALSO the source of the
// image stores: re-read with the 16x128b shape a thread's four words of a row are one 16-byte piece of the operand image, a
// quad writes 64 contiguous bytes of a row and a warp instruction touches 8 rows instead of 32 (the image stores were 193 of the
// epilogue's 477 us, LSU-wavefront bound: profiles/r03c).
template <class D, int MODE, int PERM = 0>
__global__ void __launch_bounds__(EghPlan::THREADS, 1)
k_egemm_h(const ModelRT m, const BatchRT bt, const EgArgs a, const int n_tiles) {
  pdl_launch();
  pdl_wait();
  using PL = EghPlan;
  static_assert(MODE == EG_MSG || MODE == EG_MSGA, "gate-fused message linears of GVP 1 (image out) and GVP 2 (segment sum)");
  static_assert(!PERM || MODE == EG_MSG, "permuted features: the linear that writes operand images");
  static_assert(D::S == 256, "256 output features, eight 32-feature chunks");
  constexpr int S = D::S;
  constexpr int K = D::K1;
  constexpr int NSLAB = (K + 63) / 64;
  constexpr int LAST_KSTEPS = ((K - 1) % 64) / 16 + 1;
  constexpr int NST = PL::NST;
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
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&x_full[i], 1); tc::mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);                       // the four hf = 2 warps, which read the gate accumulator
      tc::mbar_init(&a_ready[i], PL::NEW);
      tc::mbar_init(&gate_full[i], 1);
    }
    for (int i = 0; i < PL::NROWBUF; ++i) tc::mbar_init(&rows_full[i], 1);
    tc::mbar_init(wg_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- weight producer (+ the resident gate weights, once): as k_egemm_g ------------------------------------------------------------
      if (n_my > 0) {
        tc::mbar_arrive_expect_tx(wg_full, PL::WG_BYTES);
        for (int u = 0; u < 8; ++u) tc::bulk_g2s(wg + u * 4096, reinterpret_cast<const uint8_t*>(a.g_units) + (size_t)u * 4096, 4096, wg_full);
      }
      uint32_t g = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < NSLAB; ++j, ++g) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.units) + (size_t)(4 * j) * TC_UNIT;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (g > 0) tc::mbar_wait(&w_empty[u], (g - 1) & 1);
            if (a.dbg & 1) { tc::mbar_arrive_expect_tx(&w_full[u], 0u); continue; }
            tc::mbar_arrive_expect_tx(&w_full[u], TC_UNIT);
            tc::bulk_g2s(ring + u * TC_UNIT, src + u * TC_UNIT, TC_UNIT, &w_full[u]);
          }
        }
      }
    } else if (lane == 1) {
      // ---- activation producer: the four image k-slabs of s' and the norms slab, one 32 KB bulk copy each ----------------------------------
      uint32_t g = 0;
      for (int it = 0; it < n_my; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
        for (int s = 0; s < NSLAB; ++s, ++g) {
          const uint32_t st = g % NST, use = g / NST;
          if (use > 0) tc::mbar_wait(&x_empty[st], (use - 1) & 1);
          if (a.dbg & 4) { tc::mbar_arrive_expect_tx(&x_full[st], 0u); continue; }
          tc::mbar_arrive_expect_tx(&x_full[st], PL::XSTAGE);
          const uint8_t* srcp = (s == NSLAB - 1) ? reinterpret_cast<const uint8_t*>(a.sh_img) + (size_t)tile * PL::XSTAGE
                                                 : reinterpret_cast<const uint8_t*>(a.in_img) + ((size_t)tile * (S / 64) + s) * PL::XSTAGE;
          tc::bulk_g2s(xst + st * PL::XSTAGE, srcp, PL::XSTAGE, &x_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ---- main MMA issuer ---------------------------------------------------------------------------------------------------------------
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::idesc_f16(128, 128);          // two N = 128 MMAs per k-step and product: one per 128-feature weight unit
    const uint32_t w_base = tc::smem_u32(ring) >> 4, x_lo = tc::smem_u32(xst) >> 4;
    uint32_t g = 0;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      if (it >= 2) { tc::mbar_wait(&acc_empty[b], ((it >> 1) - 1) & 1); tc::tc_fence_after(); }
      for (int j = 0; j < NSLAB; ++j, ++g) {
        const uint32_t st = g % NST, ksteps = (j == NSLAB - 1) ? LAST_KSTEPS : 4;
        tc::mbar_wait(&x_full[st], (g / NST) & 1);
        const uint32_t xh = x_lo + st * (PL::XSTAGE >> 4), xl = xh + (LO_OFF >> 4);
#pragma unroll
        for (uint32_t ft = 0; ft < 2; ++ft) {
          const uint32_t d = tmem + (uint32_t)(b * 256) + ft * 128;
          const uint32_t wh = w_base + (2 * ft) * (TC_UNIT >> 4), wl = wh + (TC_UNIT >> 4);
          tc::mbar_wait(&w_full[2 * ft], g & 1);
          tc::tc_fence_after();
          if (leader) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) {
              if (ks < ksteps && !(a.dbg & 2)) {
                const uint64_t dw = tc::desc_sw128_lo(wh + 2 * ks);
                tc::umma_f16(d, tc::desc_sw128_lo(xl + 2 * ks), dw, idesc, (j > 0 || ks > 0) ? 1u : 0u);
                tc::umma_f16(d, tc::desc_sw128_lo(xh + 2 * ks), dw, idesc, 1u);
              }
            }
            tc::umma_commit(&w_empty[2 * ft]);
          }
          tc::mbar_wait(&w_full[2 * ft + 1], g & 1);
          tc::tc_fence_after();
          if (leader) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
              if (ks < ksteps && !(a.dbg & 2)) tc::umma_f16(d, tc::desc_sw128_lo(xh + 2 * ks), tc::desc_sw128_lo(wl + 2 * ks), idesc, 1u);
            tc::umma_commit(&w_empty[2 * ft + 1]);
          }
        }
        if (leader) tc::umma_commit(&x_empty[st]);
      }
      if (leader) tc::umma_commit(&acc_full[b]);
    }
  } else if (warp == 2) {
    // ---- gate MMA issuer; MSGA: destination bookkeeping (node, first / last in-edge flags) of every row, two tiles ahead -------------------
    auto rowinfo = [&](int it) {
      int* r_row = reinterpret_cast<int*>(smem_dyn + PL::OFF_ROW) + (it % PL::NROWBUF) * PL::T;
#pragma unroll
      for (int k = 0; k < PL::T / 32; ++k) {
        const int r = 32 * k + lane;
        const long long slot = ((long long)blockIdx.x + (long long)it * gridDim.x) * PL::T + r;
        int info = -1;
        if (slot < a.EP) {
          const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
          const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
          if (le < n * (n - 1)) {
            const int j = le / (n - 1), rem = le - j * (n - 1);
            info = ((bt.mol_node[mol] + j) << 2) | (rem == n - 2 ? 2 : 0) | (rem == 0 ? 1 : 0);   // dst node | last in-edge | first in-edge
          }
        }
        r_row[r] = info;
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&rows_full[it % PL::NROWBUF])) : "memory");
    };
    if (AGG) {
      if (n_my > 0) rowinfo(0);
      if (n_my > 1) rowinfo(1);
    }
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::idesc_f16(128, 32);
    const uint32_t wg_lo = tc::smem_u32(wg) >> 4, pk = tc::smem_u32(park) >> 4;
    if (n_my > 0) tc::mbar_wait(wg_full, 0);
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      // rows of tile it + 2 go into the ring slot of tile it - 1, whose epilogue is over (a_ready of tile it - 1 has been seen)
      if (AGG && it + 2 < n_my) rowinfo(it + 2);
      tc::mbar_wait(&a_ready[b], (it >> 1) & 1);
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
            const uint64_t bh = tc::desc_sw128_lo(wg_lo + (2 * sl) * (4096 >> 4) + 2 * kb);
            const uint64_t bl = tc::desc_sw128_lo(wg_lo + (2 * sl + 1) * (4096 >> 4) + 2 * kb);
            if (c == 7) {                                                     // parked chunk: ordinary shared-memory A operand
              tc::umma_f16(dg, tc::desc_sw128_lo(pk + (LO_OFF >> 4) + 2 * kb), bh, idesc, first);
              tc::umma_f16(dg, tc::desc_sw128_lo(pk + 2 * kb), bh, idesc, 1u);
              tc::umma_f16(dg, tc::desc_sw128_lo(pk + 2 * kb), bl, idesc, 1u);
            } else {
              const uint32_t ah = cb + 32 * c + 8 * k2, al = ah + 16;
              umma_f16_ts(dg, al, bh, idesc, first);
              umma_f16_ts(dg, ah, bh, idesc, 1u);
              umma_f16_ts(dg, ah, bl, idesc, 1u);
            }
            first = 1u;
          }
        }
        tc::umma_commit(&gate_full[b]);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue (twelve warps) ------------------------------------------------------------------------------------------------------------------
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
        for (int e = 0; e < 8; ++e) {
          const float zg = gacc[8 * i8 + e] * g_unscale + __ldg(a.g_bias + 8 * i8 + e);
          o[e] = (a.flags & EGF_IDENTITY) ? zg : sigmoid_fast(zg);
        }
        st_global_256(gp + 8 * i8, make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])),
                      make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&acc_empty[bt_])) : "memory");
    };
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      // Gate rows of the PREVIOUS tile first: its accumulator buffer is the one the main MMAs of tile it + 1 need, and they can only
      // start once these four warps have read G out of it.  (Done one chunk into this tile -- i.e. after acc_full of THIS tile -- the
      // main issuer could never run ahead: MMAs alone took 420 us per launch against a tensor floor of ~260 us, profiles/r02i.)
      if (hf == 2 && it > 0) gate_epilogue(it - 1);
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
      for (int c = hf; c < 8; c += 3) {                          // hf 0: chunks 0 3 6;  hf 1: 1 4 7 (the parked chunk last);  hf 2: 2 5 + the gate rows
        const int s = c >> 1, ch_ = c & 1;                       // k-slab of the output images and the 32-feature half inside it
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
        if (MODE == EG_MSG && PERM && !(a.dbg & 32)) {
          uint8_t* slab = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)tile * (S / 64) + s) * PL::XSTAGE;
          if (c == 7) {
            // parked chunk: its words never reach tensor memory; piece j (physical features 8 j .. 8 j + 7 of the chunk) = the words
            // of the accumulator columns 8 n + 2 j, + 1 (n = 0..3): row-per-thread stores as before, two pieces per 32 bytes
            uint8_t* ob = slab + (size_t)row * 128;
#pragma unroll
            for (int pr_ = 0; pr_ < 2; ++pr_) {
              const uint32_t p0 = (uint32_t)(ch_ * 4 + 2 * pr_), pos = (p0 ^ x7) & ~1u;
              const bool swap = (x7 & 1u) != 0;
              const int j0 = 2 * pr_, j1 = 2 * pr_ + 1;
              const uint4 ha = make_uint4(h2[j0], h2[4 + j0], h2[8 + j0], h2[12 + j0]), hb = make_uint4(h2[j1], h2[4 + j1], h2[8 + j1], h2[12 + j1]);
              const uint4 la = make_uint4(l2[j0], l2[4 + j0], l2[8 + j0], l2[12 + j0]), lb = make_uint4(l2[j1], l2[4 + j1], l2[8 + j1], l2[12 + j1]);
              st_global_256(ob + pos * 16, swap ? hb : ha, swap ? ha : hb);
              st_global_256(ob + LO_OFF + pos * 16, swap ? lb : la, swap ? la : lb);
            }
          } else {
            tmem_st_wait();
            const int tq = lane & 3, rq = lane >> 2;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t ta = tmem + ((uint32_t)(q * 32 + 16 * half) << 16) + (uint32_t)(b * 256 + c * 32);
              uint32_t wh_[8], wl_[8];
              tmem_ld_16x128b_x4(ta, wh_);
              tmem_ld_16x128b_x4(ta + 16, wl_);
              tc::tmem_ld_wait();
#pragma unroll
              for (int r8 = 0; r8 < 2; ++r8) {
                const int rr_ = q * 32 + 16 * half + rq + 8 * r8;                         // row of the tile
                uint8_t* dstp = slab + (size_t)rr_ * 128 + ((((uint32_t)(ch_ * 4 + tq)) ^ (uint32_t)(rr_ & 7)) << 4);
                *reinterpret_cast<uint4*>(dstp) = make_uint4(wh_[r8], wh_[2 + r8], wh_[4 + r8], wh_[6 + r8]);
                *reinterpret_cast<uint4*>(dstp + LO_OFF) = make_uint4(wl_[r8], wl_[2 + r8], wl_[4 + r8], wl_[6 + r8]);
              }
            }
          }
        } else
        if (MODE == EG_MSG && !(a.dbg & 32)) {
          uint8_t* ob = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)tile * (S / 64) + s) * PL::XSTAGE + (size_t)row * 128;
#pragma unroll
          for (int pr_ = 0; pr_ < 2; ++pr_) {
            const uint32_t p0 = (uint32_t)(ch_ * 4 + 2 * pr_), pos = (p0 ^ x7) & ~1u;
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
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&a_ready[b])) : "memory");
    }
    if (hf == 2 && n_my > 0) gate_epilogue(n_my - 1);
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace fm
