// k_egemm_c -- a CHAIN of two linears with a 128-wide hidden layer in one kernel, edges on M, the hidden activations never leaving
// the SM:   rows -> linear 1 (K1 -> 128) -> add-in + SiLU -> [tensor memory] -> linear 2 (128 -> N2) -> mode epilogue.
//
// What it replaces.  EdgeUpdate (vector_field.py:844-880) ran as two k_egemm_p launches (EG_EU1 + EG_EU2, 375 + 513 us at GEOM-512,
// five times per evaluation): the hidden activations h made a round trip through HBM as operand images (1 KB / edge) and the
// LayerNorm of EU2 -- "features on M": one thread = one feature -- was two transposing 32 x 32 warp reductions plus three named
// barriers per 32 rows.  Here (edges on M: one thread = one edge ROW, as k_egemm_e / k_egemm_g):
//   * h = SiLU(z1) is written back IN PLACE over the fp32 accumulator columns it was read from, as packed fp16 (hi, lo) pairs, and
//     linear 2 takes its A operand from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc) -- no image, no HBM, no shared memory;
//   * linear 2's weights (64 KB of (hi, lo) images) are RESIDENT in shared memory; only linear 1's stream through a ring;
//   * the LayerNorm needs no transposition: a row's 128 features are the accumulator columns of TWO threads (64 each, kept in
//     registers), which exchange their partial sums through shared memory.
// HBM traffic per edge: 512 B edge-feature image in, 512 B fp32 residual in, 512 B fp32 + 512 B image out = 2 KB (was 3 KB + the
// per-launch weight streams).
//
// Roles (16 warps, one CTA per SM, persistent over 128-row tiles):
//   warp 0      lane 0: weight producer (linear 2's units once, linear 1's units through a 3-slot ring, bulk TMA);
//               lane 1: MMA issuer of linear 2 (A = D1's columns re-written as fp16 hi | lo, B = resident unit, D2 = columns [128, 256))
//   warp 1      MMA issuer of linear 1 (A = activation stage, B = weight unit, D1 = accumulator columns [0, 128) of buffer b)
//   warps 2-3   activation loaders: image k-slabs by bulk TMA, computed k-slabs (rbf of the edge length) converted in registers;
//               they also publish the tile's row bookkeeping (source node, dst - src) for the epilogues
//   warps 4-7   epilogue 1: TMEM lane quarter q = warp % 4; z1 = D1 * unscale + EA[src] + EB[dst], h = SiLU(z1) -> TMEM
//   warps 8-15  epilogue 2 (two per lane quarter, 64 features each, values in registers): y = ef + SiLU(D2 * unscale + b2),
//               LayerNorm (partial sums exchanged inside the pair), fp32 rows + operand images of the new edge features
// Two TMEM buffers of 256 columns: while epilogue 2 drains tile t, epilogue 1 works on tile t + 1 and the MMAs of tile t + 2 wait
// only for buffer (t + 2) % 2 to be released by epilogue 2 of tile t.
//
// Arithmetic is k_egemm_p's: fp16x3 operands (same images, same units, same product order lo.hi, hi.hi, hi.lo per k-slab), fp32
// accumulation, the same epilogue expressions; only the LayerNorm sums run in a different order (sequential per row instead of a
// shuffle tree), so results agree with the two-launch path to fp32 rounding, not bit for bit (tests state the tolerance).
#pragma once
#include "egemm_e.cuh"

namespace fm {

enum ChainMode : int { CH_EU = 0 };

struct EgcPlan {
  static constexpr int T = 128;
  static constexpr int NST = 3;                              // activation stages (EU: exactly one tile's three k-slabs)
  static constexpr int XSTAGE = 32768;
  static constexpr int RING = 3;                             // weight ring of linear 1 (16 KB units)
  static constexpr int W2_BYTES = 4 * TC_UNIT;               // linear 2: 2 k-slabs x (hi, lo) x [128 features][64 k]
  static constexpr int NLW = 2, NE1 = 4, NE2 = 8;             // 16 warps = 512 threads: the register file still gives 128 per thread
  static constexpr int THREADS = (2 + NLW + NE1 + NE2) * 32;
  static constexpr int W_LOAD0 = 2, W_E1 = 2 + NLW, W_E2 = W_E1 + NE1;
  static constexpr int NROWBUF = 4;
  static constexpr int ROWBYTES = T * 6;                     // int src[T]; short dd[T]
  static constexpr int OFF_X = 0;
  static constexpr int OFF_RING = NST * XSTAGE;
  static constexpr int OFF_W2 = OFF_RING + RING * TC_UNIT;
  static constexpr int OFF_ROW = OFF_W2 + W2_BYTES;
  static constexpr int OFF_XCH = OFF_ROW + NROWBUF * ROWBYTES;   // LayerNorm partials: 2 exchanges x 8 epilogue-2 warps x 32 lanes
  static constexpr int EB_SLOTS = 4;                         // distinct destination rows an epilogue-1 warp stages per tile
  static constexpr int OFF_EB = OFF_XCH + 2 * NE2 * 32 * 4;  // NE1 x EB_SLOTS x 128 floats
  static constexpr int OFF_PRM = OFF_EB + NE1 * EB_SLOTS * 128 * 4;   // b2 | LayerNorm gamma | beta
  static constexpr int OFF_BAR = OFF_PRM + 3 * 128 * 4;
  static constexpr int NBAR = 2 * RING + 2 * NST + 8 + NROWBUF + 1;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES;
  static_assert(BYTES <= 232448, "227 KB of shared memory per CTA");
  static_assert(OFF_RING % 1024 == 0 && OFF_W2 % 1024 == 0, "operand tiles are 1024-byte aligned");
};

// 32 bytes from global memory in one instruction (a full sector per lane)
__device__ __forceinline__ void ld_global_256(const void* p, float4& lo, float4& hi) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
               : "l"(p));
}
__device__ __forceinline__ void ld_global_256_rw(const void* p, float4& lo, float4& hi) {      // data this kernel also writes
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
               : "l"(p)
               : "memory");
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// tcgen05.ld / st shapes with FOUR THREADS PER ROW (measured mapping, tools/gpu_tmem_shapes.py).  The address names the first of 16
// lanes; thread T holds rows T / 4 and T / 4 + 8 of them.
//   16x256b.xN: columns 8 n + 2 (T % 4) + {0, 1} of repeat n: registers 4 n + {0, 1} (row T / 4), 4 n + {2, 3} (row T / 4 + 8)
//   16x128b.xN: column 4 n + T % 4 of repeat n: registers 2 n (row T / 4), 2 n + 1 (row T / 4 + 8)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                 "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float (&v)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_16x128b_x4(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// QL = 1 ("quad layout", option eu_quad): both epilogues work with FOUR THREADS PER ROW.  With one thread per row every 32-byte
// access of a warp touches 32 different 128-byte lines and the kernel was bound by LSU wavefronts (65 % of peak, profiles/r02_summary).
// The hidden features of linear 1 and the output features of linear 2 arrive PERMUTED (weights.py:quad_perm; packed entries
// EUPD_TC1_HP / EUPD_TC2_HP): accumulator column 8 n + 2 t + c holds physical feature 32 (n / 4) + 8 t + 2 (n % 4) + c, so that the
// thread T % 4 = t of a row's quad owns the 8 consecutive features 32 j + 8 t .. + 7 of every group j of 32.  Every global access
// of the epilogues (EA[src], EB[dst], the residual, the fp32 rows, the 16-byte pieces of the operand images) is then one 32- or
// 16-byte access per lane with the quad covering 128 (64) CONTIGUOUS bytes of the row: a warp instruction touches 8 rows.
// h goes back to tensor memory IN PLACE with the 16x128b shape (packed k-pair column 4 n + t of its group of 32).
template <class D, int MODE, int QL = 0>
__global__ void __launch_bounds__(EgcPlan::THREADS, 1)
k_egemm_c(const ModelRT m, const BatchRT bt, const EgArgs a, const int n_tiles) {
  pdl_launch();
  pdl_wait();
  using PL = EgcPlan;
  static_assert(MODE == CH_EU, "chain modes");
  static_assert(D::F == 128 && D::R == 32, "hidden width 128 = one N = 128 MMA; rbf chunk = 32 k values");
  constexpr int F = D::F;
  constexpr int K1 = F + D::R;                               // ef | rbf(d)
  constexpr int NSLAB = (K1 + 63) / 64;
  constexpr int LAST_KSTEPS = ((K1 - 1) % 64) / 16 + 1;
  constexpr int NIMG = F / 64;
  constexpr int NU1 = NSLAB * 2;                             // linear 1's weight units per tile (hi, lo per k-slab)
  constexpr int NST = PL::NST, RING = PL::RING;
  constexpr int LO_OFF = 16384;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* xst = smem_dyn + PL::OFF_X;
  uint8_t* ring = smem_dyn + PL::OFF_RING;
  uint8_t* w2 = smem_dyn + PL::OFF_W2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + PL::OFF_BAR);
  uint64_t *w_full = bars, *w_empty = bars + RING, *x_full = bars + 2 * RING, *x_empty = x_full + NST;
  uint64_t *d1_full = x_empty + NST, *a_ready = d1_full + 2, *d2_full = a_ready + 2, *buf_empty = d2_full + 2;
  uint64_t *rows_full = buf_empty + 2, *w2_full = rows_full + PL::NROWBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w2_full + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&x_full[i], PL::NLW); tc::mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&d1_full[i], 1);
      tc::mbar_init(&a_ready[i], PL::NE1);
      tc::mbar_init(&d2_full[i], 1);
      tc::mbar_init(&buf_empty[i], PL::NE2);
    }
    for (int i = 0; i < PL::NROWBUF; ++i) tc::mbar_init(&rows_full[i], PL::NLW);
    tc::mbar_init(w2_full, 1);
    tc::fence_mbar_init();
  }
  float* prm = reinterpret_cast<float*>(smem_dyn + PL::OFF_PRM);
  for (int i = tid; i < 3 * F; i += PL::THREADS) prm[i] = i < F ? a.bias[i] : (i < 2 * F ? a.ln_w[i - F] : a.ln_b[i - 2 * F]);
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- weight producer (lane 0) and the issuer of linear 2 (lane 1): two single-thread roles sharing one warp, each in its own
    // divergent branch (independent thread scheduling lets both wait on their mbarriers) -- the warp this frees is a second loader
    if (lane == 0) {
      if (n_my > 0) {
      tc::mbar_arrive_expect_tx(w2_full, PL::W2_BYTES);
      for (int u = 0; u < 4; ++u) tc::bulk_g2s(w2 + u * TC_UNIT, reinterpret_cast<const uint8_t*>(a.g_units) + (size_t)u * TC_UNIT, TC_UNIT, w2_full);
      uint32_t u = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int k = 0; k < NU1; ++k, ++u) {
          const uint32_t sl = u % RING, use = u / RING;
          if (use > 0) tc::mbar_wait(&w_empty[sl], (use - 1) & 1);
          if (a.dbg & 1) { tc::mbar_arrive_expect_tx(&w_full[sl], 0u); continue; }
          tc::mbar_arrive_expect_tx(&w_full[sl], TC_UNIT);
          tc::bulk_g2s(ring + sl * TC_UNIT, reinterpret_cast<const uint8_t*>(a.units) + (size_t)k * TC_UNIT, TC_UNIT, &w_full[sl]);
        }
      }
      }
    } else if (lane == 1) {
      // ---- MMA issuer, linear 2 (one thread): D2 = H W2^T with H (fp16 hi | lo, written by epilogue 1 over D1) as the TMEM A operand ------
      const uint32_t idesc = tc::idesc_f16(128, 128);
      const uint32_t w2_lo = tc::smem_u32(w2) >> 4;
      if (n_my > 0) tc::mbar_wait(w2_full, 0);
      for (int it = 0; it < n_my; ++it) {
        const int b = it & 1;
        tc::mbar_wait(&a_ready[b], (it >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t cb = tmem + (uint32_t)(b * 256), d2 = cb + 128;
        uint32_t first = 0;
#pragma unroll
        for (uint32_t s = 0; s < 2; ++s) {                              // k-slab of W2 = hidden features [64 s, 64 s + 64) = chunks 2 s, 2 s + 1
          if (a.dbg & 16) break;
#pragma unroll
          for (uint32_t kb = 0; kb < 4; ++kb) {
            const uint32_t ah = cb + 32 * (2 * s + (kb >> 1)) + 8 * (kb & 1), al = ah + 16;
            const uint64_t bh = tc::desc_sw128_lo(w2_lo + (2 * s) * (TC_UNIT >> 4) + 2 * kb);
            umma_f16_ts(d2, al, bh, idesc, first);
            umma_f16_ts(d2, ah, bh, idesc, 1u);
            first = 1u;
          }
#pragma unroll
          for (uint32_t kb = 0; kb < 4; ++kb) {
            const uint32_t ah = cb + 32 * (2 * s + (kb >> 1)) + 8 * (kb & 1);
            umma_f16_ts(d2, ah, tc::desc_sw128_lo(w2_lo + (2 * s + 1) * (TC_UNIT >> 4) + 2 * kb), idesc, 1u);
          }
        }
        tc::umma_commit(&d2_full[b]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer, linear 1: D1[128 rows][128] += X[128 rows][64 k] . W1[128][64 k]^T per k-slab, three products ------------------------
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::idesc_f16(128, 128);
    const uint32_t ring_lo = tc::smem_u32(ring) >> 4, x_lo = tc::smem_u32(xst) >> 4;
    uint32_t g = 0, u = 0;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      if (it >= 2) { tc::mbar_wait(&buf_empty[b], ((it >> 1) - 1) & 1); tc::tc_fence_after(); }
      const uint32_t d = tmem + (uint32_t)(b * 256);
      for (int j = 0; j < NSLAB; ++j, ++g) {
        const uint32_t st = g % NST, ksteps = (j == NSLAB - 1) ? LAST_KSTEPS : 4;
        tc::mbar_wait(&x_full[st], (g / NST) & 1);
        const uint32_t xh = x_lo + st * (PL::XSTAGE >> 4), xl = xh + (LO_OFF >> 4);
        {
          const uint32_t sl = u % RING;
          tc::mbar_wait(&w_full[sl], (u / RING) & 1);
          tc::tc_fence_after();
          const uint32_t wh = ring_lo + sl * (TC_UNIT >> 4);
          if (leader) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) {
              if (ks < ksteps && !(a.dbg & 2)) {
                const uint64_t dw = tc::desc_sw128_lo(wh + 2 * ks);
                tc::umma_f16(d, tc::desc_sw128_lo(xl + 2 * ks), dw, idesc, (j > 0 || ks > 0) ? 1u : 0u);
                tc::umma_f16(d, tc::desc_sw128_lo(xh + 2 * ks), dw, idesc, 1u);
              }
            }
            tc::umma_commit(&w_empty[sl]);
          }
          ++u;
        }
        {
          const uint32_t sl = u % RING;
          tc::mbar_wait(&w_full[sl], (u / RING) & 1);
          tc::tc_fence_after();
          const uint32_t wl = ring_lo + sl * (TC_UNIT >> 4);
          if (leader) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
              if (ks < ksteps && !(a.dbg & 2)) tc::umma_f16(d, tc::desc_sw128_lo(xh + 2 * ks), tc::desc_sw128_lo(wl + 2 * ks), idesc, 1u);
            tc::umma_commit(&w_empty[sl]);
            tc::umma_commit(&x_empty[st]);
          }
          ++u;
        }
      }
      if (leader) tc::umma_commit(&d1_full[b]);
    }
  } else if (warp < PL::W_E1) {
    // ---- activation loaders ---------------------------------------------------------------------------------------------------------------------
    // two warps, 64 rows each (lane l <-> rows wrow0 + l and wrow0 + 32 + l): the role is two bulk copies and one computed 32-k chunk
    // per tile; the warps it frees go to epilogue 2
    constexpr int RPL = PL::T / (PL::NLW * 32);
    const int wrow0 = (warp - PL::W_LOAD0) * (32 * RPL), lg = lane >> 3, ch = lane & 7;
    const float inv_sigma = (float)D::R / m.rbf_dmax;
    const float4 mu4 = *reinterpret_cast<const float4*>(m.g(G_RBF_MU) + ch * 4);
    // row bookkeeping of a tile: source node, dst - src, edge length.  A chain of dependent global loads: computed one tile ahead
    // (between the image copies and the rbf slab of the tile before) and published at the top of its tile.
    int n_ok[RPL], n_s[RPL], n_dd[RPL];
    float n_dist[RPL];
    auto rowinfo = [&](int it) {
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const long long slot = ((long long)blockIdx.x + (long long)it * gridDim.x) * PL::T + wrow0 + 32 * k + lane;
        n_ok[k] = 0; n_s[k] = -1; n_dd[k] = 0; n_dist[k] = 0.f;
        if (slot < a.EP) {
          const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
          const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
          if (le < n * (n - 1)) {
            int i, j;
            edge_src_dst(le, n, i, j);
            const int nb = bt.mol_node[mol];
            float dx, dy, dz;
            n_dist[k] = pair_dist(a.x, nb + i, nb + j, dx, dy, dz);
            n_ok[k] = 1;
            n_s[k] = nb + i;
            n_dd[k] = j - i;
          }
        }
      }
    };
    if (n_my > 0) rowinfo(0);
    uint32_t g = 0;
    for (int it = 0; it < n_my; ++it) {
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      int r_ok[RPL];
      float r_dist[RPL];
      {
        uint8_t* rb = smem_dyn + PL::OFF_ROW + (it % PL::NROWBUF) * PL::ROWBYTES;
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          r_ok[k] = n_ok[k]; r_dist[k] = n_dist[k];
          reinterpret_cast<int*>(rb)[wrow0 + 32 * k + lane] = n_s[k];
          reinterpret_cast<short*>(rb + PL::T * 4)[wrow0 + 32 * k + lane] = (short)n_dd[k];
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&rows_full[it % PL::NROWBUF])) : "memory");
      }
#pragma unroll
      for (int s = 0; s < NSLAB; ++s, ++g) {
        const uint32_t st = g % NST, use = g / NST;
        if (use > 0) tc::mbar_wait(&x_empty[st], (use - 1) & 1);
        if (s < NIMG) {
          if (lane == 0) {
            if (warp == PL::W_LOAD0) {
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
        if (it + 1 < n_my) rowinfo(it + 1);
        // rbf(d): k values [0, 32) of the last k-slab; a load-free chunk: lanes 8 g .. 8 g + 7 cover the 32 centres of row 4 i + g
        uint8_t* hi = xst + st * PL::XSTAGE;
        if (!(a.dbg & 512))
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + lg, rr_ = wrow0 + 32 * k + rl;
            const bool ok = __shfl_sync(0xffffffffu, r_ok[k], rl) != 0;
            const float dd = __shfl_sync(0xffffffffu, r_dist[k], rl);
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) val = make_float4(rbf_fast(dd, mu4.x, inv_sigma), rbf_fast(dd, mu4.y, inv_sigma), rbf_fast(dd, mu4.z, inv_sigma),
                                      rbf_fast(dd, mu4.w, inv_sigma));
            uint2 vh, vl;
            tc::split_h16x2(val.x * tc::ACT_SCALE_H16, val.y * tc::ACT_SCALE_H16, vh.x, vl.x);
            tc::split_h16x2(val.z * tc::ACT_SCALE_H16, val.w * tc::ACT_SCALE_H16, vh.y, vl.y);
            const uint32_t off = tc::sw128_off_h(rr_, ch * 4);
            *reinterpret_cast<uint2*>(hi + off) = vh;
            *reinterpret_cast<uint2*>(hi + LO_OFF + off) = vl;
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&x_full[st])) : "memory");
      }
    }
  } else if (QL && warp < PL::W_E2) {
    // ---- epilogue 1, four threads per row: h = SiLU(D1 * unscale + EA[src] + EB[dst]) -> packed fp16 (hi | lo) in tensor memory ------------------
    const int q = warp & 3, tq = lane & 3, rq = lane >> 2;
    const float unscale = a.units[(size_t)NU1 * (TC_UNIT / 4)];
    float omax = 0.f;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      tc::mbar_wait(&rows_full[it % PL::NROWBUF], (it / PL::NROWBUF) & 1);
      const uint8_t* rb = smem_dyn + PL::OFF_ROW + (it % PL::NROWBUF) * PL::ROWBYTES;
      // this thread's four rows: (half, r8) -> row 32 q + 16 half + lane / 4 + 8 r8; EA[src] / EB[dst] at its 8 features of every group
      const float *pa[2][2], *pb[2][2];
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int r8 = 0; r8 < 2; ++r8) {
          const int r = q * 32 + 16 * half + rq + 8 * r8;
          const int sn = max(reinterpret_cast<const int*>(rb)[r], 0);
          const int dn = sn + (int)reinterpret_cast<const short*>(rb + PL::T * 4)[r];
          pa[half][r8] = a.P + (size_t)sn * 2 * F + 8 * tq;
          pb[half][r8] = a.P + (size_t)dn * 2 * F + F + 8 * tq;
        }
      // step = (half, group j of 32 features): one 16x256b.x4 load, 2 rows x (EA, EB) x 32 bytes, one step of loads in flight ahead
      float4 ca[2][2], cbv[2][2], na[2][2], nbv[2][2];
      auto gather = [&](int step, float4 (&ea)[2][2], float4 (&eb)[2][2]) {
        const int half = step >> 2, j = step & 3;
#pragma unroll
        for (int r8 = 0; r8 < 2; ++r8) {
          ld_global_256(pa[half][r8] + 32 * j, ea[r8][0], ea[r8][1]);
          ld_global_256(pb[half][r8] + 32 * j, eb[r8][0], eb[r8][1]);
        }
      };
      if (a.dbg & 128) {                                       // knock-out: barrier protocol only
        tc::mbar_wait(&d1_full[b], (it >> 1) & 1);
        tc::tc_fence_after();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&a_ready[b])) : "memory");
        continue;
      }
      // (Measured and dropped, profiles/r02_kprof_e1ring.txt: EB[dst] from the per-warp staging buffer of the one-thread-per-row path and
      // the EA ring two steps deep -- this role alone is bound by the L2 round trip of its gathers, 362 us per launch in the knock-out
      // timings, but with all roles running the change was neutral to slightly slower.)
      gather(0, ca, cbv);
      tc::mbar_wait(&d1_full[b], (it >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int step = 0; step < 8; ++step) {
        const int half = step >> 2, j = step & 3;
        const uint32_t cbh = tmem + ((uint32_t)(q * 32 + 16 * half) << 16) + (uint32_t)(b * 256);
        float acc[16];
        tmem_ld_16x256b_x4(cbh + (uint32_t)(32 * j), acc);
        if (step + 1 < 8) gather(step + 1, na, nbv);
        tc::tmem_ld_wait();
        uint32_t hw[8], lw[8];
        if (!(a.dbg & 8)) {
#pragma unroll
          for (int r8 = 0; r8 < 2; ++r8) {
            const float ea[8] = {ca[r8][0].x, ca[r8][0].y, ca[r8][0].z, ca[r8][0].w, ca[r8][1].x, ca[r8][1].y, ca[r8][1].z, ca[r8][1].w};
            const float eb[8] = {cbv[r8][0].x, cbv[r8][0].y, cbv[r8][0].z, cbv[r8][0].w, cbv[r8][1].x, cbv[r8][1].y, cbv[r8][1].z, cbv[r8][1].w};
#pragma unroll
            for (int n4 = 0; n4 < 4; ++n4) {                    // repeat n4: accumulator columns 32 j + 8 n4 + 2 t + {0, 1} = features 32 j + 8 t + 2 n4 + {0, 1}
              float o[2];
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const float z = acc[4 * n4 + 2 * r8 + c] * unscale + __fadd_rn(ea[2 * n4 + c], eb[2 * n4 + c]);
                o[c] = z * sigmoid_fast(z);
                omax = fmaxf(omax, fabsf(o[c]));
              }
              tc::split_h16x2(o[0], o[1], hw[2 * n4 + r8], lw[2 * n4 + r8]);
            }
          }
          // in place, as the one-thread-per-row epilogue: the 32 accumulator columns of group j become 16 columns of hi pairs and 16 of
          // lo pairs (packed k-pair column 4 n4 + t of the group); every column of the group has been read by this step's load
          tmem_st_16x128b_x4(cbh + (uint32_t)(32 * j), hw);
          tmem_st_16x128b_x4(cbh + (uint32_t)(32 * j + 16), lw);
        }
#pragma unroll
        for (int r8 = 0; r8 < 2; ++r8)
#pragma unroll
          for (int k = 0; k < 2; ++k) { ca[r8][k] = na[r8][k]; cbv[r8][k] = nbv[r8][k]; }
      }
      tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&a_ready[b])) : "memory");
    }
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  } else if (warp < PL::W_E2) {
    // ---- epilogue 1: h = SiLU(D1 * unscale + EA[src] + EB[dst]) -> fp16 (hi, lo) in place in tensor memory -------------------------------------
    const int q = warp & 3, row = q * 32 + lane;
    const float unscale = a.units[(size_t)NU1 * (TC_UNIT / 4)];
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float* ebw = reinterpret_cast<float*>(smem_dyn + PL::OFF_EB) + (warp - PL::W_E1) * (PL::EB_SLOTS * F);
    float omax = 0.f;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      tc::mbar_wait(&rows_full[it % PL::NROWBUF], (it / PL::NROWBUF) & 1);
      const uint8_t* rb = smem_dyn + PL::OFF_ROW + (it % PL::NROWBUF) * PL::ROWBYTES;
      const int sn = max(reinterpret_cast<const int*>(rb)[row], 0);
      const int dn = sn + (int)reinterpret_cast<const short*>(rb + PL::T * 4)[row];
      const float* pa = a.P + (size_t)sn * 2 * F;             // EA[src]
      const float* pb = a.P + (size_t)dn * 2 * F + F;         // EB[dst]
      // EA rows come from L2 (~800 cycles) and this warp has little company on its scheduler: an 8-deep ring of 8-feature steps (one
      // 32-byte load each) keeps eight steps in flight while one is consumed.  EB[dst] would double that -- but edges are dst-major:
      // the 32 rows of a warp share at most a few destinations, so the warp stages those rows ONCE in shared memory (one coalesced
      // 512-byte load each) and every lane reads its destination's values as a broadcast; only with more than EB_SLOTS distinct
      // destinations (molecules of fewer than ~10 atoms) the lanes gather EB themselves.
      float4 qa[8][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) ld_global_256(pa + 8 * j, qa[j][0], qa[j][1]);
      const int dn_prev = __shfl_up_sync(0xffffffffu, dn, 1);
      const unsigned newd = __ballot_sync(0xffffffffu, lane == 0 || dn != dn_prev);
      const bool staged = __popc(newd) <= PL::EB_SLOTS;
      const float* ebs = ebw + (__popc(newd & (0xffffffffu >> (31 - lane))) - 1) * F;
      if (staged) {
        unsigned rem = newd;
        for (int k = 0; rem; ++k, rem &= rem - 1) {
          const int dk = __shfl_sync(0xffffffffu, dn, __ffs(rem) - 1);
          *reinterpret_cast<float4*>(ebw + k * F + lane * 4) = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)dk * 2 * F + F) + lane);
        }
        __syncwarp();
      }
      tc::mbar_wait(&d1_full[b], (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t cb = tmem + lane_addr + (uint32_t)(b * 256);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float acc[32];
        tc::tmem_ld32(cb + (uint32_t)(c * 32), acc);
        tc::tmem_ld_wait();
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          const int j = 4 * c + sl, rs = j & 7;                  // features [8 j, 8 j + 8), ring slot
          uint32_t h2[4], l2[4];
          if (!(a.dbg & 8)) {
            float4 vbb[2];
            if (staged) { vbb[0] = *reinterpret_cast<const float4*>(ebs + 8 * j); vbb[1] = *reinterpret_cast<const float4*>(ebs + 8 * j + 4); }
            else ld_global_256(pb + 8 * j, vbb[0], vbb[1]);
#pragma unroll
            for (int i4 = 0; i4 < 2; ++i4) {
              const float4 va = qa[rs][i4], vb = vbb[i4];
              const float pre[4] = {__fadd_rn(va.x, vb.x), __fadd_rn(va.y, vb.y), __fadd_rn(va.z, vb.z), __fadd_rn(va.w, vb.w)};
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float z = acc[8 * sl + 4 * i4 + e] * unscale + pre[e];
                o[e] = z * sigmoid_fast(z);
                omax = fmaxf(omax, fabsf(o[e]));
              }
              tc::split_h16x2(o[0], o[1], h2[2 * i4], l2[2 * i4]);
              tc::split_h16x2(o[2], o[3], h2[2 * i4 + 1], l2[2 * i4 + 1]);
            }
            // chunk c: columns [32 c, 32 c + 16) hold the packed hi pairs of its 32 features, [32 c + 16, 32 c + 32) the lo pairs
            tmem_st4(cb + (uint32_t)(c * 32 + sl * 4), h2);
            tmem_st4(cb + (uint32_t)(c * 32 + 16 + sl * 4), l2);
          }
          if (j + 8 < F / 8) ld_global_256(pa + 8 * (j + 8), qa[rs][0], qa[rs][1]);
        }
      }
      tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&a_ready[b])) : "memory");
    }
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  } else if (QL) {
    // ---- epilogue 2, four threads per row: warp (q, hf) owns output features [64 hf, 64 hf + 64) of rows [32 q, 32 q + 32); a thread
    // holds 16 of them (its 8 features of both groups of 32) for each of its four rows -------------------------------------------------------------
    const int q = warp & 3, hf = (warp - PL::W_E2) >> 2, tq = lane & 3, rq = lane >> 2;
    const float unscale = a.g_units[(size_t)4 * (TC_UNIT / 4)];
    float* xs = reinterpret_cast<float*>(smem_dyn + PL::OFF_XCH);
    float* xq = xs + PL::NE2 * 32;
    float omax = 0.f;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (q == 0 && hf == 0 && lane == 0 && it + 1 < n_my) prefetch_l2_bulk(a.in_s + (size_t)(tile + gridDim.x) * PL::T * F, PL::T * F * 4);
      // row step rs = 2 half + r8 -> row 32 q + 16 half + lane / 4 + 8 r8; residual: 2 groups x 32 bytes per row step, two steps in flight
      const size_t fo = (size_t)(hf * 64 + 8 * tq);              // this thread's first feature (group 0)
      auto rowof = [&](int rs) { return q * 32 + 16 * (rs >> 1) + rq + 8 * (rs & 1); };
      float4 rr[2][4];
      auto resid = [&](int rs, float4 (&dst_)[4]) {
        const float* rp = a.in_s + ((size_t)tile * PL::T + rowof(rs)) * F + fo;
        ld_global_256_rw(rp, dst_[0], dst_[1]);
        ld_global_256_rw(rp + 32, dst_[2], dst_[3]);
      };
      if (a.dbg & 256) {                                       // knock-out: barrier protocol only
        tc::mbar_wait(&d2_full[b], (it >> 1) & 1);
        tc::tc_fence_after();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&buf_empty[b])) : "memory");
        continue;
      }
      resid(0, rr[0]);
      resid(1, rr[1]);
      tc::mbar_wait(&d2_full[b], (it >> 1) & 1);
      tc::tc_fence_after();
      float y[2][32];                                          // [half][4 n + 2 r8 + c], n = 4 group + n4
#pragma unroll
      for (int half = 0; half < 2; ++half)
        tmem_ld_16x256b_x8(tmem + ((uint32_t)(q * 32 + 16 * half) << 16) + (uint32_t)(b * 256 + 128 + hf * 64), y[half]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&buf_empty[b])) : "memory");
      float psum[4];
#pragma unroll
      for (int rs = 0; rs < 4; ++rs) {
        const int half = rs >> 1, r8 = rs & 1;
        const float4* rv = rr[rs & 1];
        float sacc = 0.f;
        if (!(a.dbg & 8)) {
#pragma unroll
          for (int gj = 0; gj < 2; ++gj) {
            const float4 b0 = *reinterpret_cast<const float4*>(prm + fo + 32 * gj), b1 = *reinterpret_cast<const float4*>(prm + fo + 32 * gj + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const float rs8[8] = {rv[2 * gj].x, rv[2 * gj].y, rv[2 * gj].z, rv[2 * gj].w, rv[2 * gj + 1].x, rv[2 * gj + 1].y, rv[2 * gj + 1].z, rv[2 * gj + 1].w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {                       // feature fo + 32 gj + i = accumulator column 8 (4 gj + i / 2) + 2 t + i % 2
              const int idx = 4 * (4 * gj + (i >> 1)) + 2 * r8 + (i & 1);
              const float z = y[half][idx] * unscale + bb[i];
              const float v = __fadd_rn(rs8[i], z * sigmoid_fast(z));
              y[half][idx] = v;
              sacc += v;
            }
          }
        }
        psum[rs] = sacc;
        if (rs + 2 < 4) resid(rs + 2, rr[rs & 1]);
      }
      // LayerNorm statistics of a row: its quad (shuffles) x the two warps of the lane quarter (shared memory)
      float mean[4], rstd[4];
#pragma unroll
      for (int rs = 0; rs < 4; ++rs) {
        psum[rs] += __shfl_xor_sync(0xffffffffu, psum[rs], 1);
        psum[rs] += __shfl_xor_sync(0xffffffffu, psum[rs], 2);
        if (tq == 0) xs[(q * 2 + hf) * 32 + 16 * (rs >> 1) + rq + 8 * (rs & 1)] = psum[rs];
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
#pragma unroll
      for (int rs = 0; rs < 4; ++rs) {
        const float other = xs[(q * 2 + (hf ^ 1)) * 32 + 16 * (rs >> 1) + rq + 8 * (rs & 1)];
        mean[rs] = (hf ? other + psum[rs] : psum[rs] + other) * (1.0f / 128.0f);
        float sq = 0.f;
#pragma unroll
        for (int gj = 0; gj < 2; ++gj)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dl = y[rs >> 1][4 * (4 * gj + (i >> 1)) + 2 * (rs & 1) + (i & 1)] - mean[rs];
            sq += dl * dl;
          }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        psum[rs] = sq;
        if (tq == 0) xq[(q * 2 + hf) * 32 + 16 * (rs >> 1) + rq + 8 * (rs & 1)] = sq;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
#pragma unroll
      for (int rs = 0; rs < 4; ++rs) {
        const float other = xq[(q * 2 + (hf ^ 1)) * 32 + 16 * (rs >> 1) + rq + 8 * (rs & 1)];
        rstd[rs] = rsqrtf((hf ? other + psum[rs] : psum[rs] + other) * (1.0f / 128.0f) + 1e-5f);
      }
      if (a.dbg & (8 | 32)) continue;
#pragma unroll
      for (int rs = 0; rs < 4; ++rs) {
        const int r = rowof(rs);
        float* op = a.out + ((size_t)tile * PL::T + r) * F + fo;
        uint8_t* ib = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)tile * (F / 64) + hf) * PL::XSTAGE + (size_t)r * 128;
#pragma unroll
        for (int gj = 0; gj < 2; ++gj) {
          const float4 g0 = *reinterpret_cast<const float4*>(prm + F + fo + 32 * gj), g1 = *reinterpret_cast<const float4*>(prm + F + fo + 32 * gj + 4);
          const float4 t0 = *reinterpret_cast<const float4*>(prm + 2 * F + fo + 32 * gj), t1 = *reinterpret_cast<const float4*>(prm + 2 * F + fo + 32 * gj + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[i] = (y[rs >> 1][4 * (4 * gj + (i >> 1)) + 2 * (rs & 1) + (i & 1)] - mean[rs]) * rstd[rs] * gg[i] + tt[i];
            omax = fmaxf(omax, fabsf(o[i]));
          }
          st_global_256(op + 32 * gj, make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])),
                        make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
          uint32_t h4[4], l4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::split_h16x2(o[2 * k], o[2 * k + 1], h4[k], l4[k]);
          // features 32 gj + 8 t .. + 7 of k-slab hf = its 16-byte piece 4 gj + t
          uint8_t* pp = ib + ((((uint32_t)(4 * gj + tq)) ^ (uint32_t)(r & 7)) << 4);
          *reinterpret_cast<uint4*>(pp) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
          *reinterpret_cast<uint4*>(pp + LO_OFF) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
        }
      }
    }
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  } else {
    // ---- epilogue 2: y = ef + SiLU(D2 * unscale + b2); LayerNorm over the row's 128 features; fp32 row + operand images -------------------------
    // Two warps per TMEM lane quarter: warp (q, hf) owns features [64 hf, 64 hf + 64) of rows [32 q, 32 q + 32) -- one k-slab of the
    // output images.  A thread keeps its 64 y values in registers (ONE pass over tensor memory, the accumulator buffer is released
    // right after it) and the pair exchanges the LayerNorm partial sums through shared memory (two 64-thread named barriers per tile).
    const int q = warp & 3, hf = (warp - PL::W_E2) >> 2, row = q * 32 + lane;
    const float unscale = a.g_units[(size_t)4 * (TC_UNIT / 4)];
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t x7 = (uint32_t)(row & 7);
    float* xs = reinterpret_cast<float*>(smem_dyn + PL::OFF_XCH);
    float* xq = xs + PL::NE2 * 32;
    const int mine = (q * 2 + hf) * 32 + lane, theirs = (q * 2 + (hf ^ 1)) * 32 + lane;
    float omax = 0.f;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long slot = tile * PL::T + row;
      const float* rp = a.in_s + (size_t)slot * F + hf * 64;  // residual: the edge features this kernel overwrites
      float* op = a.out + (size_t)slot * F + hf * 64;
      // the residual rows come from HBM: pull the NEXT tile's 64 KB into L2 now (one bulk-prefetch instruction per tile)
      if (q == 0 && hf == 0 && lane == 0 && it + 1 < n_my) prefetch_l2_bulk(a.in_s + (size_t)(tile + gridDim.x) * PL::T * F, PL::T * F * 4);
      float4 rr[2][4];                                        // residual ring: 2 steps of 16 features in flight
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) { ld_global_256_rw(rp + 16 * sp, rr[sp][0], rr[sp][1]); ld_global_256_rw(rp + 16 * sp + 8, rr[sp][2], rr[sp][3]); }
      tc::mbar_wait(&d2_full[b], (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t d2 = tmem + lane_addr + (uint32_t)(b * 256 + 128 + hf * 64);
      float y[64];
      {
        float (&y0)[32] = *reinterpret_cast<float (*)[32]>(&y[0]);
        float (&y1)[32] = *reinterpret_cast<float (*)[32]>(&y[32]);
        tc::tmem_ld32(d2, y0);
        tc::tmem_ld32(d2 + 32, y1);
        tc::tmem_ld_wait();
      }
      // all of this warp's accumulator columns are in registers: release the buffer
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&buf_empty[b])) : "memory");
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      if (!(a.dbg & 8)) {
#pragma unroll
        for (int sp = 0; sp < 4; ++sp) {                       // 16 features per step
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 bq = *(reinterpret_cast<const float4*>(prm + hf * 64 + sp * 16) + i4);
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
            const float4 rv = rr[sp & 1][i4];
            const float rs[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = y[sp * 16 + 4 * i4 + e] * unscale + bb[e];
              const float v = __fadd_rn(rs[e], z * sigmoid_fast(z));
              y[sp * 16 + 4 * i4 + e] = v;
              sum4[e] += v;
            }
          }
          if (sp + 2 < 4) {
            ld_global_256_rw(rp + 16 * (sp + 2), rr[sp & 1][0], rr[sp & 1][1]);
            ld_global_256_rw(rp + 16 * (sp + 2) + 8, rr[sp & 1][2], rr[sp & 1][3]);
          }
        }
      }
      const float psum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      xs[mine] = psum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float osum = xs[theirs];
      const float mean = (hf ? osum + psum : psum + osum) * (1.0f / 128.0f);
      float sq4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) { const float dl = y[i] - mean; sq4[i & 3] += dl * dl; }
      const float psq = (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]);
      xq[mine] = psq;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float osq = xq[theirs];
      const float rstd = rsqrtf((hf ? osq + psq : psq + osq) * (1.0f / 128.0f) + 1e-5f);
      if (a.dbg & (8 | 32)) continue;
      // operand images of the new edge features: this warp's 64 features are k-slab hf, chunk cc = pieces 4 cc .. 4 cc + 3 (as k_egemm_e)
      uint8_t* ob = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)tile * (F / 64) + hf) * PL::XSTAGE + (size_t)row * 128;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t h2[16], l2[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 gq = *(reinterpret_cast<const float4*>(prm + F + hf * 64 + cc * 32) + i4);
          const float4 tq = *(reinterpret_cast<const float4*>(prm + 2 * F + hf * 64 + cc * 32) + i4);
          const float gg[4] = {gq.x, gq.y, gq.z, gq.w}, tt[4] = {tq.x, tq.y, tq.z, tq.w};
          float* yy = &y[cc * 32 + 4 * i4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            yy[e] = (yy[e] - mean) * rstd * gg[e] + tt[e];
            omax = fmaxf(omax, fabsf(yy[e]));
          }
          tc::split_h16x2(yy[0], yy[1], h2[2 * i4], l2[2 * i4]);
          tc::split_h16x2(yy[2], yy[3], h2[2 * i4 + 1], l2[2 * i4 + 1]);
        }
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {
          const float* yy = &y[cc * 32 + 8 * i8];
          st_global_256(op + cc * 32 + 8 * i8,
                        make_uint4(__float_as_uint(yy[0]), __float_as_uint(yy[1]), __float_as_uint(yy[2]), __float_as_uint(yy[3])),
                        make_uint4(__float_as_uint(yy[4]), __float_as_uint(yy[5]), __float_as_uint(yy[6]), __float_as_uint(yy[7])));
        }
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const uint32_t p0 = (uint32_t)(cc * 4 + 2 * pr), pos = (p0 ^ x7) & ~1u;
          const bool swap = (x7 & 1u) != 0;
          const uint4 ha = make_uint4(h2[8 * pr], h2[8 * pr + 1], h2[8 * pr + 2], h2[8 * pr + 3]);
          const uint4 hb = make_uint4(h2[8 * pr + 4], h2[8 * pr + 5], h2[8 * pr + 6], h2[8 * pr + 7]);
          const uint4 la = make_uint4(l2[8 * pr], l2[8 * pr + 1], l2[8 * pr + 2], l2[8 * pr + 3]);
          const uint4 lb = make_uint4(l2[8 * pr + 4], l2[8 * pr + 5], l2[8 * pr + 6], l2[8 * pr + 7]);
          st_global_256(ob + pos * 16, swap ? hb : ha, swap ? ha : hb);
          st_global_256(ob + LO_OFF + pos * 16, swap ? lb : la, swap ? la : lb);
        }
      }
    }
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace fm
