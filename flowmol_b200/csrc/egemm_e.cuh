// k_egemm_e -- the message linears of the image chain with EDGES ON M (accumulator lane = edge, column = feature).
//
// Why a second orientation.  ncu on k_egemm_p (profiles/r01s, r02): 47 % issue utilisation with two epilogue warps per scheduler,
// ~22 instructions per output element -- the kernel is bound by its epilogue, and most of that epilogue is the price of
// "features on M": a thread owns ONE feature of 64 edges, so the K-major operand image of the next linear (row = edge, 64
// consecutive features per 128 bytes) has to be assembled with a lane-pair shuffle, four byte permutes and two 4-byte stores per
// element pair, and the per-node pre-activations P[src] arrive as 32 scalar loads per chunk.  With edges on M a thread owns one
// EDGE ROW and 32 consecutive features per tcgen05.ld: the image row segment is a run of its own registers (two 32-byte stores
// for hi, two for lo), P[src] is eight 16-byte loads, the bias is a shared-memory broadcast.  The MMAs become
//     D[128 edges][256 features] += X[128 edges][64 k] . W[256 features][64 k]^T       (M = 128, N = 256, one instruction per k-step
// and product instead of two), with the activation stage as UMMA A and one 32 KB weight tile [256 features][64 k] as UMMA B.
//
// Same operand images, same weight units, same three products per k-step in the same order as k_egemm_p (lo.hi, hi.hi, hi.lo per
// accumulator), so the results are expected to agree bit for bit with the other orientation (tests compare them).
//
// Roles (15 warps, one CTA per SM, persistent over 128-edge tiles):
//   warp 0      weight producer: the four 16 KB units of a k-slab land as [f 0..127 hi | f 128..255 hi | f 0..127 lo | f 128..255 lo]
//   warp 1      MMA issuer (one elected lane), two TMEM accumulator buffers of 256 columns
//   warps 3-6   activation loaders: image k-slabs by bulk TMA, everything else converted fp32 -> fp16 (hi, lo) (as k_egemm_p)
//   warps 7-14  epilogue: lane quarter q = warp % 4 (32 edges), the two warps of a quarter take alternate 32-feature chunks
#pragma once
#include "egemm_p.cuh"

namespace fm {

struct EgePlan {
  static constexpr int T = 128;
  static constexpr int NST = 4;
  static constexpr int XSTAGE = 32768;
  static constexpr int RING_BYTES = 4 * TC_UNIT;            // one k-slab of weights
  static constexpr int NLW = 4, NEW = 8;
  static constexpr int THREADS = (3 + NLW + NEW) * 32;
  static constexpr int W_LOAD0 = 3, W_EPI0 = 3 + NLW;
  static constexpr int NROWBUF = 4;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_RING = NST * XSTAGE;
  static constexpr int OFF_ROW = OFF_RING + RING_BYTES;     // NROWBUF x int src[T]
  static constexpr int OFF_BIAS = OFF_ROW + NROWBUF * T * 4;
  static constexpr int OFF_BAR = OFF_BIAS + 256 * 4;
  static constexpr int NBAR = 4 + 2 * NST + 4 + NROWBUF;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES;
  static_assert(BYTES <= 232448, "227 KB of shared memory per CTA");
};

// 32 bytes to global memory in one instruction (a full sector per lane)
__device__ __forceinline__ void st_global_256(void* p, uint4 lo, uint4 hi) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x),
               "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

template <class D, int MODE>
__global__ void __launch_bounds__(EgePlan::THREADS, 1)
k_egemm_e(const ModelRT m, const BatchRT bt, const EgArgs a, const int n_tiles) {
  pdl_launch();
  pdl_wait();
  using PL = EgePlan;
  static_assert(MODE == EG_MSG0 || MODE == EG_MSG, "edges-on-M orientation: the first two message linears (image in, image out)");
  static_assert(D::S == 256 && D::F % 64 == 0, "256 output features = one N = 256 MMA");
  constexpr int S = D::S;
  constexpr int K = MODE == EG_MSG0 ? D::KE0 : D::K1;
  constexpr int NSLAB = (K + 63) / 64;
  constexpr int NCH = (K + 31) / 32;
  constexpr int LAST_KSTEPS = ((K - 1) % 64) / 16 + 1;
  constexpr int NST = PL::NST;
  constexpr int NIMG = MODE == EG_MSG0 ? D::F / 64 : S / 64;
  constexpr int FIRST_CH = 2 * NIMG;
  static_assert(FIRST_CH < NCH, "the loaders publish the row bookkeeping from their first converted chunk");
  constexpr int SH_W = 40;
  constexpr int LO_OFF = 16384;
  constexpr bool NEED_ROWS = MODE == EG_MSG0;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* xst = smem_dyn + PL::OFF_X;
  uint8_t* ring = smem_dyn + PL::OFF_RING;
  float* bias_s = reinterpret_cast<float*>(smem_dyn + PL::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + PL::OFF_BAR);
  // w_full / w_empty [0]: the hi pair (ring bytes [0, 32 K)), [1]: the lo pair
  uint64_t *w_full = bars, *w_empty = bars + 2, *x_full = bars + 4, *x_empty = x_full + NST;
  uint64_t *acc_full = x_empty + NST, *acc_empty = acc_full + 2, *rows_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rows_full + PL::NROWBUF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&x_full[i], PL::NLW); tc::mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], PL::NEW); }
    for (int i = 0; i < PL::NROWBUF; ++i) tc::mbar_init(&rows_full[i], PL::NLW);
    tc::fence_mbar_init();
  }
  if (MODE != EG_MSG0) {
    for (int i = tid; i < S; i += PL::THREADS) bias_s[i] = a.bias[i];
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- weight producer: stream order per slab is (f 0..127 hi, f 0..127 lo, f 128..255 hi, f 128..255 lo) -----------------------
    if (lane == 0) {
      uint32_t g = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < NSLAB; ++j, ++g) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.units) + (size_t)(4 * j) * TC_UNIT;
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {                 // 0: the two hi units -> ring [0, 32 K), 1: the two lo units -> [32 K, 64 K)
            if (g > 0) tc::mbar_wait(&w_empty[hl], (g - 1) & 1);
            tc::mbar_arrive_expect_tx(&w_full[hl], 2 * TC_UNIT);
            tc::bulk_g2s(ring + (2 * hl) * TC_UNIT, src + hl * TC_UNIT, TC_UNIT, &w_full[hl]);
            tc::bulk_g2s(ring + (2 * hl + 1) * TC_UNIT, src + (2 + hl) * TC_UNIT, TC_UNIT, &w_full[hl]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ------------------------------------------------------------------------------------------------------------------
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::idesc_f16(128, 256);
    const uint32_t w_hi = tc::smem_u32(ring) >> 4, w_lo = w_hi + ((2 * TC_UNIT) >> 4), x_lo = tc::smem_u32(xst) >> 4;
    uint32_t g = 0;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      if (it >= 2) { tc::mbar_wait(&acc_empty[b], ((it >> 1) - 1) & 1); tc::tc_fence_after(); }
      const uint32_t d = tmem + (uint32_t)(b * 256);
      for (int j = 0; j < NSLAB; ++j, ++g) {
        const uint32_t st = g % NST, ksteps = (j == NSLAB - 1) ? LAST_KSTEPS : 4;
        tc::mbar_wait(&x_full[st], (g / NST) & 1);
        const uint32_t xh = x_lo + st * (PL::XSTAGE >> 4), xl = xh + (LO_OFF >> 4);
        tc::mbar_wait(&w_full[0], g & 1);
        tc::tc_fence_after();
        if (leader) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) {
              const uint64_t dw = tc::desc_sw128_lo(w_hi + 2 * ks);
              tc::umma_f16(d, tc::desc_sw128_lo(xl + 2 * ks), dw, idesc, (j > 0 || ks > 0) ? 1u : 0u);
              tc::umma_f16(d, tc::desc_sw128_lo(xh + 2 * ks), dw, idesc, 1u);
            }
          }
          tc::umma_commit(&w_empty[0]);
        }
        tc::mbar_wait(&w_full[1], g & 1);
        tc::tc_fence_after();
        if (leader) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            if (ks < ksteps) tc::umma_f16(d, tc::desc_sw128_lo(xh + 2 * ks), tc::desc_sw128_lo(w_lo + 2 * ks), idesc, 1u);
          tc::umma_commit(&w_empty[1]);
          tc::umma_commit(&x_empty[st]);
        }
      }
      if (leader) tc::umma_commit(&acc_full[b]);
    }
  } else if (warp >= PL::W_LOAD0 && warp < PL::W_EPI0) {
    // ---- activation loaders (k_egemm_p's: image slabs by bulk TMA, the rest converted; row bookkeeping published for the epilogue) ---
    const int wrow0 = (warp - PL::W_LOAD0) * 32, lg = lane >> 3, ch = lane & 7;
    const float inv_sigma = (float)D::R / m.rbf_dmax;
    const float4 mu4 = *reinterpret_cast<const float4*>(m.g(G_RBF_MU) + ch * 4);
    int r_ok = 0;
    float r_dist = 0.f;
    long long f_slot0 = 0;
    auto rowinfo = [&](int it) {
      f_slot0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * PL::T;
      const int r = wrow0 + lane;
      const long long slot = f_slot0 + r;
      int ok = 0, s = -1;
      float dist = 0.f;
      if (slot < a.EP) {
        const int t64 = (int)(slot >> 6), mol = bt.etile_mol[t64];
        const int n = bt.mol_n[mol], le = (int)(slot - ((long long)bt.mol_etile[mol] << 6));
        if (le < n * (n - 1)) {
          ok = 1;
          if (MODE == EG_MSG0) {
            int i, j;
            edge_src_dst(le, n, i, j);
            const int nb = bt.mol_node[mol];
            float dx, dy, dz;
            dist = pair_dist(a.x, nb + i, nb + j, dx, dy, dz);
            s = nb + i;
          }
        }
      }
      r_ok = ok;
      r_dist = dist;
      if (NEED_ROWS) {
        int* r_src = reinterpret_cast<int*>(smem_dyn + PL::OFF_ROW) + (it % PL::NROWBUF) * PL::T;
        r_src[r] = s;
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&rows_full[it % PL::NROWBUF])) : "memory");
      }
    };
    // k order: MSG0  ef(F) | rbf(32) | norms;  MSG  s'(S) | norms
    auto fetch = [&](const int j, float4 (&buf)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + lg;
        const long long sl_ = f_slot0 + wrow0 + rl;
        const bool ok = __shfl_sync(0xffffffffu, r_ok, rl) != 0;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == EG_MSG0) {
          const float dd = __shfl_sync(0xffffffffu, r_dist, rl);
          if (j == D::F / 32) {
            if (ok) val = make_float4(rbf_fast(dd, mu4.x, inv_sigma), rbf_fast(dd, mu4.y, inv_sigma), rbf_fast(dd, mu4.z, inv_sigma),
                                      rbf_fast(dd, mu4.w, inv_sigma));
          } else {
            const int k0 = (j - 1 - D::F / 32) * 32 + ch * 4;
            if (ok && k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
          }
        } else {
          const int k0 = (j - S / 32) * 32 + ch * 4;
          if (ok && k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
        }
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
    // ---- epilogue: thread = edge row (TMEM lane), 32 consecutive features per tcgen05.ld ---------------------------------------------------
    const int q = warp & 3, hf = (warp - PL::W_EPI0) >> 2, row = q * 32 + lane;
    const float unscale = a.units[(size_t)(NSLAB * 4) * (TC_UNIT / 4)];
    const uint32_t x7 = (uint32_t)(row & 7);
    float omax = 0.f;
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
      const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      const float4* prow = nullptr;
      if (NEED_ROWS) {
        tc::mbar_wait(&rows_full[it % PL::NROWBUF], (it / PL::NROWBUF) & 1);
        const int* r_src = reinterpret_cast<const int*>(smem_dyn + PL::OFF_ROW) + (it % PL::NROWBUF) * PL::T;
        prow = reinterpret_cast<const float4*>(a.P + (size_t)max(r_src[row], 0) * S);
      }
      float4 pre[8], pnext[8];
      if (MODE == EG_MSG0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pre[i] = __ldg(prow + hf * 8 + i);
      }
      tc::mbar_wait(&acc_full[b], (it >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < S / 64; ++s) {
        const int c = 2 * s + hf;                                // this warp's 32-feature chunk of slab s
        float acc[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + c * 32), acc);
        if (MODE == EG_MSG0 && s + 1 < S / 64) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pnext[i] = __ldg(prow + (c + 2) * 8 + i);
        }
        tc::tmem_ld_wait();
        uint32_t h2[16], l2[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float4 add;
          if (MODE == EG_MSG0) add = pre[i4];
          else add = *reinterpret_cast<const float4*>(bias_s + c * 32 + i4 * 4);
          const float ad[4] = {add.x, add.y, add.z, add.w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z = acc[4 * i4 + e] * unscale + ad[e];
            o[e] = z * sigmoid_fast(z);
            omax = fmaxf(omax, fabsf(o[e]));
          }
          tc::split_h16x2(o[0], o[1], h2[2 * i4], l2[2 * i4]);
          tc::split_h16x2(o[2], o[3], h2[2 * i4 + 1], l2[2 * i4 + 1]);
        }
        // image row of this edge in slab s: 128 bytes = 8 pieces of 16 bytes (8 k values), piece p stored at position p ^ (row % 8);
        // this warp owns pieces 4 hf .. 4 hf + 3, i.e. two 32-byte sectors {p, p ^ 1}
        uint8_t* ob = reinterpret_cast<uint8_t*>(a.out_img) + ((size_t)tile * (S / 64) + s) * PL::XSTAGE + (size_t)row * 128;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const uint32_t p0 = (uint32_t)(hf * 4 + 2 * pr), pos = (p0 ^ x7) & ~1u;       // sector start (in 16-byte pieces)
          const bool swap = (x7 & 1u) != 0;                       // odd rows: piece p0 sits in the upper half of its sector
          const uint4 ha = make_uint4(h2[8 * pr], h2[8 * pr + 1], h2[8 * pr + 2], h2[8 * pr + 3]);
          const uint4 hb = make_uint4(h2[8 * pr + 4], h2[8 * pr + 5], h2[8 * pr + 6], h2[8 * pr + 7]);
          const uint4 la = make_uint4(l2[8 * pr], l2[8 * pr + 1], l2[8 * pr + 2], l2[8 * pr + 3]);
          const uint4 lb = make_uint4(l2[8 * pr + 4], l2[8 * pr + 5], l2[8 * pr + 6], l2[8 * pr + 7]);
          st_global_256(ob + pos * 16, swap ? hb : ha, swap ? ha : hb);
          st_global_256(ob + LO_OFF + pos * 16, swap ? lb : la, swap ? la : lb);
        }
        if (MODE == EG_MSG0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[i] = pnext[i];
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&acc_empty[b])) : "memory");
    }
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// k_egemm_g -- a message linear of GVP 1 / 2 WITH ITS GATE LINEAR (scalar_to_vector_gates, gvp.py:125-127) in the same kernel.
//
// The separate GATE launch re-reads the 1 KB / edge activation its producer just wrote (18 launches, 3.9 ms, 24 GB per evaluation).
// Fusing it was blocked by capacity: the activation tile as a second shared-memory operand is 128 KB and TMEM is full.  With edges on
// M neither is needed: the epilogue thread that owns an edge row and has just computed 32 features of s' = SiLU(z) writes them back
// IN PLACE over the fp32 accumulator columns it read -- 16 columns of packed fp16 `hi`, 16 columns of packed `lo` (tcgen05.st) -- and
// the gate GEMM takes its A operand from TENSOR MEMORY:
//     G[128 edges][32 gates] += S'[128 edges][16 k] (TMEM) . Wg[32 gates][16 k]^T (shared, resident: 32 KB)
// (the pattern attention kernels use for P = softmax(S)).  The gate accumulator needs 32 columns of its own, and there are none
// (2 x 256 accumulator columns): the last 32-feature chunk is therefore parked in shared memory as an ordinary SW128 operand slab
// and ITS accumulator columns hold G.  A third issuer warp sends the 48 gate MMAs of tile t while the main issuer is already
// filling the other buffer with tile t + 1; four of the epilogue warps read G one chunk into the next tile (no bubble), apply bias and
// sigmoid and store the gate rows.
//
// MODE EG_MSG : s' also leaves as the operand images of the next linear (as k_egemm_e).
// MODE EG_MSGA: last message GVP -- s' is needed by nothing else: no image is written at all, the scalar messages are summed over
//               the in-edges of every destination node right here (rows are dst-major; per 32-row group, per segment one masked
//               32 x 32 transposing warp reduction) into M / partL / partF with 32-row pieces.
struct EggPlan {
  static constexpr int T = 128;
  static constexpr int NST = 3;
  static constexpr int XSTAGE = 32768;
  static constexpr int RING_BYTES = 4 * TC_UNIT;
  static constexpr int WG_BYTES = 8 * 4096;                 // gate weights: 4 k-slabs x (hi, lo) x [32 gates][64 k]
  static constexpr int PARK_BYTES = 32768;                  // the last chunk of s' as (half of) an operand slab: hi 16 KB | lo 16 KB
  static constexpr int NLW = 4, NEW = 8;
  static constexpr int THREADS = (3 + NLW + NEW) * 32;
  static constexpr int W_LOAD0 = 3, W_EPI0 = 3 + NLW;
  static constexpr int NROWBUF = 3;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_RING = NST * XSTAGE;
  static constexpr int OFF_WG = OFF_RING + RING_BYTES;
  static constexpr int OFF_PARK = OFF_WG + WG_BYTES;
  static constexpr int OFF_ROW = OFF_PARK + PARK_BYTES;     // NROWBUF x int row[T]
  static constexpr int OFF_BAR = OFF_ROW + NROWBUF * T * 4;
  static constexpr int NBAR = 8 + 2 * NST + 4 + NROWBUF + 2 + 2 + 1;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr size_t SMEM_BYTES = BYTES;
  static_assert(BYTES <= 232448, "227 KB of shared memory per CTA");
  static_assert(OFF_PARK % 1024 == 0 && OFF_WG % 1024 == 0, "operand tiles are 1024-byte aligned");
};

// registers -> TMEM: lane = this thread's row, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&f)[32]) {
  uint32_t lo[16], hi[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { lo[i] = __float_as_uint(f[i]); hi[i] = __float_as_uint(f[16 + i]); }
  tmem_st16(taddr, lo);
  tmem_st16(taddr + 16, hi);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::f16: A = M lanes x 8 columns of packed fp16 pairs (16 k values)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// IN_IMG = 0: the leading S k values arrive as fp32 rows `in_s` and are converted by the loaders (node rows, first GVP of a chain);
// OUT_F32 = 1: s' leaves as fp32 rows `out` instead of operand images (node rows, last GVP of a chain: k_node_mid reads them).
// SH_IMG = 1: the vector norms (the last k-slab) arrive as operand images too (`sh_img`, written by k_vecr_b): the loaders convert
// nothing, every k-slab of a tile is one bulk copy (knock-out timing: the conversion cost ~40 us per launch, profiles/r03d).
template <class D, int MODE, int IN_IMG = 1, int OUT_F32 = 0, int SH_IMG = 0>
__global__ void __launch_bounds__(EggPlan::THREADS, 1)
k_egemm_g(const ModelRT m, const BatchRT bt, const EgArgs a, const int n_tiles) {
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
  static_assert(!SH_IMG || IN_IMG, "norms as images: only on top of the image chain");
  constexpr int NIMG = IN_IMG ? S / 64 + (SH_IMG ? 1 : 0) : 0;
  constexpr int FIRST_CH = SH_IMG ? NCH : 2 * NIMG;
  static_assert(SH_IMG || FIRST_CH < NCH, "the loaders publish the row bookkeeping from their first converted chunk");
  static_assert(!(OUT_F32 || !IN_IMG) || MODE == EG_MSG, "fp32 rows in / out: node-row chains (no segment sum)");
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
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&x_full[i], PL::NLW); tc::mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], PL::NEW / 2);             // the four warps that read the gate accumulator
      tc::mbar_init(&a_ready[i], PL::NEW);
      tc::mbar_init(&gate_full[i], 1);
    }
    for (int i = 0; i < PL::NROWBUF; ++i) tc::mbar_init(&rows_full[i], PL::NLW);
    tc::mbar_init(wg_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- weight producer (+ the resident gate weights, once) --------------------------------------------------------------------------
    if (lane == 0) {
      if (n_my > 0) {
        tc::mbar_arrive_expect_tx(wg_full, PL::WG_BYTES);
        for (int u = 0; u < 8; ++u) tc::bulk_g2s(wg + u * 4096, reinterpret_cast<const uint8_t*>(a.g_units) + (size_t)u * 4096, 4096, wg_full);
      }
      // the four 16 KB units of a k-slab in stream order (f 0..127 hi, f 0..127 lo, f 128..255 hi, f 128..255 lo), each with its own
      // barrier pair: a unit is refilled as soon as ITS MMAs are done.  (A first version moved the two hi / lo units as one 32 KB
      // tile for N = 256 MMAs: only two refills in flight, tensor pipe 49 % busy, the epilogue warps waiting -- profiles/r02g.)
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
    // ---- gate MMA issuer: G = S' Wg^T once the eight epilogue warps have put s' of the tile in place (TMEM) / in the park (smem) ---------
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::idesc_f16(128, 32);
    const uint32_t wg_lo = tc::smem_u32(wg) >> 4, pk = tc::smem_u32(park) >> 4;
    if (n_my > 0) tc::mbar_wait(wg_full, 0);
    for (int it = 0; it < n_my; ++it) {
      const int b = it & 1;
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
      if (a.flags & EGF_NODE_ROWS) {
        ok = slot < a.EP;                                       // EP = number of nodes
      } else if (slot < a.EP) {
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
        if (!IN_IMG && j < S / 32) {
          if (ok) val = *(reinterpret_cast<const float4*>(a.in_s + (size_t)sl_ * S + j * 32) + ch);
        } else {
          const int k0 = (j - S / 32) * 32 + ch * 4;
          if (ok && k0 < SH_W) val = *(reinterpret_cast<const float4*>(a.in_sh + (size_t)sl_ * SH_W + k0));
        }
        buf[i] = val;
      }
    };
    float4 cur[8], nxt[8];
    float amax = 0.f;
    if (n_my > 0) { rowinfo(0); if (!SH_IMG) fetch(FIRST_CH, cur); }
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
              const uint8_t* srcp = (SH_IMG && s == NIMG - 1) ? reinterpret_cast<const uint8_t*>(a.sh_img) + (size_t)tile * PL::XSTAGE
                                                              : reinterpret_cast<const uint8_t*>(a.in_img) + ((size_t)tile * (S / 64) + s) * PL::XSTAGE;
              tc::bulk_g2s(xst + st * PL::XSTAGE, srcp, PL::XSTAGE, &x_full[st]);
              }
            } else {
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&x_full[st])) : "memory");
            }
          }
          __syncwarp();
          if (SH_IMG && s == NSLAB - 1 && it + 1 < n_my) rowinfo(it + 1);      // MSGA: the next tile's destination bookkeeping
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
        if (OUT_F32) {
          float* op = a.out + ((size_t)tile * PL::T + row) * S + c * 32;
#pragma unroll
          for (int i8 = 0; i8 < 4; ++i8)
            st_global_256(op + 8 * i8,
                          make_uint4(__float_as_uint(acc[8 * i8]), __float_as_uint(acc[8 * i8 + 1]), __float_as_uint(acc[8 * i8 + 2]), __float_as_uint(acc[8 * i8 + 3])),
                          make_uint4(__float_as_uint(acc[8 * i8 + 4]), __float_as_uint(acc[8 * i8 + 5]), __float_as_uint(acc[8 * i8 + 6]), __float_as_uint(acc[8 * i8 + 7])));
        } else
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
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&a_ready[b])) : "memory");
    }
    if (hf == 0 && n_my > 0) gate_epilogue(n_my - 1);
    if (!(omax < tc::ACT_LIMIT_H16) && a.status) atomicOr(a.status, 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace fm
