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

}  // namespace fm
