// Register-resident upper-edge kernels: the two 128-wide MLPs that run once per network evaluation on the UPPER edges (i < j)
//   k_edge_init_r : self-conditioning residual of the edge features (self_conditioning.py:59-82) -- [e_1_pred | rbf(d_1) - rbf(d_t)]
//                   -> 128 -> SiLU -> 128 -> SiLU, + embedded edge token, mirrored to both directed edges; writes the fp32 rows AND
//                   the fp16 (hi, lo) operand images the message / edge-update linears read (replaces k_edge_init + k_ef_image)
//   k_edge_head_r : bond-order head (vector_field.py:342-344): ef[i->j] + ef[j->i] -> 128 -> SiLU -> n_bond_types, softmax
//
// What the profile said (profiles/r02h): k_edge_init 0.98 ms + k_ef_image 0.26 ms + k_edge_head 1.07 ms = 8 % of an evaluation for
// 0.6 GB of traffic: fp32 FFMA tile GEMMs at ~22 TFLOP/s behind five CTA-wide barriers per tile.  Same recipe as the vector stages
// (vec_reg.cuh): ONE WARP owns 16 rows end to end, the GEMMs are error-compensated fp16x3 mma.sync.m16n8k16 (a_lo b_hi, a_hi b_lo,
// a_hi b_hi, fp32 accumulate) with the weights split once per CTA into shared memory, the accumulator fragment of the first linear
// IS the A fragment of the second (c0..c3 of n-tiles 2j, 2j + 1 = a0..a3 of k-step j), nothing else touches shared memory, no
// barriers after the prologue.
#pragma once
#include "vec_reg.cuh"

namespace fm {

constexpr int ER_LD = 136;                        // words per k-pair row of a 128-column weight image (% 32 == 8: conflict-free B fragments)
constexpr int ER_LD8 = 40;                        // ... of the head's last linear (n_bond_types <= 8 columns used, 32 stored)

// one k-step of 16 for NTL n-tiles starting at column n0: per accumulator the products a_lo b_hi, a_hi b_lo, a_hi b_hi (the order of
// warp_gemm_h16x3), issued round-robin over groups of four accumulators so that dependent HMMAs are four instructions apart
template <int NTL>
__device__ __forceinline__ void er_kstep16(float (&acc)[NTL][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t* __restrict__ Wh, const uint32_t* __restrict__ Wl, const int ldw, const int k0,
                                           const int n0, const int g, const int t) {
  constexpr int G = NTL < 4 ? NTL : 4;
#pragma unroll
  for (int nb = 0; nb < NTL; nb += G) {
    uint32_t bh[G][2], bl[G][2];
#pragma unroll
    for (int q = 0; q < G; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int w = ((k0 >> 1) + t + 4 * i) * ldw + n0 + 8 * (nb + q) + g;
        bh[q][i] = Wh[w]; bl[q][i] = Wl[w];
      }
#pragma unroll
    for (int q = 0; q < G; ++q) vr_mma16(acc[nb + q], al, bh[q]);
#pragma unroll
    for (int q = 0; q < G; ++q) vr_mma16(acc[nb + q], ah, bl[q]);
#pragma unroll
    for (int q = 0; q < G; ++q) vr_mma16(acc[nb + q], ah, bh[q]);
  }
}

// load_resident_h16 for a row range of a larger matrix: `K` rows are read, the image has Kpad / 2 k-pair rows (zeros beyond K)
__device__ __forceinline__ WH16 er_load_w(float* dst, const float* __restrict__ src, int K, int Kpad, int np, int ld, float* red) {
  const int tid = threadIdx.x;
  float mx = 0.f;
  for (int i = tid; i < K * np; i += NT) mx = fmaxf(mx, fabsf(__ldg(src + i)));
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = 0.f;
#pragma unroll
  for (int w = 0; w < NWARP; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const int e = mx >= 1e-30f ? (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127 : 13;
  const float scale = __uint_as_float((uint32_t)(127 + 13 - e) << 23), inv = __uint_as_float((uint32_t)(127 - 13 + e) << 23);
  const int kp2 = Kpad >> 1;
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

struct ErUnit {
  int n, ucount, lu0, ub;       // atoms, upper edges of the unit's molecule, local upper index of row 0, first compact upper edge
  long long ebase;              // first directed-edge slot of the molecule
  int nb;                       // first node
};
__device__ __forceinline__ ErUnit er_unit(const BatchRT& bt, int unit) {
  ErUnit u;
  const int tile = unit >> 2, mol = __ldg(bt.utile_mol + tile);
  u.n = __ldg(bt.mol_n + mol);
  u.nb = __ldg(bt.mol_node + mol);
  u.ucount = u.n * (u.n - 1) / 2;
  u.lu0 = (tile - __ldg(bt.mol_utile + mol)) * TM + (unit & 3) * UR;
  u.ub = __ldg(bt.mol_u + mol);
  u.ebase = (long long)__ldg(bt.mol_etile + mol) * TM;
  return u;
}

template <class D>
struct EdgeRegSmem {
  static constexpr int W128 = 2 * (D::F / 2) * ER_LD;               // words: a [128 k][128 n] matrix as (hi | lo) k-pair rows
  static constexpr int W_IN = 2 * 24 * ER_LD;                       // init, first linear: K <= 48
  static constexpr int W_OUT8 = 2 * (D::F / 2) * ER_LD8;            // head, last linear
  static constexpr size_t HEAD_BYTES = (size_t)(W128 + W_OUT8 + D::F + 32 + NWARP) * 4;
  static constexpr size_t INIT_BYTES = (size_t)(W_IN + W128 + 2 * 8 * D::F + D::F + NWARP) * 4;
};

// ------------------------------------------------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_edge_head_r(const ModelRT m, const BatchRT bt, const int n_units, const float* __restrict__ ef, float* __restrict__ pe) {
  pdl_launch();
  pdl_wait();
  static_assert(D::F == 128, "16 n-tiles / 8 k-steps");
  constexpr int F = D::F;
  extern __shared__ __align__(16) float er_smem[];
  float* w1s = er_smem;
  float* w2s = w1s + EdgeRegSmem<D>::W128;
  float* b1s = w2s + EdgeRegSmem<D>::W_OUT8;
  float* b2s = b1s + F;
  float* red = b2s + 32;
  const WH16 w1 = vr_load_w(w1s, m.g(G_EHEAD0_W), F, F, F, ER_LD, 0, red, /*permK=*/F, 0);     // permuted k order: 16-byte A loads (vec_reg.cuh)
  const WH16 w2 = load_resident_h16(w2s, m.g(G_EHEAD2_W), F, 32, ER_LD8, red);
  for (int i = threadIdx.x; i < F; i += NT) b1s[i] = m.g(G_EHEAD0_B)[i];
  if (threadIdx.x < 32) b2s[threadIdx.x] = m.g(G_EHEAD2_B)[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
  const int nw = gridDim.x * NWARP, EB = m.EB;
  for (int unit = blockIdx.x * NWARP + warp; unit < n_units; unit += nw) {
    const ErUnit u = er_unit(bt, unit);
    bool okr[2];
    const float *r0p[2], *r1p[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int lu = u.lu0 + g + 8 * hh;
      okr[hh] = lu < u.ucount;
      int i = 0, j = 1;
      if (okr[hh]) upper_ij(lu, u.n, i, j);
      r0p[hh] = ef + (size_t)(u.ebase + edge_pos(i, j, u.n)) * F;
      r1p[hh] = ef + (size_t)(u.ebase + edge_pos(j, i, u.n)) * F;
    }
    float acc1[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc1[nt][i] = 0.f;
#pragma unroll 2
    for (int ks = 0; ks < F / 16; ++ks) {
      // A fragment of ef[i->j] + ef[j->i] straight from global memory; permuted k order: the thread's positions 2t, 2t + 1 (a0 / a1)
      // and 2t + 8, 2t + 9 (a2 / a3) of this k-step are the four consecutive features 16 ks + 4t .. + 3
      uint32_t ah[4], al[4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = 16 * ks + 4 * t;
        float4 sm4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (okr[hh]) {
          const float4 a = *reinterpret_cast<const float4*>(r0p[hh] + col), b = *reinterpret_cast<const float4*>(r1p[hh] + col);
          sm4 = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
        }
        tc::split_h16x2(sm4.x, sm4.y, ah[hh], al[hh]);
        tc::split_h16x2(sm4.z, sm4.w, ah[2 + hh], al[2 + hh]);
      }
      er_kstep16<16>(acc1, ah, al, w1.hi, w1.lo, ER_LD, 16 * ks, 0, g, t);
    }
    // SiLU(. + b) on the accumulator fragment, which is the A fragment of the last linear
    float acc2[1][4] = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int j = 0; j < F / 16; ++j) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int nt = 2 * j + h2;
        const float2 bb = *reinterpret_cast<const float2*>(b1s + 8 * nt + 2 * t);
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float z = acc1[nt][i] * w1.inv + ((i & 1) ? bb.y : bb.x); o[i] = z * sigmoid_fast(z); }
        tc::split_h16x2(o[0], o[1], ah[2 * h2], al[2 * h2]);
        tc::split_h16x2(o[2], o[3], ah[2 * h2 + 1], al[2 * h2 + 1]);
      }
      er_kstep16<1>(acc2, ah, al, w2.hi, w2.lo, ER_LD8, 16 * j, 0, g, t);
    }
    // softmax over the n_bond_types logits of a row: they sit in the row's quad, columns 2t, 2t + 1
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int c0 = 2 * t;
      const float l0 = acc2[0][2 * hh] * w2.inv + b2s[c0], l1 = acc2[0][2 * hh + 1] * w2.inv + b2s[c0 + 1];
      const bool in0 = c0 < EB, in1 = c0 + 1 < EB;
      float mx = fmaxf(in0 ? l0 : -INFINITY, in1 ? l1 : -INFINITY);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float e0 = in0 ? expf(l0 - mx) : 0.f, e1 = in1 ? expf(l1 - mx) : 0.f;
      float sum = e0 + e1;
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (okr[hh]) {
        float* o = pe + (size_t)(u.ub + u.lu0 + g + 8 * hh) * EB;
        if (in0) o[c0] = __fdiv_rn(e0, sum);
        if (in1) o[c0 + 1] = __fdiv_rn(e1, sum);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// has_prev only (every evaluation but the self-conditioning pre-pass of the first step, which keeps k_edge_init + k_ef_image)
template <class D>
__global__ void __launch_bounds__(NT, 2)
k_edge_init_r(const ModelRT m, const BatchRT bt, const int n_units, const float* __restrict__ x_t, const uint8_t* __restrict__ e_t,
              const PredPtr prev, float* __restrict__ ef, float* __restrict__ img) {
  pdl_launch();
  pdl_wait();
  static_assert(D::F == 128 && D::R == 32, "16 n-tiles; 32 radial basis functions");
  constexpr int F = D::F;
  extern __shared__ __align__(16) float er_smem[];
  float* w1s = er_smem;
  float* w2s = w1s + EdgeRegSmem<D>::W_IN;
  float* tabs = w2s + EdgeRegSmem<D>::W128;          // [0, 8 F): embedded token rows; [8 F, 16 F): T1 = b1 + table . W1[0:F]  (k_edge_table)
  float* b2s = tabs + 2 * 8 * F;
  float* red = b2s + F;
  const int EB = m.EB;
  // rows [F, F + EB + R) of the first linear multiply the per-edge operand [e_1_pred | rbf differences]; image padded to 48 k values
  const WH16 w1 = er_load_w(w1s, m.g(G_SCE0_W) + (size_t)F * F, EB + D::R, 48, F, ER_LD, red);
  // second linear: output columns in the permuted order of vec_reg.cuh (a thread's values of n-tiles 2 m, 2 m + 1 = four consecutive channels)
  const WH16 w2 = vr_load_w(w2s, m.g(G_SCE2_W), F, F, F, ER_LD, 0, red, 0, /*permN=*/F);
  for (int i = threadIdx.x; i < 2 * 8 * F; i += NT) {
    const int half = i / (8 * F), r = (i - half * 8 * F) / F, c = i % F;
    tabs[i] = r <= EB ? m.eemb_table[(size_t)(half * (EB + 1) + r) * F + c] : 0.f;
  }
  for (int i = threadIdx.x; i < F; i += NT) b2s[i] = m.g(G_SCE2_B)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
  const int nw = gridDim.x * NWARP;
  const float inv_sigma = (float)D::R / m.rbf_dmax;
  const float* mu = m.g(G_RBF_MU);
  for (int unit = blockIdx.x * NWARP + warp; unit < n_units; unit += nw) {
    const ErUnit u = er_unit(bt, unit);
    bool okr[2];
    int tok[2];
    long long p0[2], p1[2];
    float dt[2], d1[2];
    const float* pe[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int lu = u.lu0 + g + 8 * hh;
      okr[hh] = lu < u.ucount;
      int i = 0, j = 1;
      tok[hh] = 0; dt[hh] = 0.f; d1[hh] = 0.f;
      pe[hh] = prev.e;
      if (okr[hh]) {
        upper_ij(lu, u.n, i, j);
        tok[hh] = e_t[u.ub + lu];
        float dx, dy, dz;
        dt[hh] = pair_dist(x_t, u.nb + i, u.nb + j, dx, dy, dz);          // self_conditioning.py:88-103 (edge_distances)
        d1[hh] = pair_dist(prev.x, u.nb + i, u.nb + j, dx, dy, dz);
        pe[hh] = prev.e + (size_t)(u.ub + lu) * EB;
      }
      p0[hh] = u.ebase + edge_pos(i, j, u.n);
      p1[hh] = u.ebase + edge_pos(j, i, u.n);
    }
    // operand element k of a row: k < EB: e_1_pred[k];  EB <= k < EB + R: rbf(d_1)[k - EB] - rbf(d_t)[k - EB];  else 0
    auto elem = [&](int hh, int k) -> float {
      if (!okr[hh]) return 0.f;
      if (k < EB) return pe[hh][k];
      if (k < EB + D::R) { const float c = __ldg(mu + (k - EB)); return __fsub_rn(rbf_fast(d1[hh], c, inv_sigma), rbf_fast(dt[hh], c, inv_sigma)); }
      return 0.f;
    };
    float acc1[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc1[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int hh = i & 1, k = 16 * ks + 2 * t + (i >> 1) * 8;
        tc::split_h16x2(elem(hh, k), elem(hh, k + 1), ah[i], al[i]);
      }
      er_kstep16<16>(acc1, ah, al, w1.hi, w1.lo, ER_LD, 16 * ks, 0, g, t);
    }
    // h = SiLU(. + T1[tok]) -> packed A fragments of the second linear (all 8 k-steps; the accumulators are dead after this)
    uint32_t fh[8][4], fl[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int nt = 2 * j + h2, col = 8 * nt + 2 * t;
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float z = acc1[nt][i] * w1.inv + tabs[8 * F + tok[i >> 1] * F + col + (i & 1)]; o[i] = z * sigmoid_fast(z); }
        tc::split_h16x2(o[0], o[1], fh[j][2 * h2], fl[j][2 * h2]);
        tc::split_h16x2(o[2], o[3], fh[j][2 * h2 + 1], fl[j][2 * h2 + 1]);
      }
    float* rowp[2][2];
    uint8_t* imgp[2][2];
    int r7[2][2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const long long slot = dir ? p1[hh] : p0[hh];
        rowp[hh][dir] = ef + (size_t)slot * F;
        imgp[hh][dir] = reinterpret_cast<uint8_t*>(img) + (size_t)(slot >> 7) * (F / 64) * 32768 + (size_t)(slot & 127) * 128;
        r7[hh][dir] = (int)(slot & 7);
      }
    // second linear in two halves of 64 output features; out = table[tok] + SiLU(. + b2), mirrored to both directed edges
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      float acc2[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc2[nt][i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) er_kstep16<8>(acc2, fh[j], fl[j], w2.hi, w2.lo, ER_LD, 16 * j, 64 * half, g, t);
#pragma unroll
      for (int mm = 0; mm < 4; ++mm) {
        const int col = 64 * half + 16 * mm + 4 * t;              // this thread's four consecutive output channels of n-tiles 2 mm, 2 mm + 1
        const float4 bb = *reinterpret_cast<const float4*>(b2s + col);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (!okr[hh]) continue;
          const float4 tb = *reinterpret_cast<const float4*>(tabs + tok[hh] * F + col);
          const float z0 = acc2[2 * mm][2 * hh] * w2.inv + bb.x, z1 = acc2[2 * mm][2 * hh + 1] * w2.inv + bb.y;
          const float z2 = acc2[2 * mm + 1][2 * hh] * w2.inv + bb.z, z3 = acc2[2 * mm + 1][2 * hh + 1] * w2.inv + bb.w;
          const float o0 = __fadd_rn(tb.x, z0 * sigmoid_fast(z0)), o1 = __fadd_rn(tb.y, z1 * sigmoid_fast(z1));
          const float o2 = __fadd_rn(tb.z, z2 * sigmoid_fast(z2)), o3 = __fadd_rn(tb.w, z3 * sigmoid_fast(z3));
          uint32_t ha, la, hb, lb;
          tc::split_h16x2(o0, o1, ha, la);
          tc::split_h16x2(o2, o3, hb, lb);
          const int piece = 2 * mm + (t >> 1);                     // 16-byte piece (8 k values) of the row inside k-slab `half`
#pragma unroll
          for (int dir = 0; dir < 2; ++dir) {
            *reinterpret_cast<float4*>(rowp[hh][dir] + col) = make_float4(o0, o1, o2, o3);
            uint8_t* ib = imgp[hh][dir] + half * 32768 + ((piece ^ r7[hh][dir]) << 4) + 8 * (t & 1);
            *reinterpret_cast<uint2*>(ib) = make_uint2(ha, hb);
            *reinterpret_cast<uint2*>(ib + 16384) = make_uint2(la, lb);
          }
        }
      }
    }
  }
}

}  // namespace fm
