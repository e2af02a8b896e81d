// Register-resident vector-channel stages of the message GVPs (edge rows of the wide tensor-core pipeline).
//
// What the profiles said about vec_stages.cuh (profiles/r01s, r01y): k_vec_a/b/c are 26 % of a network evaluation and run at
// 25-52 % of the HBM peak -- every 16-row block goes global -> registers -> shared -> (GEMM) -> shared -> (cross, norms) -> shared
// -> (GEMM) -> shared -> global behind seven 64-thread barriers, and only the first of those phases has memory requests in flight.
//
// Here ONE WARP owns a unit of 16 rows (edges) end to end and nothing but the weights lives in shared memory:
//   * an m16 MMA tile is the 16 rows of ONE xyz plane (not 16 consecutive (row, plane) pairs), the three planes are three tiles of
//     the same warp: a thread then holds x, y and z of the same (row, column) -- norms and cross products are thread-local
//     (the cross product's two column groups meet through one shuffle) -- and
//   * the accumulator fragment of the first GEMM ([Vh | Vcp] = V [Wh | Wcp]) IS the A fragment of the second (Vu = Vh_ext Wu):
//     c0..c3 of n-tiles 2j, 2j+1 are a0..a3 of k-step j, so the chain never leaves the register file;
//   * the A operand of the first GEMM is loaded from global memory in fragment layout (a quad reads 32 contiguous bytes of a row),
//     the next unit's rows are prefetched HBM -> L2 with one bulk-prefetch instruction each while this one computes.
// No barriers, no activation tiles in shared memory, warps drift freely: memory requests of ~16 warps per SM overlap all the time.
//
// The global intermediates change with it: instead of VH (= [Vh | cross], 3 x 40 floats per row, consumed by the NEXT kernel's
// Vu GEMM) a stage stores VU = Vh_ext Wu (3 x 32), i.e. both GEMMs of a GVP sit in the kernel that produced Vh, the kernel after
// the gate linear starts with V' = gate * VU, and the last stage (k_vecr_c) has no GEMM at all.  Same arithmetic on the same
// values in the same order as vec_stages.cuh (k_vecr_b / k_vecr_c: bit-identical; k_vecr_a orders K as v_src | x_diff so that its
// fragment loads are aligned -- a different summation order inside the first MMA k-steps).
//
//   k_vecr_a : gather x / v of src;  Vh0 = [v_src | x_diff] [Wh | Wcp], cross, norms -> SH;  VU = Vh0_ext Wu      (message GVP 0)
//   k_vecr_b : V' = GT * VU (GVP g);  Vh = V' [Wh | Wcp], cross, norms -> SH;  VU = Vh_ext Wu                      (GVP g + 1)
//   k_vecr_c : V' = GT * VU (GVP 2), segment-sum over the in-edges of every destination -> M / partL / partF (vector columns)
#pragma once
#include "vec_stages.cuh"

namespace fm {

constexpr int VUW = 96;       // floats per row of VU: 3 planes x 32 vector channels
constexpr int UR = 16;        // rows per unit (= one m16 tile per plane)

struct VrUnit {
  int nvalid;                 // live rows (a prefix of the unit)
  int n, nb, le0;             // atoms / first node of the unit's molecule, local edge index of row 0
};
__device__ __forceinline__ VrUnit vr_unit(const BatchRT& bt, int unit) {
  VrUnit u;
  const int tile = unit >> 2, mol = __ldg(bt.etile_mol + tile);
  u.n = __ldg(bt.mol_n + mol);
  u.nb = __ldg(bt.mol_node + mol);
  u.le0 = (tile - __ldg(bt.mol_etile + mol)) * TM + (unit & 3) * UR;
  u.nvalid = max(0, min(UR, u.n * (u.n - 1) - u.le0));
  return u;
}

// Channel permutation of the fragment layouts.  In an m16n8k16 fragment a thread owns logical k (or n) positions 2t, 2t + 1 and
// 2t + 8, 2t + 9 of every group of 16: read straight from memory that is two 8-byte accesses per row and group, and a warp
// instruction touches 8 rows x 32 bytes -- the vector stages ran at 56-64 % LSU-wavefront utilisation moving 33 bytes per wavefront
// (profiles/r02_ncu_full_other.csv).  The contraction does not care which physical channel sits at which logical position as long
// as the other operand agrees, and the weights are re-laid out in shared memory once per CTA anyway: logical position
// 16 m + 8 h + 2 t + c  <->  physical channel 16 m + 4 t + 2 h + c.  A thread's four positions of a group are then FOUR CONSECUTIVE
// channels: one 16-byte access, a quad covers 64 contiguous bytes of the row, half the memory instructions.  Memory layouts stay
// in natural channel order (k_vecr_c and the scalar linears read them unchanged).
__device__ __forceinline__ int vr_logical_of(int pc) { return (pc & ~15) | (((pc >> 1) & 1) << 3) | (((pc >> 2) & 3) << 1) | (pc & 1); }
__device__ __forceinline__ int vr_physical_of(int lc) { return (lc & ~15) | (((lc >> 1) & 3) << 2) | (((lc >> 3) & 1) << 1) | (lc & 1); }

// load_resident_h16 with an optional row rotation: rot1 != 0 puts source row 0 LAST (smem row k <- source row k + 1 for k < Kreal - 1,
// smem row Kreal - 1 <- source row 0): message GVP 0 contracts over [v_src | x_diff] instead of the reference's [x_diff | v_src]
// permK > 0: logical k rows [0, permK) hold the source rows of the permuted physical channels (A operand read as 16-byte groups);
// permN > 0: logical columns [0, permN) hold the permuted physical output channels (C fragment stored as 16-byte groups)
__device__ __forceinline__ WH16 vr_load_w(float* dst, const float* __restrict__ src, int K, int Kreal, int np, int ld, int rot1, float* red,
                                          int permK = 0, int permN = 0) {
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
  const int kp2 = ((K + 7) & ~7) >> 1;
  uint32_t* hi = reinterpret_cast<uint32_t*>(dst);
  uint32_t* lo = hi + kp2 * ld;
  auto srow = [&](int k) {
    if (k < permK) k = vr_physical_of(k);
    return rot1 ? (k < Kreal - 1 ? k + 1 : (k == Kreal - 1 ? 0 : k)) : k;
  };
  for (int i = tid; i < kp2 * np; i += NT) {
    const int r = i / np, n = i - r * np, k = 2 * r;
    const int sn = n < permN ? vr_physical_of(n) : n;          // smem (logical) column n <- source column
    const float w0 = k < K ? __ldg(src + srow(k) * np + sn) * scale : 0.f, w1 = k + 1 < K ? __ldg(src + srow(k + 1) * np + sn) * scale : 0.f;
    uint32_t h2, l2;
    tc::split_h16x2(w0, w1, h2, l2);
    hi[r * ld + n] = h2;
    lo[r * ld + n] = l2;
  }
  return WH16{hi, lo, inv};
}

// D[p][nt] += A[p] W for the three planes p, one k-step of 16 (8): per accumulator the same three products in the same order as
// warp_gemm_h16x3 (a_lo b_hi, a_hi b_lo, a_hi b_hi).  ncu (profiles/r02b): with ~4 warps per scheduler the stages spent 2 issue
// slots per instruction waiting on fixed-latency dependencies -- the three HMMAs of one accumulator back to back -- so the products
// are issued round-robin over the NTL x 3 accumulators (dependent HMMAs are NTL * 3 instructions apart) and the asm statements are
// not volatile (pure functions of their operands: ptxas may interleave them with the splits and loads around them).
__device__ __forceinline__ void vr_mma16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void vr_mma8(float (&d)[4], const uint32_t (&a)[2], const uint32_t b) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(b));
}
template <int NTL>
__device__ __forceinline__ void vr_kstep16(float (&acc)[3][NTL][4], const uint32_t (&ah)[3][4], const uint32_t (&al)[3][4],
                                           const uint32_t* __restrict__ Wh, const uint32_t* __restrict__ Wl, int ldw, int k0, int g, int t) {
  uint32_t bh[NTL][2], bl[NTL][2];
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int w = ((k0 >> 1) + t + 4 * i) * ldw + 8 * nt + g;
      bh[nt][i] = Wh[w]; bl[nt][i] = Wl[w];
    }
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int p = 0; p < 3; ++p) vr_mma16(acc[p][nt], al[p], bh[nt]);
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int p = 0; p < 3; ++p) vr_mma16(acc[p][nt], ah[p], bl[nt]);
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int p = 0; p < 3; ++p) vr_mma16(acc[p][nt], ah[p], bh[nt]);
}
template <int NTL>
__device__ __forceinline__ void vr_kstep8(float (&acc)[3][NTL][4], const uint32_t (&ah)[3][2], const uint32_t (&al)[3][2],
                                          const uint32_t* __restrict__ Wh, const uint32_t* __restrict__ Wl, int ldw, int k0, int g, int t) {
  uint32_t bh[NTL], bl[NTL];
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt) {
    const int w = ((k0 >> 1) + t) * ldw + 8 * nt + g;
    bh[nt] = Wh[w]; bl[nt] = Wl[w];
  }
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int p = 0; p < 3; ++p) vr_mma8(acc[p][nt], al[p], bh[nt]);
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int p = 0; p < 3; ++p) vr_mma8(acc[p][nt], ah[p], bl[nt]);
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
    for (int p = 0; p < 3; ++p) vr_mma8(acc[p][nt], ah[p], bh[nt]);
}

// The tail every stage shares: acc1 = [Vh | cross | 0] (already unscaled, cross products in place, columns >= hc zero) of the warp's
// 16 rows x 3 planes -> norms to SH, Vu = Vh_ext Wu (K = hc padded to 40: k-steps 16, 16, 8) to VU.  NT1 >= 5 n-tiles.
template <int NT1>
__device__ __forceinline__ void vr_tail(float (&acc1)[3][NT1][4], const int hc, const WH16& wu, const size_t r0, const bool ok0, const bool ok1,
                                        float* __restrict__ VU, float* __restrict__ SH, float* __restrict__ SHI = nullptr) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // gvp.py:14-21 sqrt(clamp(x^2 + y^2 + z^2, 1e-8)) per hidden vector channel; dead rows / padding columns store 0
  uint32_t wh[5][2], wl[5][2];                               // SHI: packed (hi, hi) / (lo, lo) words of this lane's two columns per n-tile and row
#pragma unroll
  for (int nt = 0; nt < 5; ++nt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int col = 8 * nt + 2 * t;
      const bool okr = hh ? ok1 : ok0;
      float nn[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float vx = acc1[0][nt][2 * hh + c], vy = acc1[1][nt][2 * hh + c], vz = acc1[2][nt][2 * hh + c];
        const float q = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        nn[c] = (okr && col + c < hc) ? sqrt_pos(fmaxf(q, 1e-8f)) : 0.f;
      }
      if (SHI) tc::split_h16x2(nn[0] * tc::ACT_SCALE_H16, nn[1] * tc::ACT_SCALE_H16, wh[nt][hh], wl[nt][hh]);
      else *reinterpret_cast<float2*>(SH + (r0 + g + 8 * hh) * VHW + col) = make_float2(nn[0], nn[1]);
    }
  if (SHI) {
    // The norms as the LAST k-slab of the next scalar linear's operand images (egemm_p.cuh): per 128-row tile one 32 KB block
    // [hi 16 KB | lo 16 KB], row r = 128 bytes, 16-byte piece p (8 k values) at position p ^ (r % 8); columns >= 40 stay zero
    // (fm_batch_init).  A quad holds piece nt of a row as four 4-byte words: a 4 x 4 word transpose inside the quad (four shuffles)
    // gives lane t all 16 bytes of piece t, so pieces 0..3 leave as ONE 16-byte store per lane (full sectors); piece 4 as words.
    // (Stored word by word the stage was 29 % slower: 20 partial-sector stores per lane, profiles/r03e.)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const size_t slot = r0 + g + 8 * hh;
      const uint32_t r = (uint32_t)(slot & 127), r7 = r & 7u;
      uint8_t* rowb = reinterpret_cast<uint8_t*>(SHI) + (slot >> 7) * 32768 + r * 128;
#pragma unroll
      for (int hl = 0; hl < 2; ++hl) {
        uint32_t x0 = hl ? wl[0][hh] : wh[0][hh], x1 = hl ? wl[1][hh] : wh[1][hh], x2 = hl ? wl[2][hh] : wh[2][hh], x3 = hl ? wl[3][hh] : wh[3][hh];
        {   // 2 x 2 blocks: swap the off-diagonal words with the xor-1 partner
          const uint32_t s01 = (t & 1) ? x0 : x1, s23 = (t & 1) ? x2 : x3;
          const uint32_t q01 = __shfl_xor_sync(0xffffffffu, s01, 1), q23 = __shfl_xor_sync(0xffffffffu, s23, 1);
          if (t & 1) { x0 = q01; x2 = q23; } else { x1 = q01; x3 = q23; }
        }
        {   // swap the off-diagonal 2 x 2 blocks with the xor-2 partner
          const uint32_t sa = (t & 2) ? x0 : x2, sb = (t & 2) ? x1 : x3;
          const uint32_t qa = __shfl_xor_sync(0xffffffffu, sa, 2), qb = __shfl_xor_sync(0xffffffffu, sb, 2);
          if (t & 2) { x0 = qa; x1 = qb; } else { x2 = qa; x3 = qb; }
        }
        *reinterpret_cast<uint4*>(rowb + hl * 16384 + (((uint32_t)t ^ r7) << 4)) = make_uint4(x0, x1, x2, x3);
        // piece 4 (columns 32..39) shares its 32-byte sector with the all-zero piece 5: the quad writes the whole sector (8 bytes per
        // lane), so that no sector of the image is ever written partially
        {
          const uint32_t w4 = hl ? wl[4][hh] : wh[4][hh];
          const int q0 = (lane & ~3) + 2 * (t & 1);
          const uint32_t wa = __shfl_sync(0xffffffffu, w4, q0), wb = __shfl_sync(0xffffffffu, w4, q0 + 1);
          const uint32_t p4 = 4u ^ r7;                        // position of piece 4; its sector starts at position p4 & ~1
          const bool mine = (uint32_t)(t >> 1) == (p4 & 1u);
          *reinterpret_cast<uint2*>(rowb + hl * 16384 + ((p4 & ~1u) << 4) + 8 * t) = mine ? make_uint2(wa, wb) : make_uint2(0u, 0u);
        }
      }
    }
  }
  float acc2[3][4][4];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc2[p][nt][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint32_t ah[3][4], al[3][4];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      tc::split_h16x2(acc1[p][2 * j][0], acc1[p][2 * j][1], ah[p][0], al[p][0]);
      tc::split_h16x2(acc1[p][2 * j][2], acc1[p][2 * j][3], ah[p][1], al[p][1]);
      tc::split_h16x2(acc1[p][2 * j + 1][0], acc1[p][2 * j + 1][1], ah[p][2], al[p][2]);
      tc::split_h16x2(acc1[p][2 * j + 1][2], acc1[p][2 * j + 1][3], ah[p][3], al[p][3]);
    }
    vr_kstep16<4>(acc2, ah, al, wu.hi, wu.lo, WLD_U, 16 * j, g, t);
  }
  {
    uint32_t ah[3][2], al[3][2];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      tc::split_h16x2(acc1[p][4][0], acc1[p][4][1], ah[p][0], al[p][0]);
      tc::split_h16x2(acc1[p][4][2], acc1[p][4][3], ah[p][1], al[p][1]);
    }
    vr_kstep8<4>(acc2, ah, al, wu.hi, wu.lo, WLD_U, 32, g, t);
  }
  // Wu's columns sit in shared memory in the permuted order (vr_load_w, permN): n-tiles 2 m, 2 m + 1 hold the four consecutive
  // output channels 16 m + 4 t .. + 3 of this thread -- one 16-byte store per plane, group and row
  const float inv = wu.inv;
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int mg = 0; mg < 2; ++mg)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
        *reinterpret_cast<float4*>(VU + (r0 + g + 8 * hh) * VUW + p * 32 + 16 * mg + 4 * t) =
            make_float4(acc2[p][2 * mg][2 * hh] * inv, acc2[p][2 * mg][2 * hh + 1] * inv, acc2[p][2 * mg + 1][2 * hh] * inv,
                        acc2[p][2 * mg + 1][2 * hh + 1] * inv);
}

// cross product of two 3-vectors with the roundings of torch.linalg.cross as vec_stage1 does them
__device__ __forceinline__ void vr_cross(float ax, float ay, float az, float qx, float qy, float qz, float& rx, float& ry, float& rz) {
  rx = __fsub_rn(__fmul_rn(ay, qz), __fmul_rn(az, qy));
  ry = __fsub_rn(__fmul_rn(az, qx), __fmul_rn(ax, qz));
  rz = __fsub_rn(__fmul_rn(ax, qy), __fmul_rn(ay, qx));
}

constexpr int VR_W1_WORDS = 2 * 20 * WLD_HCP;      // [Wh | Wcp] hi + lo, K <= 40
constexpr int VR_W2_WORDS = 2 * 20 * WLD_U;        // Wu hi + lo, K <= 40

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vecr_b(const BatchRT bt, const float* __restrict__ whcp, const float* __restrict__ wu, const int n_units,
         float* __restrict__ VU, float* __restrict__ SH, const float* __restrict__ GT, float* __restrict__ SHI) {
  pdl_launch();
  pdl_wait();
  static_assert(D::V == 32 && D::CP == 4, "fragment mapping: 32 vector channels, 4 cross-product features");
  __shared__ __align__(16) float wsm[VR_W1_WORDS + VR_W2_WORDS];
  __shared__ float red[NWARP];
  const WH16 w1 = vr_load_w(wsm, whcp, D::V, D::V, 32 * D::CPT_HC, WLD_HCP, 0, red, /*permK=*/D::V, 0);
  const WH16 w2 = vr_load_w(wsm + VR_W1_WORDS, wu, pad4(D::V + D::CP), pad4(D::V + D::CP), 32, WLD_U, 0, red, 0, /*permN=*/32);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nw = gridDim.x * NWARP;
  for (int unit = blockIdx.x * NWARP + (threadIdx.x >> 5); unit < n_units; unit += nw) {
    const VrUnit u = vr_unit(bt, unit);
    const size_t r0 = (size_t)unit * UR;
    if (lane == 0 && unit + nw < n_units) {
      prefetch_l2(VU + (r0 + (size_t)nw * UR) * VUW, UR * VUW * 4);
      prefetch_l2(GT + (r0 + (size_t)nw * UR) * 32, UR * 32 * 4);
    }
    const bool ok0 = g < u.nvalid, ok1 = g + 8 < u.nvalid;
    // A fragments of V' = gate * VU straight from global memory.  Permuted channel order (vr_load_w, permK): the thread's logical
    // positions 2t, 2t + 1 (a0 / a1) and 2t + 8, 2t + 9 (a2 / a3) of k-step ks are the physical channels 16 ks + 4t .. + 3
    float4 gq[2][2], vq[3][2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const size_t row = r0 + g + hh * 8;
        const int col = 16 * ks + 4 * t;
        const bool ok = hh ? ok1 : ok0;
        gq[ks][hh] = ok ? *reinterpret_cast<const float4*>(GT + row * 32 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 3; ++p)
          vq[p][ks][hh] = ok ? *reinterpret_cast<const float4*>(VU + row * VUW + p * 32 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    float acc1[3][5][4];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int nt = 0; nt < 5; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc1[p][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ah[3][4], al[3][4];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float4 gg = gq[ks][hh], vw = vq[p][ks][hh];
          tc::split_h16x2(__fmul_rn(gg.x, vw.x), __fmul_rn(gg.y, vw.y), ah[p][hh], al[p][hh]);
          tc::split_h16x2(__fmul_rn(gg.z, vw.z), __fmul_rn(gg.w, vw.w), ah[p][2 + hh], al[p][2 + hh]);
        }
      vr_kstep16<5>(acc1, ah, al, w1.hi, w1.lo, WLD_HCP, 16 * ks, g, t);
    }
    {
      const float inv = w1.inv;
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int nt = 0; nt < 5; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc1[p][nt][i] *= inv;
    }
    // cross products: n-tile 4 holds Vcp -- quads' lanes t = 0, 1 the columns 32 + 2t, 33 + 2t (first factors), lanes t + 2 the
    // matching second factors 36 + 2t, 37 + 2t: one xor-2 exchange; the results replace the first factors, the rest becomes K padding
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float qx = __shfl_xor_sync(0xffffffffu, acc1[0][4][i], 2), qy = __shfl_xor_sync(0xffffffffu, acc1[1][4][i], 2),
                  qz = __shfl_xor_sync(0xffffffffu, acc1[2][4][i], 2);
      float rx, ry, rz;
      vr_cross(acc1[0][4][i], acc1[1][4][i], acc1[2][4][i], qx, qy, qz, rx, ry, rz);
      acc1[0][4][i] = t < 2 ? rx : 0.f;
      acc1[1][4][i] = t < 2 ? ry : 0.f;
      acc1[2][4][i] = t < 2 ? rz : 0.f;
    }
    vr_tail<5>(acc1, D::V + D::CP, w2, r0, ok0, ok1, VU, SH, SHI);
  }
}

template <class D>
__global__ void __launch_bounds__(NT, 2)
k_vecr_a(const ModelRT m, const BatchRT bt, const int layer, const int n_units, const float* __restrict__ x, const float* __restrict__ v,
         float* __restrict__ VU, float* __restrict__ SH) {
  pdl_launch();
  pdl_wait();
  static_assert(D::V == 32 && D::CP == 4 && D::VIN0 == 33 && D::H0 == 33, "fragment mapping: [v_src(32) | x_diff], 4 cross-product features");
  __shared__ __align__(16) float wsm[VR_W1_WORDS + VR_W2_WORDS];
  __shared__ float red[NWARP];
  __shared__ float xch[NWARP][UR * 3 * 8];                    // per warp: the 8 Vcp columns of its 16 rows x 3 planes
  const WH16 w1 = vr_load_w(wsm, m.c(layer, C_MSG0_WHCP), pad4(D::VIN0), D::VIN0, 32 * D::CPT_HC0, WLD_HCP, 1, red, /*permK=*/D::V, 0);
  const WH16 w2 = vr_load_w(wsm + VR_W1_WORDS, m.c(layer, C_MSG0_WU), pad4(D::H0 + D::CP), pad4(D::H0 + D::CP), 32, WLD_U, 0, red, 0, /*permN=*/32);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
  float* xc = xch[warp];
  const int nw = gridDim.x * NWARP;
  constexpr int H = D::H0, HC = D::H0 + D::CP;                // 33, 37
  for (int unit = blockIdx.x * NWARP + warp; unit < n_units; unit += nw) {
    const VrUnit u = vr_unit(bt, unit);
    const size_t r0 = (size_t)unit * UR;
    const bool okr[2] = {g < u.nvalid, g + 8 < u.nvalid};
    int src[2];
    float ud[2][3];                                            // unit vector x_diff of this thread's two rows
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      src[hh] = u.nb;
      ud[hh][0] = ud[hh][1] = ud[hh][2] = 0.f;
      if (okr[hh]) {
        int i, j;
        edge_src_dst(u.le0 + g + 8 * hh, u.n, i, j);
        src[hh] = u.nb + i;
        float dx, dy, dz;
        const float dist = pair_dist(x, src[hh], u.nb + j, dx, dy, dz);
        ud[hh][0] = __fdiv_rn(dx, dist); ud[hh][1] = __fdiv_rn(dy, dist); ud[hh][2] = __fdiv_rn(dz, dist);
      }
    }
    float4 vq[3][2][2];                                        // permuted channel order: 16 bytes = the thread's four positions of a k-step
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = 16 * ks + 4 * t;
#pragma unroll
        for (int p = 0; p < 3; ++p)
          vq[p][ks][hh] = okr[hh] ? __ldg(reinterpret_cast<const float4*>(v + (size_t)src[hh] * 3 * D::V + p * D::V + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    float acc1[3][6][4];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int nt = 0; nt < 6; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc1[p][nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ah[3][4], al[3][4];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tc::split_h16x2(vq[p][ks][hh].x, vq[p][ks][hh].y, ah[p][hh], al[p][hh]);
          tc::split_h16x2(vq[p][ks][hh].z, vq[p][ks][hh].w, ah[p][2 + hh], al[p][2 + hh]);
        }
      vr_kstep16<6>(acc1, ah, al, w1.hi, w1.lo, WLD_HCP, 16 * ks, g, t);
    }
    {
      // k = 32: x_diff (lane t = 0 of every quad), 33..39: K padding
      uint32_t ah[3][2], al[3][2];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) tc::split_h16x2(t == 0 ? ud[hh][p] : 0.f, 0.f, ah[p][hh], al[p][hh]);
      vr_kstep8<6>(acc1, ah, al, w1.hi, w1.lo, WLD_HCP, 32, g, t);
    }
    {
      const float inv = w1.inv;
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int nt = 0; nt < 6; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc1[p][nt][i] *= inv;
    }
    // cross products: Vcp = columns [33, 41) (n-tiles 4 and 5) -> per-warp exchange buffer -> the owners of columns 33..36 compute
    __syncwarp();
#pragma unroll
    for (int nt = 4; nt < 6; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = 8 * nt + 2 * t + (i & 1), row = g + 8 * (i >> 1);
        if (col >= H && col < H + 2 * D::CP) {
#pragma unroll
          for (int p = 0; p < 3; ++p) xc[(row * 3 + p) * 8 + (col - H)] = acc1[p][nt][i];
        }
      }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = 32 + 2 * t + (i & 1), row = g + 8 * (i >> 1);
      if (col >= H && col < HC) {
        const float* q = xc + (row * 3) * 8 + (col - H) + D::CP;
        float rx, ry, rz;
        vr_cross(acc1[0][4][i], acc1[1][4][i], acc1[2][4][i], q[0], q[8], q[16], rx, ry, rz);
        acc1[0][4][i] = rx; acc1[1][4][i] = ry; acc1[2][4][i] = rz;
      } else if (col >= HC) {
        acc1[0][4][i] = 0.f; acc1[1][4][i] = 0.f; acc1[2][4][i] = 0.f;
      }
    }
    vr_tail<6>(acc1, HC, w2, r0, okr[0], okr[1], VU, SH);
  }
}

// V' = GT * VU of the last message GVP and its sum over the in-edges of every destination node (gvp.py:491-492, vector half):
// one warp per 64-slot tile, lane = vector channel, the three planes in three accumulators; rows are dst-major, a segment that
// is cut by the tile goes to partL (its head) / partF (a continuation) exactly as in k_vec_c / k_conv_edge.
template <class D>
__global__ void __launch_bounds__(NT)
k_vecr_c(const BatchRT bt, const int piece_rows /* 64 or 32: rows per aggregation piece (k_node_pre's agg_rows) */,
         const float* __restrict__ VU, const float* __restrict__ GT, float* __restrict__ M, float* __restrict__ partF,
         float* __restrict__ partL) {
  pdl_launch();
  pdl_wait();
  static_assert(D::V == 32, "lane = vector channel");
  const int lane = threadIdx.x & 31, piece = blockIdx.x * NWARP + (threadIdx.x >> 5);
  const int ppt = TM / piece_rows, tile = piece / ppt;          // pieces per 64-slot tile
  if (tile >= bt.n_edge_tiles) return;
  const int mol = __ldg(bt.etile_mol + tile), n = __ldg(bt.mol_n + mol), nb = __ldg(bt.mol_node + mol);
  const int le0 = (tile - __ldg(bt.mol_etile + mol)) * TM + (piece - tile * ppt) * piece_rows;
  const int nvalid = min(piece_rows, n * (n - 1) - le0), deg = n - 1;
  if (nvalid <= 0) return;
  const size_t erow0 = (size_t)piece * piece_rows;
  int j = le0 / deg, rem = le0 - j * deg;                      // destination (local) and position inside its in-edge segment of row 0
  bool head = rem == 0;                                        // the running segment started with the node's first in-edge
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  constexpr int CH = 8;
  for (int rb = 0; rb < nvalid; rb += CH) {
    float gbuf[CH], b0[CH], b1[CH], b2[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const size_t r = erow0 + min(rb + k, nvalid - 1);
      gbuf[k] = __ldg(GT + r * 32 + lane);
      b0[k] = __ldg(VU + r * VUW + lane); b1[k] = __ldg(VU + r * VUW + 32 + lane); b2[k] = __ldg(VU + r * VUW + 64 + lane);
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int row = rb + k;
      if (row < nvalid) {
        a0 = __fadd_rn(a0, __fmul_rn(gbuf[k], b0[k]));
        a1 = __fadd_rn(a1, __fmul_rn(gbuf[k], b1[k]));
        a2 = __fadd_rn(a2, __fmul_rn(gbuf[k], b2[k]));
        const bool tail = rem == deg - 1;
        if (tail || row == nvalid - 1) {
          float* dst = (head && tail) ? M + (size_t)(nb + j) * D::MW : (head ? partL + (size_t)piece * D::MW : partF + (size_t)piece * D::MW);
          dst[D::S + lane] = a0; dst[D::S + 32 + lane] = a1; dst[D::S + 64 + lane] = a2;
          a0 = a1 = a2 = 0.f;
          head = true;                                         // the next segment (if any) starts at its node's first in-edge
        }
        if (tail) { rem = 0; ++j; } else ++rem;
      }
    }
  }
}

}  // namespace fm
