// One Geometric Vector Perceptron (flowmol/models/gvp.py:90-133, vector gating) applied to a 64-row tile that lives
// in shared memory.  Rows are edges (message GVPs) or nodes (update / position GVPs).
//
//   inputs : scalars  Xs[row][0 : s_in)                      vectors  Va[plane][row][0 : v_in)   (plane = x,y,z)
//   outputs: scalars  Xs[row][0 : s_out)   (in place)        vectors  Va[plane][row][0 : v_out)  (in place)
//   scratch: Vb[plane][row][0 : h + 2cp), G[row][0 : 32), wstage (weight stream)
//
// Steps (all GEMMs through tile_gemm):
//   1. [Vh | Vcp] = V x [Wh | Wcp]                     (one GEMM, 3 planes)
//   2. cross products  cp_j = Vcp[j] x Vcp[cp + j]  -> appended after Vh;   sh = ||Vh_ext||  -> Xs[:, s_in : s_in+h+cp)
//   3. s' = SiLU( [s | sh] x W  + pre(row, col) )      `pre` = bias, or the per-node pre-activations gathered per edge
//   4. gate = s' x Wg + bg   (sigmoid unless identity)
//   5. V' = gate * (Vh_ext x Wu)
#pragma once
#include "tile_gemm.cuh"
#include "model.cuh"

namespace fm {

struct GvpShape {
  int v_in, h, cp, v_out, s_in, s_out;
  bool sigmoid_gate;
};

struct BiasPre {
  const float* b;
  __device__ __forceinline__ float operator()(int /*row*/, int col) const { return b[col]; }
};

template <int CPT_MAIN, int CPT_HC, int XLD, int LDVA, int LDVB, class Pre>
__device__ __forceinline__ void gvp_tile(float* __restrict__ Xs, float* __restrict__ Va, float* __restrict__ Vb,
                                         float* __restrict__ G, float* __restrict__ wstage, const GvpShape sh,
                                         const GvpPtr w, const Pre pre) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hc = sh.h + sh.cp;           // hidden vectors after appending the cross products
  // ---- 1. [Vh | Vcp] ------------------------------------------------------------------------------------------
  {
    float acc[3][RPW][CPT_HC];
    tile_gemm<3, CPT_HC>(Va, LDVA, TM * LDVA, pad4(sh.v_in), w.whcp, wstage, acc);
    const int ncol = sh.h + 2 * sh.cp;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int c = 0; c < CPT_HC; ++c) {
          const int col = ColMap<CPT_HC>::col(lane, c);
          if (col < ncol) Vb[(p * TM + warp * RPW + r) * LDVB + col] = acc[p][r][c];
        }
  }
  __syncthreads();
  // ---- 2. cross products (torch.linalg.cross) and norms --------------------------------------------------------------
  for (int idx = tid; idx < TM * sh.cp; idx += NT) {
    const int row = idx / sh.cp, j = idx - row * sh.cp;
    float* bx = Vb + (0 * TM + row) * LDVB;
    float* by = Vb + (1 * TM + row) * LDVB;
    float* bz = Vb + (2 * TM + row) * LDVB;
    const int ca = sh.h + j, cb = sh.h + sh.cp + j;
    const float ax = bx[ca], ay = by[ca], az = bz[ca], qx = bx[cb], qy = by[cb], qz = bz[cb];
    bx[ca] = __fsub_rn(__fmul_rn(ay, qz), __fmul_rn(az, qy));
    by[ca] = __fsub_rn(__fmul_rn(az, qx), __fmul_rn(ax, qz));
    bz[ca] = __fsub_rn(__fmul_rn(ax, qy), __fmul_rn(ay, qx));
  }
  __syncthreads();
  {
    const int hcp = pad4(hc);
    for (int idx = tid; idx < TM * hcp; idx += NT) {
      const int row = idx / hcp, c = idx - row * hcp;
      if (c < hc) {
        const float x = Vb[(0 * TM + row) * LDVB + c], y = Vb[(1 * TM + row) * LDVB + c], z = Vb[(2 * TM + row) * LDVB + c];
        Xs[row * XLD + sh.s_in + c] = norm_no_nan3(x, y, z);
      } else {                       // zero the K padding of both operands (stale Vcp columns / stale activations)
        Xs[row * XLD + sh.s_in + c] = 0.f;
        Vb[(0 * TM + row) * LDVB + c] = 0.f;
        Vb[(1 * TM + row) * LDVB + c] = 0.f;
        Vb[(2 * TM + row) * LDVB + c] = 0.f;
      }
    }
    // s_in is a multiple of 4 for every GVP on the path except none; if it were not, pad4(s_in + hc) handles the tail
    const int ktot = sh.s_in + hc, kpad = pad4(ktot);
    for (int idx = tid; idx < TM * 4; idx += NT) {
      const int row = idx >> 2, c = sh.s_in + hcp + (idx & 3);
      if (c < kpad) Xs[row * XLD + c] = 0.f;
    }
  }
  // ---- 3. scalar path ---------------------------------------------------------------------------------------------
  {
    float acc[1][RPW][CPT_MAIN];
    tile_gemm<1, CPT_MAIN>(Xs, XLD, 0, pad4(sh.s_in + hc), w.w, wstage, acc);   // ends with __syncthreads()
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = warp * RPW + r;
#pragma unroll
      for (int c = 0; c < CPT_MAIN; ++c) {
        const int col = ColMap<CPT_MAIN>::col(lane, c);
        if (col < sh.s_out) Xs[row * XLD + col] = silu_f(acc[0][r][c] + pre(row, col));
      }
    }
    // zero K padding for the gate GEMM (s_out is a multiple of 4 in all shipped configs; keep it general)
    for (int idx = tid; idx < TM * 4; idx += NT) {
      const int row = idx >> 2, c = sh.s_out + (idx & 3);
      if (c < pad4(sh.s_out)) Xs[row * XLD + c] = 0.f;
    }
  }
  // ---- 4. gates ------------------------------------------------------------------------------------------------------
  {
    float acc[1][RPW][1];
    tile_gemm<1, 1>(Xs, XLD, 0, pad4(sh.s_out), w.wg, wstage, acc);
    if (lane < sh.v_out) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float gte = acc[0][r][0] + w.bg[lane];
        G[(warp * RPW + r) * 32 + lane] = sh.sigmoid_gate ? sigmoid_f(gte) : gte;
      }
    }
  }
  // ---- 5. vector path ----------------------------------------------------------------------------------------------
  {
    float acc[3][RPW][1];
    tile_gemm<3, 1>(Vb, LDVB, TM * LDVB, pad4(hc), w.wu, wstage, acc);      // entry barrier makes G visible
    const int vpad = pad4(sh.v_out);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int row = warp * RPW + r;
        if (lane < sh.v_out) Va[(p * TM + row) * LDVA + lane] = __fmul_rn(G[row * 32 + lane], acc[p][r][0]);
        else if (lane < vpad) Va[(p * TM + row) * LDVA + lane] = 0.f;
      }
  }
  __syncthreads();
}

}  // namespace fm
