// Compile-time model dimensions, runtime model / batch descriptors shared by all kernels.
#pragma once
#include "common.cuh"
#include "weight_ids.h"

namespace fm {

constexpr int AMAX = 16;   // max atom types (incl. fake) supported by the shared-memory plan

template <int S_, int V_, int F_, int R_, int CP_, int SD_, int VD_, int TOK_, int TD_>
struct Dims {
  static constexpr int S = S_, V = V_, F = F_, R = R_, CP = CP_, SD = SD_, VD = VD_, TOK = TOK_, TD = TD_;
  static constexpr int VIN0 = 1 + V + VD;                  // x_diff | v_src | v_dst_msg      (gvp.py:526-529)
  static constexpr int H0 = VIN0 > V ? VIN0 : V;           // hidden vectors of message GVP 0
  static constexpr int KE0 = R + F + H0 + CP;              // per-edge K of message GVP 0 (s_src / s_dst rows folded per node)
  static constexpr int K1 = S + V + CP;                    // K of every other S->S GVP
  static constexpr int MW = S + 3 * V;                     // aggregated message width (scalars | x-plane | y-plane | z-plane)
  static constexpr int CPT_S = S / 32, CPT_F = F / 32;
  static constexpr int CPT_HC0 = (H0 + 2 * CP + 31) / 32, CPT_HC = (V + 2 * CP + 31) / 32;
  static constexpr int CPT_SD = (SD + 31) / 32;
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  // widest activation row any kernel builds in the scalar tile
  static constexpr int XW = cmax(cmax(cmax(K1, KE0), cmax(2 * F + R, F + 4 + R + F + 4)),
                                 cmax(2 * TOK + TD, S + AMAX + 8 + R));
  static constexpr int XLD = ((XW + 3) / 4) * 4 + 4;
  // vector-plane row pitches: K padded to 8 for the warp-MMA path, and pitch % 16 == 8 so that the 64-bit A-fragment loads /
  // accumulator stores of mma.m16n8k16 (rows g = 0..3 of a half warp, 2 words each) fall into four disjoint bank octets
  static constexpr int vpitch(int k) { return ((k + 7) / 8) * 8 + ((((k + 7) / 8) * 8) % 16 == 0 ? 8 : 0); }
  static constexpr int LDVA = vpitch(cmax(VIN0, V));
  static constexpr int LDVB = vpitch(H0 + 2 * CP);
  static_assert(S % 32 == 0 && F % 32 == 0, "S and F must be multiples of 32");
  static_assert(V % 4 == 0 && R % 4 == 0 && TOK % 4 == 0 && TD % 4 == 0, "V, R, token dims must be multiples of 4");
  // shared-memory plan of the tile kernels (floats)
  static constexpr int SM_XS = TM * XLD;
  static constexpr int SM_VA = 3 * TM * LDVA;
  static constexpr int SM_VB = 3 * TM * LDVB;
  static constexpr int SM_G = TM * 32;
  static constexpr int SM_MISC = 4 * TM;                    // src, dst, dist, valid
  static constexpr int SM_FLOATS = SM_XS + SM_VA + SM_VB + SM_G + WSTAGE_FLOATS + SM_MISC;
  static constexpr size_t SMEM_BYTES = (size_t)SM_FLOATS * 4;
};

using DimsFlowmol3 = Dims<256, 32, 128, 32, 4, 0, 0, 64, 64>;   // configs/flowmol3.yml:80-106
using DimsDev = Dims<64, 16, 64, 32, 4, 16, 4, 64, 64>;          // configs/dev.yml:78-108

// Runtime model description (device-visible, passed by value to kernels).
struct ModelRT {
  const float* w;            // packed weights (device)
  const long long* off;      // offset table (device)
  const float* eemb_table;   // [(n_bond_types+1)][F] edge-embedding outputs (weights-only constant)
  int A, C, EB;              // categories (mask index = count)
  int L, NU;                 // conv layers, updaters
  int convs_per_update, separate_updaters, self_cond, use_dst;
  float rbf_dmax, msg_norm;  // msg_norm: 0 => 'sum', -1 => 'mean', >0 => divide
  __device__ __forceinline__ const float* g(int id) const { return w + off[id]; }
  __device__ __forceinline__ const float* c(int l, int id) const { return w + off[G_COUNT + l * C_COUNT + id]; }
  __device__ __forceinline__ const float* u(int u_, int id) const { return w + off[G_COUNT + L * C_COUNT + u_ * U_COUNT + id]; }
};

// Batch descriptor: B independent complete molecular graphs (device arrays live in the workspace).
//   nodes: flat [N] in molecule order.  directed edges: internal dst-major order, each molecule's n(n-1) edges start at a
//   64-aligned slot (mol_epad) so that tile boundaries -- and therefore every reduction order -- depend only on the
//   molecule, never on its neighbours in the batch (bit-exact results under re-batching / sharding).
//   upper edges (i<j): compact [U] in the reference's order for state / predictions (C-ABI arrays), plus 64-aligned tiles.
struct BatchRT {
  int B, N, U;
  int n_edge_tiles, n_upper_tiles, n_node_tiles;
  long long EP;                 // padded directed-edge slots
  const int* mol_n;             // [B] atoms
  const int* mol_node;          // [B] first node
  const int* mol_u;             // [B] first compact upper edge
  const int* mol_etile;         // [B] first edge tile   (mol_epad = 64 * mol_etile)
  const int* mol_utile;         // [B] first upper tile
  const int* etile_mol;         // [n_edge_tiles]
  const int* utile_mol;         // [n_upper_tiles]
  const int* node_mol;          // [N]
};

struct GvpPtr { const float *whcp, *wu, *w, *b, *wg, *bg; };
__device__ __forceinline__ GvpPtr gvp_ptr_conv(const ModelRT& m, int l, int base) {
  return GvpPtr{m.c(l, base + GV_WHCP), m.c(l, base + GV_WU), m.c(l, base + GV_W), m.c(l, base + GV_B),
                m.c(l, base + GV_WG), m.c(l, base + GV_BG)};
}
__device__ __forceinline__ GvpPtr gvp_ptr_upd(const ModelRT& m, int u, int base) {
  return GvpPtr{m.u(u, base + GV_WHCP), m.u(u, base + GV_WU), m.u(u, base + GV_W), m.u(u, base + GV_B),
                m.u(u, base + GV_WG), m.u(u, base + GV_BG)};
}

}  // namespace fm
