// C ABI of the B200-native FlowMol sampling path (see include/flowmol_b200.h for the contract).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <utility>
#include <unordered_map>
#include <vector>

#include "../../include/flowmol_b200.h"
#include "ctmc.cuh"
#include "kernels.cuh"
#include "conv_tc.cuh"
#include "egemm_tc.cuh"
#include "egemm_p.cuh"
#include "egemm_e.cuh"
#include "egemm_c.cuh"
#include "egemm_g2.cuh"
#include "egemm_h.cuh"
#include "vec_stages.cuh"
#include "vec_reg.cuh"
#include "edge_reg.cuh"
#include "tc_test.cuh"

namespace {

thread_local std::string g_err;
int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}
#define CUDA_OK(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t e__ = (expr);                                                                              \
    if (e__ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));              \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// byte offsets of everything inside a workspace
struct Layout {
  int B = 0, N = 0, U = 0, nET = 0, nUT = 0, nNT = 0;
  long long EP = 0;
  size_t mol_n, mol_node, mol_u, mol_etile, mol_utile, etile_mol, utile_mol, node_mol;
  size_t s, v, x, P, Q, vd, EAB, M, partF, partL, ef;
  size_t SA, SB, VH, SH, GT;   // wide tensor-core pipeline intermediates (per padded edge slot, rounded up to 256 slots)
  size_t SHI;                  // vector norms of message GVPs 1 / 2 as the last k-slab's operand images (k_vecr_b -> k_egemm_g): 256 B per slot
  size_t EFI;                  // fp16 (hi, lo) operand images of the edge features (egemm_p.cuh), same bytes as ef
  long long EPA = 0;
  long long NPA = 0;           // node rows rounded up to 256
  size_t pred[3][4];     // [buffer][x,a,c,e]
  size_t total = 0;
};

struct Dyn {           // model dimensions known at run time
  int S, V, F, SD, VD, A, C, EB, MW;
};

Layout make_layout(const Dyn& d, const int32_t* n_atoms, int B) {
  Layout L;
  L.B = B;
  long long N = 0, U = 0, et = 0, ut = 0;
  for (int b = 0; b < B; ++b) {
    const long long n = n_atoms[b];
    N += n;
    U += n * (n - 1) / 2;
    et += (n * (n - 1) + fm::TM - 1) / fm::TM;
    ut += (n * (n - 1) / 2 + fm::TM - 1) / fm::TM;
  }
  L.N = (int)N; L.U = (int)U; L.nET = (int)et; L.nUT = (int)ut; L.nNT = (int)((N + fm::TM - 1) / fm::TM);
  L.EP = et * fm::TM;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.mol_n = take(4ull * B); L.mol_node = take(4ull * B); L.mol_u = take(4ull * B);
  L.mol_etile = take(4ull * B); L.mol_utile = take(4ull * B);
  L.etile_mol = take(4ull * L.nET); L.utile_mol = take(4ull * L.nUT); L.node_mol = take(4ull * N);
  L.s = take(4ull * N * d.S); L.v = take(4ull * N * 3 * d.V); L.x = take(4ull * N * 3);
  L.NPA = (N + 255) / 256 * 256;                  // node rows rounded up: k_egemm_tc stores whole 128-row tiles
  L.P = take(4ull * (size_t)L.NPA * d.S); L.Q = take(4ull * N * d.S * (d.SD > 0)); L.vd = take(4ull * N * 3 * d.VD);
  L.EAB = take(4ull * (size_t)L.NPA * 2 * d.F); L.M = take(4ull * N * d.MW);
  L.partF = take(8ull * L.nET * d.MW); L.partL = take(8ull * L.nET * d.MW);   // x2: the tensor-core kernel uses 32-row tiles
  L.EPA = (L.EP + 255) / 256 * 256;
  L.ef = take(4ull * (size_t)L.EPA * d.F);       // rounded up: the wide kernels store whole 128-slot tiles
  const size_t wide = d.S == 256 && d.SD == 0 ? (size_t)std::max<long long>(L.EPA, L.NPA) : 0;   // the node pipeline reuses these
  L.EFI = take(wide ? 4ull * (size_t)L.EPA * d.F : 0);
  L.SA = take(4ull * wide * d.S); L.SB = take(4ull * wide * d.S); L.VH = take(4ull * wide * 120); L.SH = take(4ull * wide * 40);
  L.GT = take(4ull * wide * 32);
  L.SHI = take(wide ? 256ull * (size_t)L.EPA : 0);
  for (int k = 0; k < 3; ++k) {
    L.pred[k][0] = take(4ull * N * 3); L.pred[k][1] = take(4ull * N * d.A); L.pred[k][2] = take(4ull * N * d.C);
    L.pred[k][3] = take(4ull * (size_t)U * d.EB);
  }
  L.total = off;
  return L;
}

}  // namespace

struct FmHandle {
  FmConfig cfg;
  int device = 0;
  int variant = 0;             // 0: flowmol3 dims, 1: dev dims
  Dyn dyn;
  float* d_w = nullptr;
  long long* d_off = nullptr;
  float* d_table = nullptr;
  fm::ModelRT rt;
  std::vector<long long> off_h;   // host copy of the weight offset table
  std::unordered_map<void*, Layout> batches;
  int64_t launches = 0;
  int eg_nh_gate = 1;
  int n_sm = 148;
  int eg_nh = 2;               // 128-edge halves per CTA of k_egemm_tc (2: 1 CTA/SM, 1: 2 CTAs/SM)
  long long* d_trace = nullptr;   // clock64 stamps of one egemm CTA (timeline experiments)
  int trace_cta = 0, trace_mode = 1;
  int tc_debug = 0;            // timing experiments (conv_tc.cuh TcCtx::dbg)
  bool has_tc = false;         // packed weights contain the UMMA operand images
  int fuse_agg = 1;            // scalar segment-sum in the epilogue of the last message linear (k_egemm_tc<EG_MSGA>)
  int node_impl = 0;           // 0: fused fp32 k_node_update, 1: node pipeline around k_egemm_tc (with conv_impl 2)
  int conv_impl = 0;           // 0: fp32 CUDA-core k_conv_edge, 1: tcgen05 3xTF32 k_conv_edge_tc (flowmol3 dims only)
  int eg_persist = 1;          // tc_prec 1: persistent role-specialised k_egemm_p (1 CTA / SM, double-buffered accumulators)
  int eg_img = 1;              // k_egemm_p: consecutive tensor-core linears hand their activations over as fp16 (hi, lo) operand images
  int eg_orient = 0;           // message linears MSG0 / MSG of the image chain: 0 = features on M (k_egemm_p, default), 1 = edges on M
                               // (k_egemm_e: bit-identical, measured slower -- MSG0 750 vs 588 us, MSG 591 vs 547 us, profiles/r02d)
  int eg_fuse_gate = 1;        // GVP 1 / 2 of the message pass: gate linear inside the scalar linear's kernel (k_egemm_g; needs the image chain
                               // and the register-resident vector stages); aggregation pieces become 32 rows
  int eg_cluster = 1;          // k_egemm_p on edge rows: CTAs per thread-block cluster sharing one multicast weight stream (1, 2, 4)
  int eg_clusters_seen = 0;    // cudaOccupancyMaxActiveClusters of the last cluster kernel configured (diagnostics)
  int eu_quad = 1;             // k_egemm_c with four threads per row in the epilogues (permuted features)
  int nemb_tc = 1;             // k_node_embed: its five 256-wide linears on mma.sync fp16x3 (tile_gemm_h16) in the fp16x3 tensor-core pipeline
  int nemb_ok = 0;             // ... the five weight matrices fit the fixed 2^10 scale (checked at fm_create)
  int eg_perm = 1;             // k_egemm_h<MSG> with permuted output features: image stores re-read from tensor memory (16-byte pieces)
  int eg_epi12 = 1;            // gate-fused edge-row linears with twelve epilogue warps (k_egemm_h; needs sh_img)
  int sh_img = 1;              // norms of message GVPs 1 / 2 as operand images: every k-slab of k_egemm_g is a bulk copy
  int node_fuse_gate = 1;      // node-row GVPs: scalar + gate linear in one k_egemm_g launch
  int pdl = 0;                 // programmatic dependent launch of every pipeline kernel (launch_k); measured 3 % slower end to end (profiles/r02z): off
  int eg_pair = 0;             // gate-fused message linears on CTA pairs (tcgen05 cta_group::2, egemm_g2.cuh)
  int edge_reg = 1;            // upper-edge MLPs (edge self-conditioning residual, bond-order head) as register-resident warp kernels (edge_reg.cuh)
  int node_img = 1;            // node-row GVP chains: operand images between the three scalar linears (as on edge rows)
  int eu_fuse = 1;             // EdgeUpdate: both linears + LayerNorm in one kernel, hidden activations in tensor memory (egemm_c.cuh)
  int vec_impl = 1;            // edge-row vector stages: 1 = register-resident warp units (vec_reg.cuh, needs the image chain), 0 = vec_stages.cuh
  int tc_prec = 0;             // operand format of k_egemm_tc: 0 = 3xTF32 images, 1 = scaled fp16 hi/lo images ("fp16x3")
  bool has_h16 = false;        // packed weights carry the fp16 images
  int* d_status = nullptr;     // device status word: bit 0 = an activation left the fp16 operand range (tc_prec 1)
  cudaStream_t cap_stream = nullptr;    // private stream for CUDA-graph capture (the legacy default stream cannot capture)
  int kprof = 0;                        // "kprof" option: a CUDA event after every launch of fm_forward (fm_debug_kprof reads them)
  std::vector<std::pair<int, cudaEvent_t>> prof;      // (api.cu line of the launch, event recorded right after it)
};

namespace {

template <class T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }

fm::BatchRT batch_rt(void* ws, const Layout& L) {
  fm::BatchRT b;
  b.B = L.B; b.N = L.N; b.U = L.U; b.n_edge_tiles = L.nET; b.n_upper_tiles = L.nUT; b.n_node_tiles = L.nNT; b.EP = L.EP;
  b.mol_n = at<int>(ws, L.mol_n); b.mol_node = at<int>(ws, L.mol_node); b.mol_u = at<int>(ws, L.mol_u);
  b.mol_etile = at<int>(ws, L.mol_etile); b.mol_utile = at<int>(ws, L.mol_utile);
  b.etile_mol = at<int>(ws, L.etile_mol); b.utile_mol = at<int>(ws, L.utile_mol); b.node_mol = at<int>(ws, L.node_mol);
  return b;
}

fm::PredPtr pred_ptr(void* ws, const Layout& L, int k) {
  return fm::PredPtr{at<float>(ws, L.pred[k][0]), at<float>(ws, L.pred[k][1]), at<float>(ws, L.pred[k][2]),
                     at<float>(ws, L.pred[k][3])};
}

template <class D>
int set_smem_attrs() {
  const int bytes = (int)D::SMEM_BYTES;
  CUDA_OK(cudaFuncSetAttribute(fm::k_node_embed<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if constexpr (D::S == 256)
    CUDA_OK(cudaFuncSetAttribute(fm::k_node_embed<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::NodeEmbedMmaSmem<D>::BYTES));
  CUDA_OK(cudaFuncSetAttribute(fm::k_edge_init<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EdgeSmem<D>::BYTES));
  CUDA_OK(cudaFuncSetAttribute(fm::k_conv_edge<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_OK(cudaFuncSetAttribute(fm::k_node_update<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_OK(cudaFuncSetAttribute(fm::k_dst_proj<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_OK(cudaFuncSetAttribute(fm::k_edge_update<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_OK(cudaFuncSetAttribute(fm::k_node_head<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_OK(cudaFuncSetAttribute(fm::k_edge_head<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EdgeSmem<D>::BYTES));
  if constexpr (D::S == 256 && D::V == 32 && D::SD == 0) {
    CUDA_OK(cudaFuncSetAttribute(fm::k_conv_edge_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::TcPlan<D>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG0, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<2>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG0, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<2>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<2>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<2>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_GATE, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<2>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_GATE, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<2>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSG, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_GATE, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_GATE, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_EU1, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_EU1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_EU2, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_EU2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_vec_a<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::VecSmem<D>::BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_vec_b<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::VecSmem<D>::BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_node_pre<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::VecSmem<D>::BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_node_mid<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::VecSmem<D>::BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_node_post<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::VecSmem<D>::BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSGA, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_MSGA, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_LIN, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_tc<D, fm::EG_LIN, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgPlan<1>::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_vec_c<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::VecSmem<D>::BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSG0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_EU1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_EU2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_LIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g<D, fm::EG_MSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g<D, fm::EG_MSGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_h<D, fm::EG_MSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EghPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_h<D, fm::EG_MSG, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EghPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_h<D, fm::EG_MSGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EghPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g<D, fm::EG_MSG, 1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g<D, fm::EG_MSGA, 1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g<D, fm::EG_MSG, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g<D, fm::EG_MSG, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g2<D, fm::EG_MSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_g2<D, fm::EG_MSGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EggPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_edge_head_r<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EdgeRegSmem<D>::HEAD_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_edge_init_r<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EdgeRegSmem<D>::INIT_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_c<D, fm::CH_EU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgcPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_c<D, fm::CH_EU, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgcPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_e<D, fm::EG_MSG0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgePlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_e<D, fm::EG_MSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgePlan::SMEM_BYTES));
    constexpr int IO = fm::EGI_IN | fm::EGI_OUT;
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSG0, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSG, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSG, fm::EGI_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSG, fm::EGI_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_MSGA, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_GATE, fm::EGI_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_EU1, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(fm::k_egemm_p<D, fm::EG_EU2, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES));
  }
  return 0;
}


// Kernel launch, optionally as a programmatic dependent launch (option "pdl"; default off: measured 7573 vs 7339 ms per 250-step
// trajectory, profiles/r02z -- the kernels fill the SMs' shared memory, so a dependent CTA cannot become resident before its
// predecessor's CTA on that SM has exited, and the early-launched grid only adds scheduling work): the kernel's CTAs may start while its predecessor
// in the stream drains; every kernel begins with pdl_launch(); ... pdl_wait() (common.cuh) before it touches activation data.
template <class... KArgs, class... Args>
inline void launch_k(FmHandle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (h && h->pdl) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// k_egemm_tc in the operand precision selected on the handle
// in-situ per-launch timing (option "kprof"): an event after every launch; consecutive differences are the launches' durations in
// the warm pipeline (what ncu's serialised cold-cache replays cannot show).  Line 0 = start marker.
inline void prof_mark(FmHandle* h, int line, cudaStream_t st) {
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  h->prof.emplace_back(line, e);
}

// operand-image hand-over between consecutive linears: only k_egemm_p knows it (callers pass IMG != 0 only when img_on(h))
inline bool img_on(const FmHandle* h) { return h->eg_img && h->tc_prec == 1 && h->eg_persist && h->eg_nh == 1 && h->eg_nh_gate == 1 && h->fuse_agg; }
// gate-fused message linears (egemm_e.cuh:k_egemm_g): on top of the image chain and the register-resident vector stages
inline bool gate_fused(const FmHandle* h) { return img_on(h) && h->vec_impl == 1 && h->eg_fuse_gate == 1 && h->eg_cluster == 1 && h->conv_impl == 2; }
// k_egemm_p as thread-block clusters of CL CTAs that share one multicast weight stream (egemm_p.cuh).  The grid is the number of
// clusters the device can hold at once (cudaOccupancyMaxActiveClusters: a cluster lives inside one GPC) times CL.
template <class D, int MODE, int IMG, int CL>
void launch_egp_cluster(FmHandle* h, int n_tiles, cudaStream_t st, const fm::ModelRT& m, const fm::BatchRT& bt, const fm::EgArgs& a) {
  auto kern = fm::k_egemm_p<D, MODE, IMG, CL>;
  static int max_clusters = -1;                 // per instantiation; every rank of a job drives the same kind of device
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(fm::EgpPlan::THREADS);
  cfg.dynamicSmemBytes = fm::EgpPlan::SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fm::EgpPlan::SMEM_BYTES);
    cfg.gridDim = dim3(CL * (h->n_sm / CL));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = 1;
    max_clusters = n;
    h->eg_clusters_seen = n;
  }
  const int want = (n_tiles + CL - 1) / CL;
  cfg.gridDim = dim3(CL * (want < max_clusters ? want : max_clusters));
  cudaLaunchKernelEx(&cfg, kern, m, bt, a, n_tiles);
}

template <class D, int MODE, int NH, int IMG = 0>
void launch_eg(FmHandle* h, int grid, cudaStream_t st, const fm::ModelRT& m, const fm::BatchRT& bt, fm::EgArgs a) {
  using PL = fm::EgPlan<NH>;
  a.status = h->d_status;
  if (NH == 1 && h->tc_prec == 1 && h->eg_persist) {      // `grid` = number of 128-row tiles
    if constexpr (IMG == (fm::EGI_IN | fm::EGI_OUT)) {     // the edge-row linears of the image chain: optional cluster launch
      if (h->eg_cluster == 2 && !(a.flags & fm::EGF_NODE_ROWS)) { launch_egp_cluster<D, MODE, IMG, 2>(h, grid, st, m, bt, a); return; }
      if (h->eg_cluster == 4 && !(a.flags & fm::EGF_NODE_ROWS)) { launch_egp_cluster<D, MODE, IMG, 4>(h, grid, st, m, bt, a); return; }
    }
    launch_k(h, fm::k_egemm_p<D, MODE, IMG>, grid < h->n_sm ? grid : h->n_sm, fm::EgpPlan::THREADS, fm::EgpPlan::SMEM_BYTES, st, m, bt, a, grid);
    return;
  }
  if (h->tc_prec) launch_k(h, fm::k_egemm_tc<D, MODE, NH, 1>, grid, PL::THREADS, PL::SMEM_BYTES, st, m, bt, a);
  else launch_k(h, fm::k_egemm_tc<D, MODE, NH, 0>, grid, PL::THREADS, PL::SMEM_BYTES, st, m, bt, a);
}
// id of a tensor-core image entry in the selected precision (the fp16 twins follow the TF32 entries, weight_layout.py)
inline int tc_c(const FmHandle* h, int id) { return id + (h->tc_prec ? (int)fm::C_MSG0_TCW_H - (int)fm::C_MSG0_TCW : 0); }
inline int tc_u(const FmHandle* h, int id) { return id + (h->tc_prec ? (int)fm::U_EUPD_TC1_H - (int)fm::U_EUPD_TC1 : 0); }

// every call site has the launching stream in a local named `st`
#define LAUNCH_OK(h)                                                                                       \
  do {                                                                                                     \
    ++(h)->launches;                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                                  \
    if (e__ != cudaSuccess) return fail(std::string("kernel launch failed: ") + cudaGetErrorString(e__)); \
    if ((h)->kprof) prof_mark((h), __LINE__, st);                                                          \
  } while (0)

// message pass of conv `l` as the wide tensor-core pipeline (egemm_tc.cuh + vec_stages.cuh): 10 launches
template <class D>
int conv_wide(FmHandle* h, void* ws, const Layout& L, const fm::BatchRT& bt, int l, cudaStream_t st) {
  if constexpr (D::S == 256 && D::V == 32 && D::SD == 0) {
    const fm::ModelRT& m = h->rt;
    auto wptr = [&](int id) { return h->d_w + h->off_h[fm::G_COUNT + l * fm::C_COUNT + id]; };
    float *v = at<float>(ws, L.v), *x = at<float>(ws, L.x), *P = at<float>(ws, L.P), *M = at<float>(ws, L.M);
    float *partF = at<float>(ws, L.partF), *partL = at<float>(ws, L.partL), *ef = at<float>(ws, L.ef);
    float *SA = at<float>(ws, L.SA), *SB = at<float>(ws, L.SB), *VH = at<float>(ws, L.VH), *SH = at<float>(ws, L.SH), *GT = at<float>(ws, L.GT);
    const int NHsel = h->eg_nh;
    const int gt = (int)(L.EPA / (128 * NHsel));
    const bool img = img_on(h);
    const size_t vsm = fm::VecSmem<D>::BYTES;
    const int vgrid = L.nET < 2 * h->n_sm ? L.nET : 2 * h->n_sm;     // persistent: 2 CTAs per SM, tiles strided
    // register-resident vector stages (vec_reg.cuh): one warp per 16-row unit, VH holds VU = Vh_ext Wu (3 x 32 per row) instead
    const bool vr = img && h->vec_impl == 1;
    const bool gf = gate_fused(h);
    const bool shi = gf && h->sh_img && !h->eg_pair;           // norms of GVP 1 / 2 handed over as operand images (k_vecr_b -> k_egemm_g)
    float* SHI = at<float>(ws, L.SHI);
    const int n_units = L.nET * (fm::TM / fm::UR);
    const int vrgrid = (n_units + fm::NWARP - 1) / fm::NWARP < 2 * h->n_sm ? (n_units + fm::NWARP - 1) / fm::NWARP : 2 * h->n_sm;
    if (!vr) launch_k(h, fm::k_vec_a<D>, vgrid, fm::NT, vsm, st, m, bt, l, x, v, VH, SH);
    else launch_k(h, fm::k_vecr_a<D>, vrgrid, fm::NT, 0, st, m, bt, l, n_units, x, v, VH, SH);
    LAUNCH_OK(h);
    const int tcw[3] = {fm::C_MSG0_TCW, fm::C_MSG1_TCW, fm::C_MSG2_TCW}, tcg[3] = {fm::C_MSG0_TCG, fm::C_MSG1_TCG, fm::C_MSG2_TCG};
    const int gb[3] = {fm::C_MSG0_WHCP, fm::C_MSG1_WHCP, fm::C_MSG2_WHCP};
    float* cur = ef;      // input activations of the current scalar linear
    float* outs[3] = {SA, SB, SA};
    for (int g = 0; g < 3; ++g) {
      fm::EgArgs a{wptr(tc_c(h, tcw[g])), wptr(gb[g] + fm::GV_B), cur, SH, P, x, outs[g], nullptr, nullptr, L.EP, h->trace_mode == (g == 0 ? 0 : 1) ? h->d_trace : nullptr, h->trace_cta, 0, h->tc_debug, M, partF, partL};
      if (img) {                         // fp16 (hi, lo) operand images between the linears (k_egemm_p only; implies fuse_agg)
        constexpr int IO = fm::EGI_IN | fm::EGI_OUT;
        a.in_img = g == 0 ? at<float>(ws, L.EFI) : cur;
        a.out_img = outs[g];
        if (g >= 1 && gf) {                                          // scalar linear + gate linear in one kernel (egemm_e.cuh)
          a.status = h->d_status;
          a.g_units = wptr(tc_c(h, tcg[g])); a.g_bias = wptr(gb[g] + fm::GV_BG); a.g_out = GT;
          const int grid_g = gt < h->n_sm ? gt : h->n_sm;
          if (h->eg_pair && gt % 2 == 0 && h->n_sm >= 2) {             // CTA pairs on one weight stream (egemm_g2.cuh)
            const int grid_p = 2 * (gt / 2 < h->n_sm / 2 ? gt / 2 : h->n_sm / 2);
            if (g == 1) launch_k(h, fm::k_egemm_g2<D, fm::EG_MSG>, grid_p, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
            else launch_k(h, fm::k_egemm_g2<D, fm::EG_MSGA>, grid_p, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
          } else
          if (shi && h->eg_epi12) {                                    // all-image inputs: twelve epilogue warps (egemm_h.cuh)
            a.sh_img = SHI;
            if (g == 1 && h->eg_perm && h->off_h[fm::G_COUNT + l * fm::C_COUNT + fm::C_MSG1_TCW_HP] >= 0) {
              // output features permuted inside every 32-chunk: image stores from the packed words in tensor memory, 16 bytes per lane
              a.units = wptr(fm::C_MSG1_TCW_HP); a.bias = wptr(fm::C_MSG1_BP); a.g_units = wptr(fm::C_MSG1_TCG_HP);
              launch_k(h, fm::k_egemm_h<D, fm::EG_MSG, 1>, grid_g, fm::EghPlan::THREADS, fm::EghPlan::SMEM_BYTES, st, m, bt, a, gt);
            } else
            if (g == 1) launch_k(h, fm::k_egemm_h<D, fm::EG_MSG>, grid_g, fm::EghPlan::THREADS, fm::EghPlan::SMEM_BYTES, st, m, bt, a, gt);
            else launch_k(h, fm::k_egemm_h<D, fm::EG_MSGA>, grid_g, fm::EghPlan::THREADS, fm::EghPlan::SMEM_BYTES, st, m, bt, a, gt);
          } else
          if (shi) {
            a.sh_img = SHI;
            if (g == 1) launch_k(h, fm::k_egemm_g<D, fm::EG_MSG, 1, 0, 1>, grid_g, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
            else launch_k(h, fm::k_egemm_g<D, fm::EG_MSGA, 1, 0, 1>, grid_g, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
          } else
          if (g == 1) launch_k(h, fm::k_egemm_g<D, fm::EG_MSG>, grid_g, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
          else launch_k(h, fm::k_egemm_g<D, fm::EG_MSGA>, grid_g, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
        } else
        if (g < 2 && h->eg_orient == 1 && h->eg_cluster == 1) {      // edges-on-M orientation (egemm_e.cuh)
          a.status = h->d_status;
          const int grid_e = gt < h->n_sm ? gt : h->n_sm;
          if (g == 0) launch_k(h, fm::k_egemm_e<D, fm::EG_MSG0>, grid_e, fm::EgePlan::THREADS, fm::EgePlan::SMEM_BYTES, st, m, bt, a, gt);
          else launch_k(h, fm::k_egemm_e<D, fm::EG_MSG>, grid_e, fm::EgePlan::THREADS, fm::EgePlan::SMEM_BYTES, st, m, bt, a, gt);
        } else if (g == 0) {
          launch_eg<D, fm::EG_MSG0, 1, IO>(h, gt, st, m, bt, a);
        } else if (g == 1) {
          launch_eg<D, fm::EG_MSG, 1, IO>(h, gt, st, m, bt, a);
        } else {
          launch_eg<D, fm::EG_MSGA, 1, IO>(h, gt, st, m, bt, a);
        }
      } else
      if (g == 2 && h->fuse_agg) {       // last scalar linear: the segment-sum over in-edges rides in the epilogue
        launch_eg<D, fm::EG_MSGA, 1>(h, (int)(L.EPA / 128), st, m, bt, a);
      } else
      if (NHsel == 2) {
        if (g == 0) launch_eg<D, fm::EG_MSG0, 2>(h, gt, st, m, bt, a);
        else launch_eg<D, fm::EG_MSG, 2>(h, gt, st, m, bt, a);
      } else {
        if (g == 0) launch_eg<D, fm::EG_MSG0, 1>(h, gt, st, m, bt, a);
        else launch_eg<D, fm::EG_MSG, 1>(h, gt, st, m, bt, a);
      }
      LAUNCH_OK(h);
      if (!(g >= 1 && gf)) {
        fm::EgArgs ag{wptr(tc_c(h, tcg[g])), wptr(gb[g] + fm::GV_BG), outs[g], nullptr, nullptr, nullptr, GT, nullptr, nullptr, L.EP, h->trace_mode == 2 ? h->d_trace : nullptr, h->trace_cta, 0, h->tc_debug};
        ag.in_img = outs[g];
        if (img) launch_eg<D, fm::EG_GATE, 1, fm::EGI_IN>(h, (int)(L.EPA / 128), st, m, bt, ag);
        else if (h->eg_nh_gate == 2) launch_eg<D, fm::EG_GATE, 2>(h, (int)(L.EPA / 256), st, m, bt, ag);
        else launch_eg<D, fm::EG_GATE, 1>(h, (int)(L.EPA / 128), st, m, bt, ag);
        LAUNCH_OK(h);
      }
      if (g < 2) {
        if (!vr)
          launch_k(h, fm::k_vec_b<D>, vgrid, fm::NT, vsm, st, bt, wptr(g == 0 ? fm::C_MSG0_WU : fm::C_MSG1_WU), (g == 0 ? D::H0 : D::V) + D::CP,
                                                     wptr(g == 0 ? fm::C_MSG1_WHCP : fm::C_MSG2_WHCP), 0, VH, SH, GT);
        else
          launch_k(h, fm::k_vecr_b<D>, vrgrid, fm::NT, 0, st, bt, wptr(g == 0 ? fm::C_MSG1_WHCP : fm::C_MSG2_WHCP),
                                                      wptr(g == 0 ? fm::C_MSG1_WU : fm::C_MSG2_WU), n_units, VH, SH, GT,
                                                      shi ? SHI : (float*)nullptr);
        LAUNCH_OK(h);
      }
      cur = outs[g];
    }
    if (!vr) launch_k(h, fm::k_vec_c<D>, vgrid, fm::NT, vsm, st, m, bt, l, h->fuse_agg ? D::S : 0, VH, GT, SA, M, partF, partL);
    else launch_k(h, fm::k_vecr_c<D>, (L.nET * (gf ? 2 : 1) + fm::NWARP - 1) / fm::NWARP, fm::NT, 0, st, bt, gf ? 32 : fm::TM, VH, GT, M, partF, partL);
    LAUNCH_OK(h);
  }
  return 0;
}

// node update of conv `l` (+ NodePositionUpdate of `upd`) as a pipeline: the six GVPs' scalar / gate linears and the per-node
// halves of the next edge phases run through k_egemm_tc on node rows, the vector stages in vec_stages.cuh.  The edge-sized
// scratch buffers (SA, SB, VH, SH, GT) are idle between two message passes and are reused.
template <class D>
int node_wide(FmHandle* h, void* ws, const Layout& L, const fm::BatchRT& bt, int l, int upd, int has_next, int agg_rows, cudaStream_t st) {
  if constexpr (D::S == 256 && D::V == 32 && D::SD == 0 && D::F == 128) {
    const fm::ModelRT& m = h->rt;
    auto wptr = [&](int ll, int id) { return h->d_w + h->off_h[fm::G_COUNT + ll * fm::C_COUNT + id]; };
    auto uptr = [&](int id) { return h->d_w + h->off_h[fm::G_COUNT + m.L * fm::C_COUNT + upd * fm::U_COUNT + id]; };
    float *s = at<float>(ws, L.s), *v = at<float>(ws, L.v), *x = at<float>(ws, L.x), *P = at<float>(ws, L.P), *M = at<float>(ws, L.M);
    float *EAB = at<float>(ws, L.EAB), *partF = at<float>(ws, L.partF), *partL = at<float>(ws, L.partL);
    float *SA = at<float>(ws, L.SA), *SB = at<float>(ws, L.SB), *VH = at<float>(ws, L.VH), *SH = at<float>(ws, L.SH), *GT = at<float>(ws, L.GT);
    const size_t vsm = fm::VecSmem<D>::BYTES;
    const int gt = (int)(L.NPA / 128);
    using PL = fm::EgPlan<1>;
    // hand-over between the three GVPs of a chain as operand images, like the edge rows (node_img): GVP 0 reads fp32 rows and writes
    // images, GVP 1 reads and writes images, GVP 2 reads images and writes fp32 rows (k_node_mid / nobody reads them as images)
    const bool nimg = img_on(h) && h->node_img;
    auto scalar = [&](const float* units, const float* bias, const float* in, float* out, int g) {
      fm::EgArgs a{units, bias, in, SH, nullptr, nullptr, out, nullptr, nullptr, (long long)L.N, nullptr, 0, fm::EGF_NODE_ROWS, 0};
      a.in_img = in; a.out_img = out;
      if (!nimg) launch_eg<D, fm::EG_MSG, 1>(h, gt, st, m, bt, a);
      else if (g == 0) launch_eg<D, fm::EG_MSG, 1, fm::EGI_OUT>(h, gt, st, m, bt, a);
      else if (g == 1) launch_eg<D, fm::EG_MSG, 1, fm::EGI_IN | fm::EGI_OUT>(h, gt, st, m, bt, a);
      else launch_eg<D, fm::EG_MSG, 1, fm::EGI_IN>(h, gt, st, m, bt, a);
    };
    auto gate = [&](const float* units, const float* bias, const float* in, int identity, int g) {
      fm::EgArgs a{units, bias, in, nullptr, nullptr, nullptr, GT, nullptr, nullptr, (long long)L.N, nullptr, 0,
                   fm::EGF_NODE_ROWS | (identity ? fm::EGF_IDENTITY : 0), 0};
      a.in_img = in;
      if (nimg && g < 2) launch_eg<D, fm::EG_GATE, 1, fm::EGI_IN>(h, gt, st, m, bt, a);
      else launch_eg<D, fm::EG_GATE, 1>(h, gt, st, m, bt, a);
    };
    auto linear = [&](const float* units, const float* bias, float* out) {
      fm::EgArgs a{units, bias, s, nullptr, nullptr, nullptr, out, nullptr, nullptr, (long long)L.N, nullptr, 0, fm::EGF_NODE_ROWS, 0};
      launch_eg<D, fm::EG_LIN, 1>(h, gt, st, m, bt, a);
    };
    // scalar linear + gate linear of a node-row GVP in ONE launch (k_egemm_g on node rows: fp32 rows in for the first GVP of a chain,
    // fp32 rows out for the last): 33 launches of ~17 us less per evaluation
    const bool nfuse = nimg && gate_fused(h) && h->node_fuse_gate;
    auto sgate = [&](const float* units, const float* bias, const float* g_units, const float* g_bias, const float* in, float* out, int g, int identity) {
      fm::EgArgs a{units, bias, in, SH, nullptr, nullptr, out, nullptr, nullptr, (long long)L.N, nullptr, 0,
                   fm::EGF_NODE_ROWS | (identity ? fm::EGF_IDENTITY : 0), 0};
      a.status = h->d_status; a.in_img = in; a.out_img = out; a.g_units = g_units; a.g_bias = g_bias; a.g_out = GT;
      const int grid = gt < h->n_sm ? gt : h->n_sm;
      if (g == 0) launch_k(h, fm::k_egemm_g<D, fm::EG_MSG, 0, 0>, grid, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
      else if (g == 1) launch_k(h, fm::k_egemm_g<D, fm::EG_MSG, 1, 0>, grid, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
      else launch_k(h, fm::k_egemm_g<D, fm::EG_MSG, 1, 1>, grid, fm::EggPlan::THREADS, fm::EggPlan::SMEM_BYTES, st, m, bt, a, gt);
    };
    const int vgrid = L.nNT < 2 * h->n_sm ? L.nNT : 2 * h->n_sm;
    launch_k(h, fm::k_node_pre<D>, L.nNT, fm::NT, vsm, st, m, bt, l, agg_rows, s, v, M, partF, partL, VH, SH);
    LAUNCH_OK(h);
    const int utw[3] = {fm::C_UPD0_TCW, fm::C_UPD1_TCW, fm::C_UPD2_TCW}, utg[3] = {fm::C_UPD0_TCG, fm::C_UPD1_TCG, fm::C_UPD2_TCG};
    const int ub[3] = {fm::C_UPD0_WHCP, fm::C_UPD1_WHCP, fm::C_UPD2_WHCP};
    const float* cur = s;
    float* outs[3] = {SA, SB, SA};
    for (int g = 0; g < 3; ++g) {
      if (nfuse) {
        sgate(wptr(l, tc_c(h, utw[g])), wptr(l, ub[g] + fm::GV_B), wptr(l, tc_c(h, utg[g])), wptr(l, ub[g] + fm::GV_BG), cur, outs[g], g, 0); LAUNCH_OK(h);
      } else {
        scalar(wptr(l, tc_c(h, utw[g])), wptr(l, ub[g] + fm::GV_B), cur, outs[g], g); LAUNCH_OK(h);
        gate(wptr(l, tc_c(h, utg[g])), wptr(l, ub[g] + fm::GV_BG), outs[g], 0, g); LAUNCH_OK(h);
      }
      if (g < 2) {
        launch_k(h, fm::k_vec_b<D>, vgrid, fm::NT, vsm, st, bt, wptr(l, ub[g] + fm::GV_WU), D::V + D::CP, wptr(l, ub[g + 1] + fm::GV_WHCP), 1, VH, SH, GT);
        LAUNCH_OK(h);
      }
      cur = outs[g];
    }
    launch_k(h, fm::k_node_mid<D>, L.nNT, fm::NT, vsm, st, m, bt, l, upd, s, v, SA, VH, SH, GT);
    LAUNCH_OK(h);
    if (has_next) { linear(wptr(l + 1, tc_c(h, fm::C_WSRC_TC)), wptr(l + 1, fm::C_BSRC), P); LAUNCH_OK(h); }
    if (upd >= 0) {
      linear(uptr(tc_u(h, fm::U_EUPD_WN_TC)), uptr(fm::U_EUPD_BN), EAB); LAUNCH_OK(h);
      const int ptw[3] = {fm::U_POS0_TCW, fm::U_POS1_TCW, fm::U_POS2_TCW}, ptg[3] = {fm::U_POS0_TCG, fm::U_POS1_TCG, fm::U_POS2_TCG};
      const int pb[3] = {fm::U_POS0_WHCP, fm::U_POS1_WHCP, fm::U_POS2_WHCP};
      cur = s;
      for (int g = 0; g < 3; ++g) {
        if (nfuse) {
          sgate(uptr(tc_u(h, ptw[g])), uptr(pb[g] + fm::GV_B), uptr(tc_u(h, ptg[g])), uptr(pb[g] + fm::GV_BG), cur, outs[g], g, g == 2); LAUNCH_OK(h);
        } else {
          scalar(uptr(tc_u(h, ptw[g])), uptr(pb[g] + fm::GV_B), cur, outs[g], g); LAUNCH_OK(h);
          gate(uptr(tc_u(h, ptg[g])), uptr(pb[g] + fm::GV_BG), outs[g], g == 2, g); LAUNCH_OK(h);
        }
        if (g < 2) {
          launch_k(h, fm::k_vec_b<D>, vgrid, fm::NT, vsm, st, bt, uptr(pb[g] + fm::GV_WU), D::V + D::CP, uptr(pb[g + 1] + fm::GV_WHCP), 1, VH, SH, GT);
          LAUNCH_OK(h);
        }
        cur = outs[g];
      }
      launch_k(h, fm::k_node_post<D>, L.nNT, fm::NT, vsm, st, m, bt, upd, x, VH, GT);
      LAUNCH_OK(h);
    }
  }
  return 0;
}

// one denoise_graph pass (embedding + optional self-conditioning residual + convs + heads)
template <class D>
int run_pass(FmHandle* h, void* ws, const Layout& L, const float* x_t, const uint8_t* a_t, const uint8_t* c_t,
             const uint8_t* e_t, float t, const fm::PredPtr prev, int has_prev, const fm::PredPtr out, int remove_com,
             int stop_after, cudaStream_t st) {
  const fm::BatchRT bt = batch_rt(ws, L);
  const fm::ModelRT& m = h->rt;
  const size_t smem = D::SMEM_BYTES;
  float *s = at<float>(ws, L.s), *v = at<float>(ws, L.v), *x = at<float>(ws, L.x), *P = at<float>(ws, L.P);
  float *Q = at<float>(ws, L.Q), *vd = at<float>(ws, L.vd), *EAB = at<float>(ws, L.EAB), *M = at<float>(ws, L.M);
  float *partF = at<float>(ws, L.partF), *partL = at<float>(ws, L.partL), *ef = at<float>(ws, L.ef);
  bool embed_done = false;
  if constexpr (D::S == 256) {
    if (h->nemb_tc && h->nemb_ok && h->conv_impl == 2 && h->tc_prec == 1) {
      launch_k(h, fm::k_node_embed<D, 1>, L.nNT, fm::NT, fm::NodeEmbedMmaSmem<D>::BYTES, st, m, bt, x_t, a_t, c_t, t, prev, has_prev, s, v, P, h->d_status);
      embed_done = true;
    }
  }
  if (!embed_done) launch_k(h, fm::k_node_embed<D>, L.nNT, fm::NT, smem, st, m, bt, x_t, a_t, c_t, t, prev, has_prev, s, v, P, (int*)nullptr);
  LAUNCH_OK(h);
  if (m.use_dst) { launch_k(h, fm::k_dst_proj<D>, L.nNT, fm::NT, smem, st, m, bt, 0, s, v, Q, vd); LAUNCH_OK(h); }
  bool edge_done = false;
  int er_grid = 0;
  if constexpr (D::S == 256 && D::V == 32 && D::SD == 0 && D::F == 128 && D::R == 32) {
    // register-resident upper-edge kernels (edge_reg.cuh): 16-row units, one warp each, 2 CTAs per SM
    const int n_units = L.nUT * (fm::TM / fm::UR), want = (n_units + fm::NWARP - 1) / fm::NWARP;
    er_grid = want < 2 * h->n_sm ? want : 2 * h->n_sm;
    if (h->edge_reg && has_prev && m.self_cond && h->conv_impl == 2 && img_on(h) && h->vec_impl == 1 && m.EB <= 7) {
      // self-conditioning residual + both output formats (fp32 rows, operand images) in one kernel; dead rows were zeroed by fm_batch_init
      launch_k(h, fm::k_edge_init_r<D>, er_grid, fm::NT, fm::EdgeRegSmem<D>::INIT_BYTES, st, m, bt, n_units, x_t, e_t, prev, ef, at<float>(ws, L.EFI));
      LAUNCH_OK(h);
      edge_done = true;
    }
  }
  if (!edge_done) {
  launch_k(h, fm::k_edge_init<D>, L.nUT, fm::NT, fm::EdgeSmem<D>::BYTES, st, m, bt, x_t, e_t, prev, has_prev, ef);
  LAUNCH_OK(h);
  if constexpr (D::S == 256 && D::V == 32 && D::SD == 0 && D::F == 128) {
    if (h->conv_impl == 2 && img_on(h)) {        // entry of the operand-image chain (egemm_p.cuh)
      const int nt = (int)(L.EPA / 128);
      launch_k(h, fm::k_ef_image<D>, nt < 4 * h->n_sm ? nt : 4 * h->n_sm, 256, 0, st, bt, ef, at<float>(ws, L.EFI), L.EP, nt);
      LAUNCH_OK(h);
    }
  }
  }
  CUDA_OK(cudaMemcpyAsync(x, x_t, sizeof(float) * 3 * L.N, cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < m.L; ++l) {
    int agg_rows = fm::TM;
    if constexpr (D::S == 256 && D::V == 32 && D::SD == 0) {
      if (h->conv_impl == 1) {
        launch_k(h, fm::k_conv_edge_tc<D>, 2 * L.nET, fm::NT, fm::TcPlan<D>::SMEM_BYTES, st, m, bt, l, x, v, ef, P, M, partF, partL, h->tc_debug);
        agg_rows = fm::TCT;
        LAUNCH_OK(h);
      }
    }
    if (h->conv_impl == 2) {
      int rc = conv_wide<D>(h, ws, L, bt, l, st);
      if (rc) return rc;
      if (gate_fused(h)) agg_rows = 32;          // k_egemm_g<EG_MSGA> / k_vecr_c write 32-row aggregation pieces
    } else if (agg_rows == fm::TM) {
      launch_k(h, fm::k_conv_edge<D>, L.nET, fm::NT, smem, st, m, bt, l, x, v, ef, P, Q, vd, M, partF, partL);
      LAUNCH_OK(h);
    }
    int upd = -1;
    if (l != 0 && (l + 1) % m.convs_per_update == 0) upd = m.separate_updaters ? l / m.convs_per_update : 0;   // vector_field.py:321-326
    const int has_next = l + 1 < m.L;
    if (h->conv_impl == 2 && h->node_impl == 1) {
      int rc = node_wide<D>(h, ws, L, bt, l, upd, has_next, agg_rows, st);
      if (rc) return rc;
    } else {
      launch_k(h, fm::k_node_update<D>, L.nNT, fm::NT, smem, st, m, bt, l, upd, has_next, agg_rows, s, v, x, M, partF, partL, P, EAB);
      LAUNCH_OK(h);
    }
    if (m.use_dst && has_next) { launch_k(h, fm::k_dst_proj<D>, L.nNT, fm::NT, smem, st, m, bt, l + 1, s, v, Q, vd); LAUNCH_OK(h); }
    if (upd >= 0) {
      bool done = false;
      if constexpr (D::S == 256 && D::V == 32 && D::SD == 0 && D::F == 128) {
        if (h->conv_impl == 2) {           // EdgeUpdate as two wide tensor-core linears (h round-trips through HBM)
          auto uptr = [&](int id) { return h->d_w + h->off_h[fm::G_COUNT + m.L * fm::C_COUNT + upd * fm::U_COUNT + id]; };
          float* H = at<float>(ws, L.SA);
          const int gt = (int)(L.EPA / 128);
          if (img_on(h) && h->eu_fuse && h->vec_impl == 1) {   // EU1 -> SiLU -> EU2 -> residual + LayerNorm in one kernel (egemm_c.cuh)
            fm::EgArgs ac{uptr(tc_u(h, fm::U_EUPD_TC1)), uptr(fm::U_EUPD_B2), ef, nullptr, EAB, x, ef, uptr(fm::U_EUPD_LN_W), uptr(fm::U_EUPD_LN_B), L.EP, nullptr, 0, 0, h->tc_debug};
            ac.status = h->d_status;
            ac.in_img = at<float>(ws, L.EFI); ac.out_img = at<float>(ws, L.EFI);
            ac.g_units = uptr(tc_u(h, fm::U_EUPD_TC2));
            if (h->eu_quad && h->off_h[fm::G_COUNT + m.L * fm::C_COUNT + upd * fm::U_COUNT + fm::U_EUPD_TC1_HP] >= 0) {
              // four threads per row in both epilogues: permuted hidden / output features (egemm_c.cuh, QL)
              ac.units = uptr(fm::U_EUPD_TC1_HP); ac.g_units = uptr(fm::U_EUPD_TC2_HP);
              launch_k(h, fm::k_egemm_c<D, fm::CH_EU, 1>, gt < h->n_sm ? gt : h->n_sm, fm::EgcPlan::THREADS, fm::EgcPlan::SMEM_BYTES, st, m, bt, ac, gt);
            } else
            launch_k(h, fm::k_egemm_c<D, fm::CH_EU>, gt < h->n_sm ? gt : h->n_sm, fm::EgcPlan::THREADS, fm::EgcPlan::SMEM_BYTES, st, m, bt, ac, gt);
            LAUNCH_OK(h);
            done = true;
          }
          if (!done) {
          fm::EgArgs a1{uptr(tc_u(h, fm::U_EUPD_TC1)), nullptr, ef, nullptr, EAB, x, H, nullptr, nullptr, L.EP, h->trace_mode == 3 ? h->d_trace : nullptr, h->trace_cta, 0, h->tc_debug};
          a1.in_img = at<float>(ws, L.EFI); a1.out_img = H;
          if (img_on(h)) launch_eg<D, fm::EG_EU1, 1, fm::EGI_IN | fm::EGI_OUT>(h, gt, st, m, bt, a1);
          else launch_eg<D, fm::EG_EU1, 1>(h, gt, st, m, bt, a1);
          LAUNCH_OK(h);
          fm::EgArgs a2{uptr(tc_u(h, fm::U_EUPD_TC2)), uptr(fm::U_EUPD_B2), H, ef, nullptr, nullptr, ef, uptr(fm::U_EUPD_LN_W), uptr(fm::U_EUPD_LN_B), L.EP, h->trace_mode == 4 ? h->d_trace : nullptr, h->trace_cta, 0, h->tc_debug};
          a2.in_img = H; a2.out_img = at<float>(ws, L.EFI);
          if (img_on(h)) launch_eg<D, fm::EG_EU2, 1, fm::EGI_IN | fm::EGI_OUT>(h, gt, st, m, bt, a2);
          else launch_eg<D, fm::EG_EU2, 1>(h, gt, st, m, bt, a2);
          LAUNCH_OK(h);
          done = true;
          }
        }
      }
      if (!done) { launch_k(h, fm::k_edge_update<D>, L.nET, fm::NT, smem, st, m, bt, upd, x, EAB, ef); LAUNCH_OK(h); }
    }
    if (l == stop_after) return 0;
  }
  launch_k(h, fm::k_node_head<D>, L.nNT, fm::NT, smem, st, m, bt, s, out.a, out.c);
  LAUNCH_OK(h);
  bool head_done = false;
  if constexpr (D::S == 256 && D::V == 32 && D::SD == 0 && D::F == 128 && D::R == 32) {
    if (h->edge_reg && m.EB <= 8) {
      launch_k(h, fm::k_edge_head_r<D>, er_grid, fm::NT, fm::EdgeRegSmem<D>::HEAD_BYTES, st, m, bt, L.nUT * (fm::TM / fm::UR), ef, out.e);
      LAUNCH_OK(h);
      head_done = true;
    }
  }
  if (!head_done) {
    launch_k(h, fm::k_edge_head<D>, L.nUT, fm::NT, fm::EdgeSmem<D>::BYTES, st, m, bt, ef, out.e);
    LAUNCH_OK(h);
  }
  launch_k(h, fm::k_com, (L.B + 7) / 8, 256, 0, st, bt, x, out.x, remove_com);
  LAUNCH_OK(h);
  return 0;
}

template <class D>
int forward_impl(FmHandle* h, void* ws, const Layout& L, const float* x_t, const uint8_t* a_t, const uint8_t* c_t,
                 const uint8_t* e_t, float t, const fm::PredPtr* prev, const fm::PredPtr out, int stop_after, cudaStream_t st) {
  const fm::PredPtr none{nullptr, nullptr, nullptr, nullptr};
  fm::PredPtr pv = prev ? *prev : none;
  int has_prev = (h->cfg.self_conditioning && prev) ? 1 : 0;
  if (h->cfg.self_conditioning && !prev && t == 0.0f) {     // first sampling step: vector_field.py:269-282
    const fm::PredPtr tmp = pred_ptr(ws, L, 2);
    int rc = run_pass<D>(h, ws, L, x_t, a_t, c_t, e_t, t, none, 0, tmp, /*remove_com=*/0, -1, st);
    if (rc) return rc;
    pv = tmp;
    has_prev = 1;
  }
  return run_pass<D>(h, ws, L, x_t, a_t, c_t, e_t, t, pv, has_prev, out, /*remove_com=*/1, stop_after, st);
}

int dispatch_forward(FmHandle* h, void* ws, const Layout& L, const float* x_t, const uint8_t* a_t, const uint8_t* c_t,
                     const uint8_t* e_t, float t, const fm::PredPtr* prev, const fm::PredPtr out, int stop_after,
                     cudaStream_t st) {
  if (h->variant == 0) return forward_impl<fm::DimsFlowmol3>(h, ws, L, x_t, a_t, c_t, e_t, t, prev, out, stop_after, st);
  return forward_impl<fm::DimsDev>(h, ws, L, x_t, a_t, c_t, e_t, t, prev, out, stop_after, st);
}

// torch.linspace(0, 1, n) in fp32 (ATen RangeFactories: step = (end-start)/(steps-1); first half start + i*step,
// second half end - (steps-1-i)*step)
void time_grid(int n, float* t) {
  if (n == 1) { t[0] = 0.f; return; }
  const float step = (1.0f - 0.0f) / (float)(n - 1);
  const int half = n / 2;
  for (int i = 0; i < n; ++i) t[i] = i < half ? 0.0f + step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

// campbell_step's jump probabilities of one step in the reference's fp32 arithmetic (ctmc_vector_field.py:430-434), linear
// schedule: alpha_t = t, alpha_t' = 1 for every modality
void step_probs(fm::StepScalars& sc, float t_i, float dt, float eta) {
  sc.t_i = t_i;
  sc.dt = dt;
  for (int m = 0; m < 3; ++m) {
    volatile float num = 1.0f + eta * t_i;
    volatile float q = dt * num;
    volatile float den = 1.0f - t_i;
    sc.unmask_prob[m] = clamp01(q / den);
    volatile float qm = dt * eta;
    sc.mask_prob[m] = clamp01(qm);
  }
  sc.inj_u = nullptr;
  sc.inj_n = 0;
}

int find_batch(FmHandle* h, void* ws, const Layout** L) {
  auto it = h->batches.find(ws);
  if (it == h->batches.end()) return fail("workspace has no batch: call fm_batch_init first");
  *L = &it->second;
  return 0;
}

}  // namespace

extern "C" {

int fm_abi_version(void) { return FM_ABI_VERSION; }
const char* fm_last_error(void) { return g_err.c_str(); }

void fm_debug_time_grid(int n, float* out_host) { time_grid(n, out_host); }

int fm_create(const FmConfig* cfg, const float* w_host, size_t n_floats, const int64_t* off_host, size_t n_off, int device,
              FmHandle** out) {
  if (!cfg || !w_host || !off_host || !out) return fail("fm_create: null argument");
  const FmConfig& c = *cfg;
  int variant = -1;
  using F3 = fm::DimsFlowmol3;
  using DV = fm::DimsDev;
  auto matches = [&](int S, int V, int F, int R, int CP, int SD, int VD, int TOK, int TD) {
    return c.n_hidden_scalars == S && c.n_vec_channels == V && c.n_hidden_edge_feats == F && c.rbf_dim == R &&
           c.n_cp_feats == CP && (c.use_dst_feats ? c.s_dst : 0) == SD && (c.use_dst_feats ? c.v_dst : 0) == VD &&
           c.token_dim == TOK && c.time_embedding_dim == TD;
  };
  if (matches(F3::S, F3::V, F3::F, F3::R, F3::CP, F3::SD, F3::VD, F3::TOK, F3::TD)) variant = 0;
  else if (matches(DV::S, DV::V, DV::F, DV::R, DV::CP, DV::SD, DV::VD, DV::TOK, DV::TD)) variant = 1;
  if (variant < 0)
    return fail("fm_create: no kernel instantiation for these dimensions (built: flowmol3 = S256/V32/F128/R32/cp4, "
                "dev = S64/V16/F64/R32/cp4 + dst feats 16/4); add a Dims<> alias in csrc/model.cuh");
  if (c.n_atom_types + 1 > fm::AMAX || c.n_atom_types + c.n_charges > 32 || c.n_bond_types + 1 > fm::KMAXC ||
      c.n_charges + 1 > fm::KMAXC || c.n_atom_types + 1 > fm::KMAXC)
    return fail("fm_create: too many categories for the kernels' fixed-size buffers");
  const size_t expect = (size_t)fm::G_COUNT + (size_t)c.n_convs * fm::C_COUNT + (size_t)c.n_updaters * fm::U_COUNT;
  if (n_off != expect) return fail("fm_create: offset table has the wrong number of entries (weight_layout mismatch)");
  for (size_t i = 0; i < n_off; ++i)
    if (off_host[i] >= (int64_t)n_floats || (off_host[i] >= 0 && off_host[i] % 4 != 0)) return fail("fm_create: bad offset table");
  if (c.n_convs < 1 || c.convs_per_update < 1) return fail("fm_create: bad layer counts");
  CUDA_OK(cudaSetDevice(device));
  FmHandle* h = new FmHandle();
  h->cfg = c; h->device = device; h->variant = variant;
  h->has_tc = off_host[fm::G_COUNT + fm::C_MSG0_TCW] >= 0;
  h->off_h.assign(off_host, off_host + n_off);
  { cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, device) == cudaSuccess) h->n_sm = pr.multiProcessorCount; }
  h->conv_impl = (variant == 0 && h->has_tc) ? 2 : 0;   // flowmol3 dims: wide tcgen05 3xTF32 pipeline by default
  h->node_impl = (h->conv_impl == 2 && off_host[fm::G_COUNT + fm::C_UPD0_TCW] >= 0) ? 1 : 0;
  h->has_h16 = h->has_tc && off_host[fm::G_COUNT + fm::C_MSG0_TCW_H] >= 0 && off_host[fm::G_TC_INFO] >= 0;
  if (h->has_h16 && w_host[off_host[fm::G_TC_INFO]] != fm::tc::ACT_SCALE_H16) {
    delete h;
    return fail("fm_create: packed fp16 images were built for a different activation scale (weights.py vs csrc/tc.cuh)");
  }
  h->tc_prec = h->has_h16 ? 1 : 0;       // default: fp16x3 (same 22 significand bits as 3xTF32 at twice the MMA rate)
  if (variant == 0) {                    // k_node_embed on mma.sync: fixed weight scale 2^10, every |w| of its five matrices must stay small
    const int S = c.n_hidden_scalars, Ksc = ((S + c.n_atom_types + c.n_charges + c.rbf_dim + 3) / 4) * 4;
    const int64_t ids[5] = {fm::G_SEMB0_W, fm::G_SEMB2_W, fm::G_SCN0_W, fm::G_SCN2_W, fm::G_COUNT + fm::C_WSRC};
    const int64_t rows[5] = {2 * c.token_dim + c.time_embedding_dim, S, Ksc, S, S};
    bool ok = true;
    for (int i = 0; i < 5 && ok; ++i) {
      const int64_t o = off_host[ids[i]];
      if (o < 0) { ok = ok && (i == 2 || i == 3) && !c.self_conditioning; continue; }      // no self-conditioning: its MLP is absent
      if (o + rows[i] * S > (int64_t)n_floats) { ok = false; break; }
      for (int64_t k = 0; k < rows[i] * S; ++k) ok = ok && std::fabs(w_host[o + k]) < fm::NE_WMAX;
    }
    h->nemb_ok = ok ? 1 : 0;
  }
  h->eg_nh = 1;
  h->dyn = Dyn{c.n_hidden_scalars, c.n_vec_channels, c.n_hidden_edge_feats, c.use_dst_feats ? c.s_dst : 0,
               c.use_dst_feats ? c.v_dst : 0, c.n_atom_types, c.n_charges, c.n_bond_types,
               c.n_hidden_scalars + 3 * c.n_vec_channels};
  CUDA_OK(cudaMalloc(&h->d_w, n_floats * sizeof(float)));
  CUDA_OK(cudaMalloc(&h->d_off, n_off * sizeof(long long)));
  CUDA_OK(cudaMalloc(&h->d_table, sizeof(float) * 2 * (c.n_bond_types + 1) * c.n_hidden_edge_feats));   // edge-embedding table + its image under the SC residual's first linear
  CUDA_OK(cudaMalloc(&h->d_status, sizeof(int)));
  CUDA_OK(cudaMemset(h->d_status, 0, sizeof(int)));
  CUDA_OK(cudaMemcpy(h->d_w, w_host, n_floats * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->d_off, off_host, n_off * sizeof(long long), cudaMemcpyHostToDevice));
  fm::ModelRT& m = h->rt;
  m.w = h->d_w; m.off = h->d_off; m.eemb_table = h->d_table;
  m.A = c.n_atom_types; m.C = c.n_charges; m.EB = c.n_bond_types; m.L = c.n_convs; m.NU = c.n_updaters;
  m.convs_per_update = c.convs_per_update; m.separate_updaters = c.separate_mol_updaters;
  m.self_cond = c.self_conditioning; m.use_dst = c.use_dst_feats; m.rbf_dmax = c.rbf_dmax; m.msg_norm = c.message_norm;
  int rc = variant == 0 ? set_smem_attrs<F3>() : set_smem_attrs<DV>();
  if (rc) { fm_destroy(h); return rc; }
  if (variant == 0) fm::k_edge_table<F3><<<c.n_bond_types + 1, 128>>>(m, h->d_table);
  else fm::k_edge_table<DV><<<c.n_bond_types + 1, 128>>>(m, h->d_table);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  *out = h;
  return 0;
}

void fm_destroy(FmHandle* h) {
  if (!h) return;
  cudaFree(h->d_w); cudaFree(h->d_off); cudaFree(h->d_table); cudaFree(h->d_status);
  for (auto& pe : h->prof) cudaEventDestroy(pe.second);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  delete h;
}

int fm_workspace_bytes(FmHandle* h, const int32_t* n_atoms, int32_t B, size_t* bytes) {
  if (!h || !n_atoms || !bytes || B <= 0) return fail("fm_workspace_bytes: bad argument");
  for (int b = 0; b < B; ++b)
    if (n_atoms[b] < 2 || n_atoms[b] > 2000) return fail("fm_workspace_bytes: every molecule needs 2..2000 atoms");
  *bytes = make_layout(h->dyn, n_atoms, B).total;
  return 0;
}

int fm_batch_init(FmHandle* h, const int32_t* n_atoms, int32_t B, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !n_atoms || !ws || B <= 0) return fail("fm_batch_init: bad argument");
  for (int b = 0; b < B; ++b)
    if (n_atoms[b] < 2 || n_atoms[b] > 2000) return fail("fm_batch_init: every molecule needs 2..2000 atoms");
  const Layout L = make_layout(h->dyn, n_atoms, B);
  if (L.total > ws_bytes) return fail("fm_batch_init: workspace too small");
  if (reinterpret_cast<uintptr_t>(ws) % 256) return fail("fm_batch_init: workspace must be 256-byte aligned");
  if ((long long)L.EP * h->dyn.F >= (1ll << 40) || L.EP >= (1ll << 31)) return fail("fm_batch_init: batch too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int> mol_n(B), mol_node(B), mol_u(B), mol_et(B), mol_ut(B), et_mol(L.nET), ut_mol(L.nUT), node_mol(L.N);
  int node = 0, u = 0, et = 0, ut = 0;
  for (int b = 0; b < B; ++b) {
    const int n = n_atoms[b];
    mol_n[b] = n; mol_node[b] = node; mol_u[b] = u; mol_et[b] = et; mol_ut[b] = ut;
    for (int i = 0; i < n; ++i) node_mol[node + i] = b;
    const int ne = (n * (n - 1) + fm::TM - 1) / fm::TM, nu = (n * (n - 1) / 2 + fm::TM - 1) / fm::TM;
    for (int k = 0; k < ne; ++k) et_mol[et + k] = b;
    for (int k = 0; k < nu; ++k) ut_mol[ut + k] = b;
    node += n; u += n * (n - 1) / 2; et += ne; ut += nu;
  }
  CUDA_OK(cudaSetDevice(h->device));
  auto up = [&](size_t off, const std::vector<int>& v) {
    return cudaMemcpyAsync(at<char>(ws, off), v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st);
  };
  CUDA_OK(up(L.mol_n, mol_n)); CUDA_OK(up(L.mol_node, mol_node)); CUDA_OK(up(L.mol_u, mol_u));
  CUDA_OK(up(L.mol_etile, mol_et)); CUDA_OK(up(L.mol_utile, mol_ut)); CUDA_OK(up(L.etile_mol, et_mol));
  CUDA_OK(up(L.utile_mol, ut_mol)); CUDA_OK(up(L.node_mol, node_mol));
  // edge-feature rows and their operand images: the padding slots of every molecule's block are never written by k_edge_init_r
  // (it stores live rows only) and every linear computes whole 128-row tiles -- start them as zeros, not as allocator garbage
  CUDA_OK(cudaMemsetAsync(at<char>(ws, L.ef), 0, 4ull * (size_t)L.EPA * h->dyn.F, st));
  if (h->dyn.S == 256 && h->dyn.SD == 0) CUDA_OK(cudaMemsetAsync(at<char>(ws, L.EFI), 0, 4ull * (size_t)L.EPA * h->dyn.F, st));
  if (h->dyn.S == 256 && h->dyn.SD == 0) CUDA_OK(cudaMemsetAsync(at<char>(ws, L.SHI), 0, 256ull * (size_t)L.EPA, st));   // k columns >= 40 are never written
  CUDA_OK(cudaStreamSynchronize(st));      // the host vectors die here
  h->batches[ws] = L;
  return 0;
}

int fm_forward(FmHandle* h, void* ws, const float* x_t, const uint8_t* a_t, const uint8_t* c_t, const uint8_t* e_t, float t,
               const FmPred* prev, const FmPred* out, int32_t stop_after_conv, void* stream) {
  if (!h || !ws || !x_t || !a_t || !c_t || !e_t) return fail("fm_forward: null argument");
  if (stop_after_conv < 0 && !out) return fail("fm_forward: out is required for a full evaluation");
  const Layout* L;
  if (find_batch(h, ws, &L)) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  h->launches = 0;
  if (h->kprof) prof_mark(h, 0, static_cast<cudaStream_t>(stream));
  fm::PredPtr pv, po{nullptr, nullptr, nullptr, nullptr};
  if (prev) pv = fm::PredPtr{prev->x, prev->a, prev->c, prev->e};
  if (out) po = fm::PredPtr{out->x, out->a, out->c, out->e};
  return dispatch_forward(h, ws, *L, x_t, a_t, c_t, e_t, t, prev ? &pv : nullptr, po, stop_after_conv,
                          static_cast<cudaStream_t>(stream));
}

int fm_integrate(FmHandle* h, void* ws, float* x, uint8_t* a, uint8_t* c, uint8_t* e, const FmSampleOpts* o, void* stream) {
  return fm_integrate_traj(h, ws, x, a, c, e, o, nullptr, stream);
}

int fm_integrate_traj(FmHandle* h, void* ws, float* x, uint8_t* a, uint8_t* c, uint8_t* e, const FmSampleOpts* o,
                      const FmTraj* traj, void* stream) {
  if (!h || !ws || !x || !a || !c || !e || !o) return fail("fm_integrate: null argument");
  if (o->n_timesteps < 2) return fail("fm_integrate: n_timesteps must be >= 2");
  if (o->dfm_type != 0 && o->dfm_type != 1) return fail("fm_integrate: dfm_type must be 0 (campbell) or 1 (gat)");
  if (o->dfm_type == 1 && (!o->fw_host || !o->bw_host)) return fail("fm_integrate: dfm_type 1 (gat) needs fw_host and bw_host");
  const Layout* Lp;
  if (find_batch(h, ws, &Lp)) return -1;
  const Layout& L = *Lp;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->launches = 0;
  const int T = o->n_timesteps;
  std::vector<float> t(T);
  if (o->tspan_host) memcpy(t.data(), o->tspan_host, sizeof(float) * T);
  else time_grid(T, t.data());
  const fm::BatchRT bt = batch_rt(ws, L);
  cudaGraph_t graph = nullptr;
  cudaStream_t user_stream = st;
  if (o->use_cuda_graph) {              // whole-trajectory graph: every launch below is recorded, then replayed once on `stream`
    if (!h->cap_stream) CUDA_OK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    st = h->cap_stream;
    CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  }
  if (traj) {                                                     // frame 0 = the prior (ctmc_vector_field.py:187-202)
    if (traj->x) CUDA_OK(cudaMemcpyAsync(traj->x, x, sizeof(float) * 3 * L.N, cudaMemcpyDeviceToDevice, st));
    if (traj->a) CUDA_OK(cudaMemcpyAsync(traj->a, a, L.N, cudaMemcpyDeviceToDevice, st));
    if (traj->c) CUDA_OK(cudaMemcpyAsync(traj->c, c, L.N, cudaMemcpyDeviceToDevice, st));
    if (traj->e && L.U > 0) CUDA_OK(cudaMemcpyAsync(traj->e, e, L.U, cudaMemcpyDeviceToDevice, st));
  }
  int rc = 0;
  for (int k = 1; k < T && rc == 0; ++k) {                        // ctmc_vector_field.py:205-232
    const float t_i = t[k - 1], s_i = t[k];
    const fm::PredPtr cur = pred_ptr(ws, L, k & 1), prv = pred_ptr(ws, L, (k - 1) & 1);
    rc = dispatch_forward(h, ws, L, x, a, c, e, t_i, k == 1 ? nullptr : &prv, cur, -1, st);
    if (rc) break;
    fm::StepScalars sc;
    step_probs(sc, t_i, s_i - t_i, o->stochasticity);
    sc.hc_thresh = o->high_confidence_threshold;
    sc.tau = o->tau_host ? o->tau_host[k - 1] : o->cat_temperature;
    sc.dfm_type = o->dfm_type;
    sc.fw = o->fw_host ? o->fw_host[k - 1] : 1.0f;
    sc.bw = o->bw_host ? o->bw_host[k - 1] : 0.0f;
    sc.inv_temp = o->inv_temp_host ? o->inv_temp_host[k - 1] : 1.0f;
    sc.last_step = k == T - 1;
    sc.step_index = k;
    sc.seed_lo = (uint32_t)(o->seed & 0xffffffffull);
    sc.seed_hi = (uint32_t)(o->seed >> 32);
    sc.mol_id_offset = o->mol_id_offset;
    fm::TrajFrame tf{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (traj) {                                                   // frame k of the state, frame k - 1 of the endpoint predictions
      const size_t N = (size_t)L.N, U = (size_t)L.U;
      if (traj->x) tf.x = traj->x + (size_t)k * N * 3;
      if (traj->a) tf.a = traj->a + (size_t)k * N;
      if (traj->c) tf.c = traj->c + (size_t)k * N;
      if (traj->e) tf.e = traj->e + (size_t)k * U;
      if (traj->x1) tf.x1 = traj->x1 + (size_t)(k - 1) * N * 3;
      if (traj->a1) tf.a1 = traj->a1 + (size_t)(k - 1) * N;
      if (traj->c1) tf.c1 = traj->c1 + (size_t)(k - 1) * N;
      if (traj->e1) tf.e1 = traj->e1 + (size_t)(k - 1) * U;
    }
    launch_k(h, fm::k_ctmc_step, L.B, 256, 0, st, bt, h->rt.A, h->rt.C, h->rt.EB, cur.x, cur.a, cur.c, cur.e, x, a, c, e, sc, tf);
    ++h->launches;
    if (cudaGetLastError() != cudaSuccess) rc = fail("k_ctmc_step launch failed");
  }
  if (o->use_cuda_graph) {
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(std::string("stream capture failed: ") + cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
    cudaError_t le = cudaGraphLaunch(exec, user_stream);
    cudaGraphDestroy(graph);
    if (le != cudaSuccess) { cudaGraphExecDestroy(exec); return fail(std::string("graph launch failed: ") + cudaGetErrorString(le)); }
    CUDA_OK(cudaStreamSynchronize(user_stream));
    cudaGraphExecDestroy(exec);
  }
  return rc;
}

int fm_sample_host(FmHandle* h, const int32_t* n_atoms, int32_t B, float* x_host, uint8_t* a_host, uint8_t* c_host,
                   uint8_t* e_host, const FmSampleOpts* o, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !n_atoms || !x_host || !a_host || !c_host || !e_host || !o || !ws) return fail("fm_sample_host: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = fm_batch_init(h, n_atoms, B, ws, ws_bytes, stream);
  if (rc) return rc;
  const Layout& L = h->batches[ws];
  // the state lives in the (otherwise unused until the first ping-pong) tail of the workspace: reuse pred buffer 2's
  // neighbours is unsafe, so state gets its own small device allocation for the duration of the call
  float* dx = nullptr; uint8_t *da = nullptr, *dc = nullptr, *de = nullptr;
  CUDA_OK(cudaMallocAsync(&dx, sizeof(float) * 3 * L.N, st));
  CUDA_OK(cudaMallocAsync(&da, L.N, st));
  CUDA_OK(cudaMallocAsync(&dc, L.N, st));
  CUDA_OK(cudaMallocAsync(&de, L.U > 0 ? L.U : 1, st));
  CUDA_OK(cudaMemcpyAsync(dx, x_host, sizeof(float) * 3 * L.N, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(da, a_host, L.N, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dc, c_host, L.N, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(de, e_host, L.U, cudaMemcpyHostToDevice, st));
  rc = fm_integrate(h, ws, dx, da, dc, de, o, stream);
  if (rc == 0) {
    CUDA_OK(cudaMemcpyAsync(x_host, dx, sizeof(float) * 3 * L.N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(a_host, da, L.N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(c_host, dc, L.N, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(e_host, de, L.U, cudaMemcpyDeviceToHost, st));
  }
  int status = 0;
  if (rc == 0) CUDA_OK(cudaMemcpyAsync(&status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  cudaFreeAsync(dx, st); cudaFreeAsync(da, st); cudaFreeAsync(dc, st); cudaFreeAsync(de, st);
  CUDA_OK(cudaStreamSynchronize(st));
  if (status & 1) {
    cudaMemsetAsync(h->d_status, 0, sizeof(int), st);
    return fail("fm_sample_host: an activation left the fp16 operand range of the tensor-core linears; "
                "re-run with fm_set_option(h, \"tc_prec\", 0) (3xTF32)");
  }
  return rc;
}

int fm_decode(FmHandle* h, void* ws, const uint8_t* a, const uint8_t* c, const uint8_t* e, int32_t fake_atom_token,
              int32_t* atom_new, int8_t* charge, int32_t* mol_kept, int32_t* bond_src, int32_t* bond_dst, uint8_t* bond_type,
              int32_t* mol_bonds, void* stream) {
  if (!h || !ws || !a || !c || !e || !atom_new || !charge || !mol_kept || !bond_src || !bond_dst || !bond_type || !mol_bonds)
    return fail("fm_decode: null argument");
  const Layout* Lp;
  if (find_batch(h, ws, &Lp)) return -1;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const fm::BatchRT bt = batch_rt(ws, *Lp);
  launch_k(h, fm::k_decode, Lp->B, 256, 0, st, bt, a, c, e, fake_atom_token, h->cfg.n_bond_types, atom_new, charge, mol_kept, bond_src, bond_dst,
                                      bond_type, mol_bonds);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int fm_workspace_tensor(FmHandle* h, void* ws, const char* name, void** ptr, size_t* n_floats) {
  if (!h || !ws || !name || !ptr || !n_floats) return fail("fm_workspace_tensor: null argument");
  const Layout* Lp;
  if (find_batch(h, ws, &Lp)) return -1;
  const Layout& L = *Lp;
  const Dyn& d = h->dyn;
  const std::string n(name);
  size_t off, cnt;
  if (n == "s") { off = L.s; cnt = (size_t)L.N * d.S; }
  else if (n == "v") { off = L.v; cnt = (size_t)L.N * 3 * d.V; }
  else if (n == "x") { off = L.x; cnt = (size_t)L.N * 3; }
  else if (n == "P") { off = L.P; cnt = (size_t)L.N * d.S; }
  else if (n == "M") { off = L.M; cnt = (size_t)L.N * d.MW; }
  else if (n == "EAB") { off = L.EAB; cnt = (size_t)L.N * 2 * d.F; }
  else if (n == "ef") { off = L.ef; cnt = (size_t)L.EP * d.F; }
  else return fail("fm_workspace_tensor: unknown tensor name");
  *ptr = at<char>(ws, off);
  *n_floats = cnt;
  return 0;
}

int64_t fm_last_launch_count(FmHandle* h) { return h ? h->launches : -1; }

// Time the dominant kernel of the wide pipeline alone: the 292 -> 256 message linear (k_egemm_tc<EG_MSG>) of conv `layer`,
// GVP 1, on the buffers left by the last fm_forward (CUDA events on the launching stream).
int fm_time_egemm_msg(FmHandle* h, void* ws, int32_t layer, int32_t iters, float* ms_avg, void* stream) {
  if (!h || !ws || !ms_avg || iters < 1 || layer < 0 || layer >= h->cfg.n_convs) return fail("fm_time_egemm_msg: bad argument");
  if (h->variant != 0 || !h->has_tc) return fail("fm_time_egemm_msg: tensor-core pipeline not available for this model");
  const Layout* Lp;
  if (find_batch(h, ws, &Lp)) return -1;
  const Layout& L = *Lp;
  using D = fm::DimsFlowmol3;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const fm::BatchRT bt = batch_rt(ws, L);
  auto wptr = [&](int id) { return h->d_w + h->off_h[fm::G_COUNT + layer * fm::C_COUNT + id]; };
  fm::EgArgs a{wptr(tc_c(h, fm::C_MSG1_TCW)), wptr(fm::C_MSG1_WHCP + fm::GV_B), at<float>(ws, L.SA), at<float>(ws, L.SH), nullptr, nullptr,
               at<float>(ws, L.SB), nullptr, nullptr, L.EP, nullptr, 0, 0};
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  CUDA_OK(cudaStreamSynchronize(st));
  int status_before = 0;
  CUDA_OK(cudaMemcpy(&status_before, h->d_status, sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) {
    a.in_img = a.in_s; a.out_img = a.out;
    if (img_on(h)) launch_eg<D, fm::EG_MSG, 1, fm::EGI_IN | fm::EGI_OUT>(h, (int)(L.EPA / 128), st, h->rt, bt, a);
    else if (h->eg_nh == 2) launch_eg<D, fm::EG_MSG, 2>(h, (int)(L.EPA / 256), st, h->rt, bt, a);
    else launch_eg<D, fm::EG_MSG, 1>(h, (int)(L.EPA / 128), st, h->rt, bt, a);
  }
  CUDA_OK(cudaEventRecord(e1, st));
  CUDA_OK(cudaEventSynchronize(e1));
  CUDA_OK(cudaGetLastError());
  // the scratch buffers hold whatever the last phase left there (not this linear's real input): a range flag raised by this
  // timing run says nothing about the model
  CUDA_OK(cudaMemcpyAsync(h->d_status, &status_before, sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_avg = ms / (float)iters;
  return 0;
}

// One campbell_step of the production step kernel on crafted inputs (tests/golden/ctmc_cases.npz: h = 0, h = m, m = 0, last step,
// eta = 0): the atom-type modality of k_ctmc_step with K classes, the given sampling distribution p (not sharpened again) and
// injected uniforms instead of Philox.  Host buffers; the other two modalities run on one-class dummies.
int fm_debug_ctmc_step(const int32_t* n_atoms_host, int32_t B, int32_t K, const float* p_host, uint8_t* state_host, uint8_t* x1_host,
                       const float* uniforms_host, float t_i, float dt, float eta, float hc_thresh, int32_t last_step, int device) {
  if (!n_atoms_host || !p_host || !state_host || !x1_host || !uniforms_host || B < 1 || K < 1 || K > fm::KMAXC)
    return fail("fm_debug_ctmc_step: bad argument");
  CUDA_OK(cudaSetDevice(device));
  std::vector<int> mol_n(B), mol_node(B), mol_u(B);
  int N = 0, U = 0;
  for (int b = 0; b < B; ++b) {
    if (n_atoms_host[b] < 1) return fail("fm_debug_ctmc_step: empty molecule");
    mol_n[b] = n_atoms_host[b]; mol_node[b] = N; mol_u[b] = U;
    N += mol_n[b]; U += mol_n[b] * (mol_n[b] - 1) / 2;
  }
  int *d_n = nullptr, *d_node = nullptr, *d_u = nullptr;
  float *d_p = nullptr, *d_one = nullptr, *d_x = nullptr, *d_px = nullptr, *d_inj = nullptr;
  uint8_t *d_a = nullptr, *d_zero = nullptr, *d_x1 = nullptr;
  const size_t big = (size_t)(N > U ? N : U) + 1;
  CUDA_OK(cudaMalloc(&d_n, 4 * B)); CUDA_OK(cudaMalloc(&d_node, 4 * B)); CUDA_OK(cudaMalloc(&d_u, 4 * B));
  CUDA_OK(cudaMalloc(&d_p, sizeof(float) * N * K)); CUDA_OK(cudaMalloc(&d_one, sizeof(float) * big));
  CUDA_OK(cudaMalloc(&d_x, sizeof(float) * 3 * N)); CUDA_OK(cudaMalloc(&d_px, sizeof(float) * 3 * N)); CUDA_OK(cudaMalloc(&d_inj, sizeof(float) * 3 * N));
  CUDA_OK(cudaMalloc(&d_a, N)); CUDA_OK(cudaMalloc(&d_zero, big)); CUDA_OK(cudaMalloc(&d_x1, N));
  CUDA_OK(cudaMemcpy(d_n, mol_n.data(), 4 * B, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_node, mol_node.data(), 4 * B, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_u, mol_u.data(), 4 * B, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_p, p_host, sizeof(float) * N * K, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_inj, uniforms_host, sizeof(float) * 3 * N, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_a, state_host, N, cudaMemcpyHostToDevice));
  std::vector<float> ones(big, 1.0f);
  CUDA_OK(cudaMemcpy(d_one, ones.data(), sizeof(float) * big, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemset(d_zero, 0, big));
  CUDA_OK(cudaMemset(d_x, 0, sizeof(float) * 3 * N));
  CUDA_OK(cudaMemset(d_px, 0, sizeof(float) * 3 * N));
  fm::BatchRT bt{};
  bt.B = B; bt.N = N; bt.U = U;
  bt.mol_n = d_n; bt.mol_node = d_node; bt.mol_u = d_u;
  fm::StepScalars sc;
  step_probs(sc, t_i, dt, eta);
  sc.hc_thresh = hc_thresh;
  sc.tau = 0.f;                          // p is used as given
  sc.dfm_type = 0; sc.fw = 1.f; sc.bw = 0.f; sc.inv_temp = 1.f;
  sc.last_step = last_step; sc.step_index = 1; sc.seed_lo = 0; sc.seed_hi = 0; sc.mol_id_offset = 0;
  sc.inj_u = d_inj; sc.inj_n = N;
  fm::TrajFrame tf{nullptr, nullptr, nullptr, nullptr, nullptr, d_x1, nullptr, nullptr};
  // charges / bonds: one class with probability 1 on an unmasked state (a no-op for those modalities)
  fm::k_ctmc_step<<<B, 256>>>(bt, K, 1, 1, d_px, d_p, d_one, d_one, d_x, d_a, d_zero, d_zero, sc, tf);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(state_host, d_a, N, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(x1_host, d_x1, N, cudaMemcpyDeviceToHost));
  cudaFree(d_n); cudaFree(d_node); cudaFree(d_u); cudaFree(d_p); cudaFree(d_one); cudaFree(d_x); cudaFree(d_px); cudaFree(d_inj);
  cudaFree(d_a); cudaFree(d_zero); cudaFree(d_x1);
  return 0;
}

int fm_set_option(FmHandle* h, const char* name, int32_t value) {
  if (!h || !name) return fail("fm_set_option: null argument");
  const std::string n(name);
  if (n == "conv_impl") {
    if (value < 0 || value > 2) return fail("fm_set_option: conv_impl must be 0 (fp32 CUDA cores), 1 (fused tcgen05) or 2 (wide tcgen05 pipeline)");
    if (value >= 1 && h->variant != 0) return fail("fm_set_option: the tcgen05 kernel is built for the flowmol3 dimensions only");
    if (value >= 1 && !h->has_tc) return fail("fm_set_option: packed weights carry no tensor-core images");
    h->conv_impl = value;
    return 0;
  }
  if (n == "node_impl") {
    if (value < 0 || value > 1) return fail("fm_set_option: node_impl must be 0 (fused fp32 kernel) or 1 (tensor-core node pipeline)");
    if (value == 1 && (h->variant != 0 || h->off_h[fm::G_COUNT + fm::C_UPD0_TCW] < 0)) return fail("fm_set_option: no tensor-core images for the node pipeline");
    h->node_impl = value;
    return 0;
  }
  if (n == "fuse_agg") { h->fuse_agg = value ? 1 : 0; return 0; }
  if (n == "eg_persist") { h->eg_persist = value ? 1 : 0; return 0; }
  if (n == "eg_img") { h->eg_img = value ? 1 : 0; return 0; }
  if (n == "vec_impl") { h->vec_impl = value ? 1 : 0; return 0; }
  if (n == "eg_orient") { h->eg_orient = value ? 1 : 0; return 0; }
  if (n == "eg_fuse_gate") { h->eg_fuse_gate = value ? 1 : 0; return 0; }
  if (n == "eu_fuse") { h->eu_fuse = value ? 1 : 0; return 0; }
  if (n == "node_img") { h->node_img = value ? 1 : 0; return 0; }
  if (n == "edge_reg") { h->edge_reg = value ? 1 : 0; return 0; }
  if (n == "eg_pair") { h->eg_pair = value ? 1 : 0; return 0; }
  if (n == "pdl") { h->pdl = value ? 1 : 0; return 0; }
  if (n == "node_fuse_gate") { h->node_fuse_gate = value ? 1 : 0; return 0; }
  if (n == "sh_img") { h->sh_img = value ? 1 : 0; return 0; }
  if (n == "eg_epi12") { h->eg_epi12 = value ? 1 : 0; return 0; }
  if (n == "eg_perm") { h->eg_perm = value ? 1 : 0; return 0; }
  if (n == "eu_quad") { h->eu_quad = value ? 1 : 0; return 0; }
  if (n == "node_embed_tc") { h->nemb_tc = value ? 1 : 0; return 0; }
  if (n == "eg_cluster") {
    if (value != 1 && value != 2 && value != 4) return fail("fm_set_option: eg_cluster must be 1, 2 or 4");
    h->eg_cluster = value;
    return 0;
  }
  if (n == "tc_prec") {
    if (value < 0 || value > 1) return fail("fm_set_option: tc_prec must be 0 (3xTF32) or 1 (fp16x3)");
    if (value == 1 && !h->has_h16) return fail("fm_set_option: packed weights carry no fp16 operand images");
    h->tc_prec = value;
    return 0;
  }
  if (n == "kprof") {
    for (auto& pe : h->prof) cudaEventDestroy(pe.second);
    h->prof.clear();
    h->kprof = value ? 1 : 0;
    return 0;
  }
  if (n == "tc_debug") { h->tc_debug = value; return 0; }
  if (n == "tc_trace_mode") { h->trace_mode = value; return 0; }
  if (n == "tc_trace") {       // value < 0: off; otherwise the CTA index whose timeline is recorded (every egemm launch overwrites it)
    if (value < 0) { h->trace_cta = 0; if (h->d_trace) { cudaFree(h->d_trace); h->d_trace = nullptr; } return 0; }
    if (!h->d_trace) { CUDA_OK(cudaMalloc(&h->d_trace, 64 * sizeof(long long))); }
    CUDA_OK(cudaMemset(h->d_trace, 0, 64 * sizeof(long long)));
    h->trace_cta = value;
    return 0;
  }
  if (n == "eg_nh_gate") { if (value != 1 && value != 2) return fail("fm_set_option: eg_nh_gate must be 1 or 2"); h->eg_nh_gate = value; return 0; }
  if (n == "eg_nh") { if (value != 1 && value != 2) return fail("fm_set_option: eg_nh must be 1 or 2"); h->eg_nh = value; return 0; }
  return fail("fm_set_option: unknown option");
}
int fm_debug_read_trace(FmHandle* h, int64_t* out64_host) {
  if (!h || !out64_host || !h->d_trace) return fail("fm_debug_read_trace: tracing is off");
  CUDA_OK(cudaMemcpy(out64_host, h->d_trace, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}
// launches recorded since fm_set_option(h, "kprof", 1): api.cu line of each launch and its duration (ms, event to event)
int fm_debug_kprof(FmHandle* h, int32_t* lines_host, float* ms_host, int32_t cap, int32_t* n_out) {
  if (!h || !lines_host || !ms_host || !n_out) return fail("fm_debug_kprof: null argument");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaDeviceSynchronize());
  int n = 0;
  for (size_t i = 1; i < h->prof.size() && n < cap; ++i) {
    if (h->prof[i].first == 0) continue;                 // start marker of the next forward
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, h->prof[i - 1].second, h->prof[i].second));
    lines_host[n] = h->prof[i].first;
    ms_host[n] = ms;
    ++n;
  }
  *n_out = n;
  return 0;
}
int fm_get_option(FmHandle* h, const char* name, int32_t* value) {
  if (!h || !name || !value) return fail("fm_get_option: null argument");
  if (std::string(name) == "conv_impl") { *value = h->conv_impl; return 0; }
  if (std::string(name) == "node_impl") { *value = h->node_impl; return 0; }
  if (std::string(name) == "tc_prec") { *value = h->tc_prec; return 0; }
  if (std::string(name) == "eg_persist") { *value = h->eg_persist; return 0; }
  if (std::string(name) == "eg_img") { *value = h->eg_img; return 0; }
  if (std::string(name) == "vec_impl") { *value = h->vec_impl; return 0; }
  if (std::string(name) == "eg_orient") { *value = h->eg_orient; return 0; }
  if (std::string(name) == "eg_fuse_gate") { *value = h->eg_fuse_gate; return 0; }
  if (std::string(name) == "eu_fuse") { *value = h->eu_fuse; return 0; }
  if (std::string(name) == "node_img") { *value = h->node_img; return 0; }
  if (std::string(name) == "edge_reg") { *value = h->edge_reg; return 0; }
  if (std::string(name) == "eg_pair") { *value = h->eg_pair; return 0; }
  if (std::string(name) == "pdl") { *value = h->pdl; return 0; }
  if (std::string(name) == "node_fuse_gate") { *value = h->node_fuse_gate; return 0; }
  if (std::string(name) == "sh_img") { *value = h->sh_img; return 0; }
  if (std::string(name) == "eg_epi12") { *value = h->eg_epi12; return 0; }
  if (std::string(name) == "eg_perm") { *value = h->eg_perm; return 0; }
  if (std::string(name) == "eu_quad") { *value = h->eu_quad; return 0; }
  if (std::string(name) == "node_embed_tc") { *value = h->nemb_tc && h->nemb_ok; return 0; }
  if (std::string(name) == "eg_cluster") { *value = h->eg_cluster; return 0; }
  if (std::string(name) == "eg_clusters_seen") { *value = h->eg_clusters_seen; return 0; }
  if (std::string(name) == "status") {      // synchronising read-and-clear of the device status word (see fm_check_status)
    int v = 0;
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(&v, h->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemset(h->d_status, 0, sizeof(int)));
    *value = v;
    return 0;
  }
  return fail("fm_get_option: unknown option");
}

// stand-alone tcgen05 check (host buffers): out[128][64] = W[128][K] . X[64][K]^T, passes = 1 (plain TF32) or 3 (3xTF32)
int fm_debug_tc_gemm(const float* w_host, const float* x_host, int32_t K, float* out_host, int32_t passes, int device) {
  // passes: 1 plain TF32, 3 error-compensated 3xTF32, 16 scaled fp16 hi/lo ("fp16x3", K a multiple of 64)
  if (!w_host || !x_host || !out_host || K < 32 || K % 32 || K > 128 || (passes != 1 && passes != 3 && passes != 16) ||
      (passes == 16 && K % 64))
    return fail("fm_debug_tc_gemm: bad argument");
  CUDA_OK(cudaSetDevice(device));
  float *dw = nullptr, *dx = nullptr, *dout = nullptr;
  CUDA_OK(cudaMalloc(&dw, sizeof(float) * 128 * K));
  CUDA_OK(cudaMalloc(&dx, sizeof(float) * 64 * K));
  CUDA_OK(cudaMalloc(&dout, sizeof(float) * 128 * 64));
  CUDA_OK(cudaMemcpy(dw, w_host, sizeof(float) * 128 * K, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dx, x_host, sizeof(float) * 64 * K, cudaMemcpyHostToDevice));
  const int smem = (K / 32) * 48 * 1024 + 1024;
  CUDA_OK(cudaFuncSetAttribute(fm::k_tc_gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (passes == 16) {
    float wmax = 0.f;
    for (int i = 0; i < 128 * K; ++i) wmax = fmaxf(wmax, fabsf(w_host[i]));
    const float w_scale = wmax > 0.f ? exp2f(13.0f - floorf(log2f(wmax))) : 1.0f;      // weights.py:h16_weight_scale
    CUDA_OK(cudaFuncSetAttribute(fm::k_tc_gemm_test_h16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    fm::k_tc_gemm_test_h16<<<1, 128, smem>>>(dw, dx, K, dout, w_scale);
  } else {
    fm::k_tc_gemm_test<<<1, 128, smem>>>(dw, dx, K, dout, passes);
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(out_host, dout, sizeof(float) * 128 * 64, cudaMemcpyDeviceToHost));
  cudaFree(dw); cudaFree(dx); cudaFree(dout);
  return 0;
}

// probe of the tcgen05.ld shapes (tc_test.cuh:k_tmem_shape_probe): out_host int32 [128][16]
int fm_debug_tmem_shapes(int32_t* out_host, int device) {
  if (!out_host) return fail("fm_debug_tmem_shapes: bad argument");
  CUDA_OK(cudaSetDevice(device));
  int* d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(int) * 128 * 16));
  fm::k_tmem_shape_probe<<<1, 128>>>(d);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(out_host, d, sizeof(int) * 128 * 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

// Re-launch the hot kernel (k_conv_edge of `layer`) `iters` times on the workspace state left by the last fm_forward and
// time it with CUDA events on the launching stream (bench.py's roofline line).
int fm_time_conv_edge(FmHandle* h, void* ws, int32_t layer, int32_t iters, float* ms_avg, void* stream) {
  if (!h || !ws || !ms_avg || iters < 1 || layer < 0 || layer >= h->cfg.n_convs) return fail("fm_time_conv_edge: bad argument");
  const Layout* Lp;
  if (find_batch(h, ws, &Lp)) return -1;
  const Layout& L = *Lp;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const fm::BatchRT bt = batch_rt(ws, L);
  float *v = at<float>(ws, L.v), *x = at<float>(ws, L.x), *P = at<float>(ws, L.P), *Q = at<float>(ws, L.Q);
  float *vd = at<float>(ws, L.vd), *M = at<float>(ws, L.M), *partF = at<float>(ws, L.partF), *partL = at<float>(ws, L.partL);
  float* ef = at<float>(ws, L.ef);
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) {
    if (h->variant == 0 && h->conv_impl == 2) {
      int rc = conv_wide<fm::DimsFlowmol3>(h, ws, L, bt, layer, st);
      if (rc) return rc;
    } else if (h->variant == 0 && h->conv_impl == 1)
      launch_k(h, fm::k_conv_edge_tc<fm::DimsFlowmol3>, 2 * L.nET, fm::NT, fm::TcPlan<fm::DimsFlowmol3>::SMEM_BYTES, st, h->rt, bt, layer, x, v, ef, P, M, partF, partL, h->tc_debug);
    else if (h->variant == 0)
      launch_k(h, fm::k_conv_edge<fm::DimsFlowmol3>, L.nET, fm::NT, fm::DimsFlowmol3::SMEM_BYTES, st, h->rt, bt, layer, x, v, ef, P, Q, vd, M, partF, partL);
    else
      launch_k(h, fm::k_conv_edge<fm::DimsDev>, L.nET, fm::NT, fm::DimsDev::SMEM_BYTES, st, h->rt, bt, layer, x, v, ef, P, Q, vd, M, partF, partL);
  }
  CUDA_OK(cudaEventRecord(e1, st));
  CUDA_OK(cudaEventSynchronize(e1));
  CUDA_OK(cudaGetLastError());
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_avg = ms / (float)iters;
  return 0;
}

}  // extern "C"
