/* flowmol_b200 -- C ABI of the B200-native FlowMol sampling hot path.
 *
 * The reference (Dunni3/FlowMol) is pure Python and has no FFI; its seam for this path is an attribute call on an
 * nn.Module.  Each entry point below names the reference interface it replaces (paths relative to the reference root):
 *
 *   fm_create / fm_destroy   weights of `FlowMol.vector_field` (CTMCVectorField.__init__, flowmol/models/
 *                            ctmc_vector_field.py:23-69 + vector_field.py:16-197) as loaded by
 *                            FlowMol.load_from_checkpoint (flowmol/__init__.py:30-56)
 *   fm_batch_init            graph construction in FlowMol.sample (flowmol/models/flowmol.py:509-529:
 *                            build_edge_idxs / dgl.batch / get_upper_edge_mask / get_batch_idxs)
 *   fm_forward               CTMCVectorField.forward == EndpointVectorField.forward(g, t, node_batch_idx, upper_edge_mask,
 *                            apply_softmax=True, remove_com=True, prev_dst_dict)      (vector_field.py:212-293), the fine seam
 *                            at ctmc_vector_field.py:318
 *   fm_integrate             CTMCVectorField.integrate(g, node_batch_idx, upper_edge_mask, n_timesteps, stochasticity,
 *                            high_confidence_threshold, ...)  (ctmc_vector_field.py:145-285), the coarse seam at
 *                            flowmol/models/flowmol.py:557
 *   fm_integrate_traj        the same with visualize=True: per-step frames of the state and of the predicted / sampled endpoint
 *                            (ctmc_vector_field.py:187-202,235-283), written on the device by the step kernel
 *   fm_decode                the decode at the end of FlowMol.sample (flowmol.py:564-587, molecule_builder.py:217-265): argmax
 *                            state -> surviving atoms, charges, compact bond lists, on the device
 *   fm_sample_host           FlowMol.sample(n_atoms, n_timesteps, prior=...) from host buffers to host buffers
 *                            (flowmol.py:489-589 minus rdkit), i.e. integrate + the H2D / D2H around it
 *
 * Conventions
 *   - plain C types only; every array is caller-owned; device pointers unless the name ends in `_host`.
 *   - categorical state is carried as token indices (uint8): the argmax of the reference's one-hot tensors; the mask
 *     token is index n_classes (ctmc_vector_field.py:64-68).  Edge state / edge predictions are per UPPER-triangle
 *     edge in the reference's order (flowmol/data_processing/utils.py:4-17): molecule-major, (i<j) row-major.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*) except fm_create / fm_sample_host.
 *   - return 0 on success, negative on error; fm_last_error() gives the message.  No C++ exception crosses the ABI.
 *   - a handle is bound to one device; one in-flight call per workspace.
 */
#ifndef FLOWMOL_B200_H
#define FLOWMOL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_ABI_VERSION 2

typedef struct FmHandle FmHandle;

typedef struct FmConfig {
  int32_t n_atom_types;      /* incl. fake-atom type; mask token index == n_atom_types */
  int32_t n_charges, n_bond_types;
  int32_t n_hidden_scalars, n_vec_channels, n_hidden_edge_feats, n_cp_feats;
  int32_t rbf_dim, time_embedding_dim, token_dim;      /* a/c/e token dims must be equal */
  int32_t n_convs, n_updaters, convs_per_update, separate_mol_updaters;
  int32_t self_conditioning, use_dst_feats, s_dst, v_dst;
  float rbf_dmax;
  float message_norm;        /* 0: 'sum', -1: 'mean', >0: divide aggregated messages by this */
} FmConfig;

typedef struct FmPred {      /* predicted endpoint, the reference's dst_dict (softmaxed, COM-free) */
  float* x;                  /* [N,3] */
  float* a;                  /* [N,n_atom_types] */
  float* c;                  /* [N,n_charges] */
  float* e;                  /* [U,n_bond_types] upper edges */
} FmPred;

typedef struct FmSampleOpts {
  int32_t n_timesteps;
  float stochasticity;       /* eta */
  float high_confidence_threshold;
  float cat_temperature;     /* tau (0.05) */
  uint64_t seed;             /* Philox key */
  int32_t mol_id_offset;     /* global id of molecule 0 of this batch (sharding-invariant noise) */
  const float* tspan_host;   /* optional [n_timesteps] fp32 time grid; NULL => fp32 linspace(0,1,n), within 1 ulp of torch.linspace (pass torch's values for bit parity) */
  int32_t use_cuda_graph;    /* capture the per-step launch sequence once and replay it */
  int32_t dfm_type;          /* 0 = 'campbell' (mask / unmask jumps, ctmc_vector_field.py:414-461), 1 = 'gat' (:463-510) */
  const float* tau_host;     /* optional [n_timesteps-1]: categorical temperature of every step = cat_temp_func(t_i) (:71-81,353); NULL => cat_temperature */
  const float* fw_host;      /* dfm_type 1: [n_timesteps-1] forward weights forward_weight_func(t_i) (:83-95,385) */
  const float* bw_host;      /* dfm_type 1: [n_timesteps-1] backward weights fw - 1, evaluated by the caller in the reference's arithmetic (:491) */
  const float* inv_temp_host;/* optional [n_timesteps-1]: inv_temp_func(t_i) factor of the position update (:334); NULL => 1 */
} FmSampleOpts;

typedef struct FmTraj {      /* optional per-step frames, device memory owned by the caller; any pointer may be NULL.
                              * The reference collects these on the host every step (ctmc_vector_field.py:187-202,235-255); here
                              * the step kernel writes them as it goes, no extra launches or host round trips. */
  float* x;                  /* [T,N,3]  positions: frame 0 = x_0, frame k = x_t after step k */
  uint8_t* a;                /* [T,N]    atom-type tokens */
  uint8_t* c;                /* [T,N]    charge tokens */
  uint8_t* e;                /* [T,U]    bond tokens (upper edges) */
  float* x1;                 /* [T-1,N,3] predicted endpoint positions of every step (x_1_pred) */
  uint8_t* a1;               /* [T-1,N]   endpoint tokens sampled in campbell_step (a_1_pred) */
  uint8_t* c1;               /* [T-1,N] */
  uint8_t* e1;               /* [T-1,U] */
} FmTraj;

int fm_abi_version(void);
const char* fm_last_error(void);

/* weights: packed blob + offset table produced by flowmol_b200.weights.pack (layout: flowmol_b200/weight_layout.py) */
int fm_create(const FmConfig* cfg, const float* packed_weights_host, size_t n_floats, const int64_t* offsets_host,
              size_t n_offsets, int device, FmHandle** out);
void fm_destroy(FmHandle* h);

/* workspace sizing for a batch with the given atom counts (host array) */
int fm_workspace_bytes(FmHandle* h, const int32_t* n_atoms_host, int32_t n_molecules, size_t* bytes);
/* writes the batch descriptor into the workspace (must be 256-byte aligned device memory of fm_workspace_bytes bytes) */
int fm_batch_init(FmHandle* h, const int32_t* n_atoms_host, int32_t n_molecules, void* workspace, size_t workspace_bytes,
                  void* stream);

/* one network evaluation.  prev == NULL and t == 0 runs the self-conditioning pre-pass first (vector_field.py:269-282).
 * stop_after_conv >= 0 stops after that conv layer (+ its molecule update) for layer-wise parity tests; -1 = full. */
int fm_forward(FmHandle* h, void* workspace, const float* x_t, const uint8_t* a_t, const uint8_t* c_t,
               const uint8_t* e_t_upper, float t, const FmPred* prev, const FmPred* out, int32_t stop_after_conv,
               void* stream);

/* full trajectory: state arrays are updated in place from (x_0, a_0, c_0, e_0) to (x_1, a_1, c_1, e_1) */
int fm_integrate(FmHandle* h, void* workspace, float* x, uint8_t* a, uint8_t* c, uint8_t* e_upper,
                 const FmSampleOpts* opts, void* stream);

/* same with trajectory capture (`traj` may be NULL): replaces integrate(..., visualize=True), ctmc_vector_field.py:145-285 */
int fm_integrate_traj(FmHandle* h, void* workspace, float* x, uint8_t* a, uint8_t* c, uint8_t* e_upper,
                      const FmSampleOpts* opts, const FmTraj* traj, void* stream);

/* host-to-host convenience: H2D of the prior, fm_batch_init, fm_integrate, D2H of the result (synchronous) */
int fm_sample_host(FmHandle* h, const int32_t* n_atoms_host, int32_t n_molecules, float* x_host, uint8_t* a_host,
                   uint8_t* c_host, uint8_t* e_upper_host, const FmSampleOpts* opts, void* workspace,
                   size_t workspace_bytes, void* stream);

/* finalisation on the device: FlowMol.sample's decode (flowmol/models/flowmol.py:564-587 -> extract_moldata_from_graph,
 * flowmol/analysis/molecule_builder.py:217-265) of the batch last initialised on `workspace`, from final token state (device) to
 * compact per-molecule arrays (device): atom_new [N] = index of the atom after fake atoms (token == fake_atom_token; pass -1 for a
 * model without fake atoms) are dropped, or -1; charge [N] = charge token - 2; mol_kept [B] = surviving atoms; the bond list of
 * molecule b (upper-triangle edges in the reference's order whose order is neither 0 nor the mask token n_bond_types and whose atoms
 * both survive, renumbered src < dst) sits at bond_src / bond_dst / bond_type [mol_u[b] .. mol_u[b] + mol_bonds[b]) where
 * mol_u[b] = sum over earlier molecules of n(n-1)/2 (so every output array has the size of the corresponding input). */
int fm_decode(FmHandle* h, void* workspace, const uint8_t* a, const uint8_t* c, const uint8_t* e_upper, int32_t fake_atom_token,
              int32_t* atom_new, int8_t* charge, int32_t* mol_kept, int32_t* bond_src, int32_t* bond_dst, uint8_t* bond_type,
              int32_t* mol_bonds, void* stream);

/* introspection for tests: device pointer + size (floats) of a named workspace tensor after fm_batch_init:
 * "s" [N,S], "v" [N,3,V], "x" [N,3], "ef" [EP,F] (internal padded dst-major order), "P", "M" */
int fm_workspace_tensor(FmHandle* h, void* workspace, const char* name, void** ptr, size_t* n_floats);
/* host-only helper: the fallback time grid used when tspan_host == NULL (torch.linspace-like fp32 grid) */
void fm_debug_time_grid(int32_t n, float* out_host);
/* measurement helper: re-launch the hot kernel (fused gather + 3 message GVPs + segment-sum of conv `layer`) `iters` times on
 * the workspace state left by the last fm_forward and return its mean duration (CUDA events on `stream`). */
int fm_time_conv_edge(FmHandle* h, void* workspace, int32_t layer, int32_t iters, float* ms_avg, void* stream);
/* stand-alone check of the tcgen05 building blocks (host buffers): out[128][64] = W[128][K] . X[64][K]^T, K in {32,64,96,128};
 * passes = 1 (plain TF32) or 3 (error-compensated 3xTF32) */
int fm_debug_tc_gemm(const float* w_host, const float* x_host, int32_t K, float* out_host, int32_t passes, int device);
/* same, for the dominant kernel of the default flowmol3 pipeline: the 292 -> 256 message linear k_egemm_tc<EG_MSG> (GVP 1) */
/* probe of the tcgen05.ld shapes 16x256b / 16x128b / 16x64b: which (TMEM lane, column) every register of every thread receives
 * (out_host int32 [128][16]: cell = lane * 256 + column; [0..7] 16x256b.x2, [8..11] 16x128b.x2, [12..13] 16x64b.x2) */
int fm_debug_tmem_shapes(int32_t* out_host, int device);
int fm_time_egemm_msg(FmHandle* h, void* workspace, int32_t layer, int32_t iters, float* ms_avg, void* stream);
/* options: "conv_impl" = 0 fp32 CUDA-core message kernel (bit-for-bit the reference's fp32 arithmetic up to summation order),
 *          1 fused tcgen05 3xTF32 message kernel (experimental), 2 wide tcgen05 3xTF32 pipeline (default for the flowmol3
 *          dimensions; message linears + EdgeUpdate on the tensor cores); "eg_nh" / "eg_nh_gate" = 128-edge halves per CTA of
 *          k_egemm_tc (1: 2 CTAs/SM, 2: 1 CTA/SM); "tc_prec" = operand format of the wide pipeline's tensor-core linears:
 *          1 (default) scaled fp16 hi/lo images, three kind::f16 MMAs per product ("fp16x3": the 22 significand bits of 3xTF32 at
 *          twice the MMA rate and half the weight-image bytes; activations must stay below 65504 in magnitude), 0 3xTF32;
 *          "eg_persist" = 1 (default) persistent k_egemm_p / 0 one-tile-per-CTA k_egemm_tc; "eg_img" = 1 (default) consecutive
 *          tensor-core linears hand their activations over as fp16 (hi, lo) operand images fetched by bulk TMA / 0 as fp32 rows
 *          converted by the consumer's loader warps (bit-identical results); "vec_impl" = 1 (default) register-resident
 *          vector stages of the message GVPs (one warp per 16 edges, csrc/vec_reg.cuh) / 0 shared-memory tile kernels;
 *          "tc_debug", "tc_trace", "tc_trace_mode": timing experiments.
 *          "eg_fuse_gate" = 1 (default): the gate linear of message GVPs 1 / 2 runs inside the scalar linear's kernel (k_egemm_g, A
 *          operand in tensor memory); "eu_fuse" = 1 (default): EdgeUpdate (both linears, residual, LayerNorm) in one kernel
 *          (k_egemm_c); "node_img" = 1 (default): operand images between the node-row linears (bit-identical to 0);
 *          "edge_reg" = 1 (default): upper-edge MLPs as register-resident warp kernels (edge_reg.cuh) instead of the fp32 tile
 *          kernels; "eg_pair" = 1: gate-fused linears on CTA pairs (tcgen05 cta_group::2, bit-identical, measured slower:
 *          default 0); "eg_orient" = 1: edges-on-M k_egemm_e for MSG0 / MSG without gate fusion; "tc_debug": knock-out timing
 *          bits of k_egemm_g / k_egemm_c (results are garbage; tools/gpu_knockout.py); "pdl" = 1: programmatic dependent launch of
 *          every pipeline kernel (measured 3 % slower end to end: default 0); "node_fuse_gate" = 1 (default): scalar + gate linear of
 *          the node-row GVPs in one k_egemm_g launch; "sh_img" = 1 (default): vector norms of message GVPs 1 / 2 handed over as
 *          operand images (bit-identical); "eg_epi12" = 1 (default): k_egemm_h, the gate-fused edge-row linears with twelve epilogue
 *          warps (bit-identical to k_egemm_g); "eg_perm" = 1 (default): k_egemm_h<MSG> with permuted output features (image stores
 *          from tensor memory, 16 bytes per lane); "eu_quad" = 1 (default): k_egemm_c with four threads per row in both epilogues
 *          (tcgen05.ld 16x256b / st 16x128b, permuted hidden / output features; 0: one / two threads per row, agrees to 1e-6);
 *          "node_embed_tc" = 1 (default): the five 256-wide linears of k_node_embed on mma.sync fp16x3 instead of fp32 FFMA (active in
 *          the fp16x3 pipeline when every |w| of those matrices is below 32; fm_get_option reports whether it is in effect);
 *          fm_get_option(h, "status", &v) synchronises the device and reads-and-clears the status word: bit 0 = an activation
 *          left the fp16 operand range since the last read (results invalid; switch to tc_prec 0).  fm_sample_host checks it. */
int fm_set_option(FmHandle* h, const char* name, int32_t value);
int fm_get_option(FmHandle* h, const char* name, int32_t* value);
/* timeline experiments: after fm_set_option(h, "tc_trace", cta) every k_egemm_tc launch records clock64 stamps of that CTA */
int fm_debug_read_trace(FmHandle* h, int64_t* out64_host);
/* in-situ per-launch timing: after fm_set_option(h, "kprof", 1) every kernel launch of fm_forward is followed by a CUDA event;
 * returns, for up to `cap` launches, the csrc/api.cu line of the launch and its duration in ms (event to event, warm pipeline).
 * fm_set_option(h, "kprof", 0 or 1) clears the record. */
int fm_debug_kprof(FmHandle* h, int32_t* lines_host, float* ms_host, int32_t cap, int32_t* n_out);
/* test hook for CTMCVectorField.campbell_step + purity_sampling (ctmc_vector_field.py:414-461, flowmol/utils/ctmc_utils.py:4-35) as
 * the production step kernel k_ctmc_step computes them: one step of ONE modality with `K` classes (mask index K) on `B` molecules
 * with n_atoms_host[b] items each, the sampling distribution p_host [N][K] used as given (no temperature), uniforms_host [3][N] =
 * (categorical draw, unmask draw, re-mask draw) in place of the Philox noise, the linear schedule's jump probabilities at
 * (t_i, dt, eta).  state_host [N] is updated in place, x1_host [N] receives the sampled endpoint tokens.  Host buffers, synchronous. */
int fm_debug_ctmc_step(const int32_t* n_atoms_host, int32_t n_molecules, int32_t K, const float* p_host, uint8_t* state_host,
                       uint8_t* x1_host, const float* uniforms_host, float t_i, float dt, float eta, float hc_thresh,
                       int32_t last_step, int device);
/* number of kernels launched by the last fm_forward / fm_integrate call on this handle */
int64_t fm_last_launch_count(FmHandle* h);

#ifdef __cplusplus
}
#endif
#endif
