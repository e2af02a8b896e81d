// One integrator step after the network call (CTMCVectorField.step / campbell_step / purity_sampling:
// flowmol/models/ctmc_vector_field.py:328-409,414-461, flowmol/utils/ctmc_utils.py:4-35) -- one CTA per molecule.
//
// The reference does this with ~60 ATen launches per step, boolean-mask indexing (host syncs), torch_scatter
// segment_csr for the per-molecule purity counts and the global torch RNG.  Here: one launch per step, per-molecule
// counts by a block reduction, counter-based Philox noise keyed by (item, global molecule id, step, modality).
// Memory-bound: per upper edge 16 B (p_hat) + 1 B state read, 1 B written.
#pragma once
#include "model.cuh"

namespace fm {

constexpr int KMAXC = 16;   // max categories of one modality (the gat sampler adds the mask class: K + 1 <= KMAXC + 1)

struct StepScalars {
  float t_i, dt;
  float unmask_prob[3], mask_prob[3];    // per modality a, c, e   (clamped, ctmc_vector_field.py:430-434)
  float hc_thresh, tau;
  int last_step, step_index;
  uint32_t seed_lo, seed_hi;
  int mol_id_offset;
  int dfm_type;                          // 0 campbell, 1 gat (ctmc_vector_field.py:377-394,463-510)
  float fw, bw;                          // gat: forward weight of this step and the reference's `forward_weight - 1`
  float inv_temp;                        // factor of the position update (inv_temp_func(t_i), :334); 1 by default
  const float* inj_u;                    // test hook (fm_debug_ctmc_step): [3][N] uniforms of the atom-type modality replacing Philox
  int inj_n;                             //   (categorical draw, unmask draw, re-mask draw of node i at inj_u[k * inj_n + i]); NULL otherwise
};

// p = softmax(log(p_hat) / tau); returns max_k p_k ("purity")          (ctmc_vector_field.py:354-356)
__device__ __forceinline__ float sharpen(const float* __restrict__ phat, int K, float tau, float* p) {
  if (tau <= 0.f) {                      // test hook: p_hat already is the sampling distribution (campbell_step called directly)
    float pur = 0.f;
    for (int k = 0; k < K; ++k) { p[k] = phat[k]; pur = fmaxf(pur, p[k]); }
    return pur;
  }
  float mx = -INFINITY;
#pragma unroll 4
  for (int k = 0; k < K; ++k) { p[k] = __fdiv_rn(logf(phat[k]), tau); mx = fmaxf(mx, p[k]); }
  float sum = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) { p[k] = expf(__fsub_rn(p[k], mx)); sum = __fadd_rn(sum, p[k]); }
  float pur = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) { p[k] = __fdiv_rn(p[k], sum); pur = fmaxf(pur, p[k]); }
  return pur;
}

__device__ __forceinline__ int block_sum_int(int v, int* scratch /*NWARP ints*/) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}

// one modality of one molecule: `cnt` items, categories K (mask index == K)
// `frame` / `x1_frame` (optional): this step's trajectory frames of the new state and of the sampled endpoint x1
// (the reference's `<feat>_t` and `<feat>_1_pred` after the step, ctmc_vector_field.py:408-409 / 235-255)
__device__ __forceinline__ void campbell_modality(const float* __restrict__ phat /*[cnt][K]*/, uint8_t* __restrict__ state,
                                                  int cnt, int K, int modality, uint32_t mol_gid, const StepScalars& sc,
                                                  int* scratch, uint8_t* __restrict__ frame, uint8_t* __restrict__ x1_frame,
                                                  int item0 = 0) {
  const float q_u = sc.unmask_prob[modality], q_m = sc.mask_prob[modality];
  float p[KMAXC];
  if (sc.dfm_type == 1) {
    // gat_step (ctmc_vector_field.py:463-510), linear schedule alpha = t, alpha' = 1: Euler step of the probability velocity
    // fw * u_forward - (fw - 1) * u_backward over K + 1 classes (mask = K), clamp, inverse-CDF draw with the item's first uniform
    const float cf = __fdiv_rn(1.0f, __fsub_rn(1.0f, sc.t_i)), cb = __fdiv_rn(1.0f, __fadd_rn(sc.t_i, 1e-8f));
    for (int it = threadIdx.x; it < cnt; it += blockDim.x) {
      const int z = state[it];
      sharpen(phat + (size_t)it * K, K, sc.tau, p);
      const Philox4 rnd = philox4x32_10((uint32_t)it, mol_gid, (uint32_t)sc.step_index, (uint32_t)modality, sc.seed_lo, sc.seed_hi);
      float c = 0.f, cum[KMAXC + 1], best = -1.f;
      int amax = 0;
      for (int k = 0; k <= K; ++k) {
        const float p1 = k < K ? p[k] : 0.f, d = k == z ? 1.f : 0.f, dm = k == K ? 1.f : 0.f;
        if (k < K && p1 > best) { best = p1; amax = k; }
        const float uf = __fmul_rn(cf, __fsub_rn(p1, d)), ub = __fmul_rn(cb, __fsub_rn(d, dm));
        const float pvel = __fsub_rn(__fmul_rn(sc.fw, uf), __fmul_rn(sc.bw, ub));
        const float ps = fminf(fmaxf(__fadd_rn(d, __fmul_rn(sc.dt, pvel)), 1.0e-9f), 1.0f);
        c = __fadd_rn(c, ps);
        cum[k] = c;
      }
      const float thr = __fmul_rn(u24(rnd.x), c);
      int zn = 0;
      for (int k = 0; k <= K; ++k) zn += (cum[k] <= thr) ? 1 : 0;
      zn = min(zn, K);
      state[it] = (uint8_t)zn;
      if (frame) frame[it] = (uint8_t)zn;
      if (x1_frame) x1_frame[it] = (uint8_t)amax;         // the reference records p itself as the endpoint frame: its argmax
    }
    return;
  }
  // pass 1: per-molecule counts of masked / high-confidence masked items      (ctmc_utils.py:6-18)
  int m_loc = 0, h_loc = 0;
  if (sc.hc_thresh > 0.f) {
    for (int it = threadIdx.x; it < cnt; it += blockDim.x) {
      if (state[it] == K) {
        ++m_loc;
        if (sharpen(phat + (size_t)it * K, K, sc.tau, p) >= sc.hc_thresh) ++h_loc;
      }
    }
  }
  const int m_cnt = block_sum_int(m_loc, scratch);
  const int h_cnt = block_sum_int(h_loc, scratch);
  float ph = 0.f, pl = 0.f;
  if (sc.hc_thresh > 0.f) {
    const float qm = __fmul_rn(q_u, (float)m_cnt);
    const float ph_max = h_cnt == 0 ? INFINITY : __fdiv_rn(qm, (float)h_cnt);   // ctmc_utils.py:21-22
    ph = fminf(ph_max, 1.0f);
    pl = __fdiv_rn(__fsub_rn(qm, __fmul_rn(ph, (float)h_cnt)), (float)(m_cnt - h_cnt));   // ctmc_utils.py:26 (unused if m == h)
  }
  // pass 2: draw x1, decide unmask / re-mask                                      (ctmc_vector_field.py:428-457)
  for (int it = threadIdx.x; it < cnt; it += blockDim.x) {
    const int z = state[it];
    const float pur = sharpen(phat + (size_t)it * K, K, sc.tau, p);
    const Philox4 rnd = philox4x32_10((uint32_t)it, mol_gid, (uint32_t)sc.step_index, (uint32_t)modality, sc.seed_lo, sc.seed_hi);
    float u_cat = u24(rnd.x), u_unmask = u24(rnd.y), u_mask = u24(rnd.z);
    if (sc.inj_u && modality == 0) {
      u_cat = sc.inj_u[item0 + it]; u_unmask = sc.inj_u[sc.inj_n + item0 + it]; u_mask = sc.inj_u[2 * sc.inj_n + item0 + it];
    }
    // inverse-CDF categorical draw (shared definition with oracle/flowmol_oracle.py:sample_categorical)
    float c = 0.f, cum[KMAXC];
    for (int k = 0; k < K; ++k) { c = __fadd_rn(c, p[k]); cum[k] = c; }
    const float thr = __fmul_rn(u_cat, c);
    int x1 = 0;
    for (int k = 0; k < K; ++k) x1 += (cum[k] <= thr) ? 1 : 0;
    x1 = min(x1, K - 1);
    const bool masked = z == K;
    float prob;
    if (sc.hc_thresh > 0.f) prob = masked ? (pur >= sc.hc_thresh ? ph : pl) : 0.f;
    else prob = masked ? q_u : 0.f;
    const bool will_unmask = u_unmask < prob;
    int zn = z;
    if (!sc.last_step && (u_mask < q_m) && !masked) zn = K;
    if (will_unmask) zn = x1;
    state[it] = (uint8_t)zn;
    if (frame) frame[it] = (uint8_t)zn;
    if (x1_frame) x1_frame[it] = (uint8_t)x1;
  }
}

// per-step trajectory frames (device pointers of THIS step's frame, any may be null)
struct TrajFrame {
  float* x;                 // [N,3] positions after the step
  uint8_t *a, *c, *e;       // token state after the step ([N], [N], [U])
  float* x1;                // [N,3] predicted endpoint positions (x_1_pred)
  uint8_t *a1, *c1, *e1;    // sampled endpoint tokens (a/c/e_1_pred)
};

__global__ void __launch_bounds__(256)
k_ctmc_step(const BatchRT bt, int A, int C, int EB, const float* __restrict__ px, const float* __restrict__ pa,
            const float* __restrict__ pc, const float* __restrict__ pe, float* __restrict__ x_t,
            uint8_t* __restrict__ a_t, uint8_t* __restrict__ c_t, uint8_t* __restrict__ e_t, const StepScalars sc,
            const TrajFrame tf) {
  pdl_launch();
  pdl_wait();
  __shared__ int scratch[32];
  const int mol = blockIdx.x;
  const int n = bt.mol_n[mol], nb = bt.mol_node[mol], ub = bt.mol_u[mol], ucount = n * (n - 1) / 2;
  // positions: x_t += dt * (alpha'/(1-alpha)) * (x1_hat - x_t), linear schedule alpha = t  (ctmc_vector_field.py:331-334)
  const float coef = __fdiv_rn(1.0f, __fsub_rn(1.0f, sc.t_i));
  for (int i = threadIdx.x; i < n * 3; i += blockDim.x) {
    const int g = nb * 3 + i;
    const float vf = __fmul_rn(coef, __fsub_rn(px[g], x_t[g]));
    const float xn = __fadd_rn(x_t[g], __fmul_rn(__fmul_rn(sc.dt, vf), sc.inv_temp));
    x_t[g] = xn;
    if (tf.x) tf.x[g] = xn;
    if (tf.x1) tf.x1[g] = px[g];
  }
  const uint32_t gid = (uint32_t)(mol + sc.mol_id_offset);
  campbell_modality(pa + (size_t)nb * A, a_t + nb, n, A, 0, gid, sc, scratch, tf.a ? tf.a + nb : nullptr, tf.a1 ? tf.a1 + nb : nullptr, nb);
  campbell_modality(pc + (size_t)nb * C, c_t + nb, n, C, 1, gid, sc, scratch, tf.c ? tf.c + nb : nullptr, tf.c1 ? tf.c1 + nb : nullptr);
  campbell_modality(pe + (size_t)ub * EB, e_t + ub, ucount, EB, 2, gid, sc, scratch, tf.e ? tf.e + ub : nullptr, tf.e1 ? tf.e1 + ub : nullptr);
}

// ---- finalisation on the device (flowmol/models/flowmol.py:564-587 -> flowmol/analysis/molecule_builder.py:217-265) ------------------------
// extract_moldata_from_graph as one launch, one CTA per molecule: atoms whose type is the fake-atom token are dropped and the rest
// renumbered in order (atom_new[i] = new index or -1), charge = token - 2, and the bond list = upper-triangle edges in the reference's
// order whose order is neither 0 nor the mask token and whose two atoms survive, compacted per molecule at offset mol_u[b] with the
// renumbered (src < dst) indices.  The host then slices compact arrays instead of looping over one-hot graphs.
__device__ __forceinline__ int block_excl_scan_flag(bool flag, int* scratch /* >= 33 ints */, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  __syncthreads();                                     // previous use of scratch is over
  if (lane == 0) scratch[warp] = __popc(bal);
  __syncthreads();
  int before = 0, tot = 0;
  for (int w = 0; w < nw; ++w) { const int c = scratch[w]; if (w < warp) before += c; tot += c; }
  total = tot;
  return before + in_warp;
}

__global__ void __launch_bounds__(256)
k_decode(const BatchRT bt, const uint8_t* __restrict__ a, const uint8_t* __restrict__ c, const uint8_t* __restrict__ e, int fake_token,
         int mask_bond, int* __restrict__ atom_new, int8_t* __restrict__ charge, int* __restrict__ mol_kept,
         int* __restrict__ bond_src, int* __restrict__ bond_dst, uint8_t* __restrict__ bond_type, int* __restrict__ mol_bonds) {
  pdl_launch();
  pdl_wait();
  __shared__ int scratch[40];
  __shared__ int new_idx[2048];                        // n <= 2000 (fm_batch_init)
  const int mol = blockIdx.x, n = bt.mol_n[mol], nb = bt.mol_node[mol], ub = bt.mol_u[mol], ucount = n * (n - 1) / 2;
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool keep = i < n && (int)a[nb + i] != fake_token;
    int tot;
    const int pos = block_excl_scan_flag(keep, scratch, tot);
    if (i < n) {
      const int v = keep ? base + pos : -1;
      new_idx[i] = v;
      atom_new[nb + i] = v;
      charge[nb + i] = (int8_t)((int)c[nb + i] - 2);
    }
    base += tot;
  }
  if (threadIdx.x == 0) mol_kept[mol] = base;
  __syncthreads();
  int bbase = 0;
  for (int l0 = 0; l0 < ucount; l0 += blockDim.x) {
    const int lu = l0 + threadIdx.x;
    int i = 0, j = 0, btp = 0;
    bool sel = false;
    if (lu < ucount) {
      upper_ij(lu, n, i, j);
      btp = e[ub + lu];
      if (btp == mask_bond) btp = 0;
      sel = btp != 0 && new_idx[i] >= 0 && new_idx[j] >= 0;
    }
    int tot;
    const int pos = block_excl_scan_flag(sel, scratch, tot);
    if (sel) {
      bond_src[ub + bbase + pos] = new_idx[i];
      bond_dst[ub + bbase + pos] = new_idx[j];
      bond_type[ub + bbase + pos] = (uint8_t)btp;
    }
    bbase += tot;
  }
  if (threadIdx.x == 0) mol_bonds[mol] = bbase;
}

}  // namespace fm
