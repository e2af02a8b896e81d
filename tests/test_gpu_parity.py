"""GPU parity tests: CUDA path (through the C ABI) vs golden vectors produced by the reference's own code, and vs the
CPU oracle on fresh seeded inputs.  Run on the B200 box with `pytest -m gpu`.

Tolerances (the yardstick is the reference's own fp32-vs-fp64 noise floor, SURVEY.md section 8c: max|dx| 6e-7,
max|dp| 3e-7 per forward).  The flowmol3 configuration runs its message linears as error-compensated 3xTF32 on the
tensor cores by default (measured: hidden state 1.1e-5 abs / 2.5e-6 rel after 6 layers, outputs 2.4e-7; the fp32
CUDA-core kernels of the dev configuration and of `conv_impl=0` sit at 1.4e-6 / 1.8e-7):
    per forward : |dx| <= 5e-6, |dp| <= 3e-6, argmax(a, c, e) exact
    hidden state: rtol 1e-5 / atol 3e-5 per layer
    trajectory  : final categorical state identical for every molecule, |dx| <= 1e-4
"""
import numpy as np
import pytest
import torch

from flowmol_b200 import graph as G
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from oracle import flowmol_oracle as O
from tests.helpers import golden_schedules, load_golden, model_from_golden, t

pytestmark = pytest.mark.gpu

FWD = ["fwd_dev_taps", "fwd_flowmol3_taps", "fwd_flowmol3_geom", "fwd_dev_qm9"]
ITG = ["itg_dev_T10", "itg_dev_T50", "itg_flowmol3_T10", "itg_flowmol3_T25"]
TOL_X, TOL_P, TOL_H, ATOL_H = 5e-6, 3e-6, 1e-5, 3e-5

_models = {}


def cuda_model(cfg_name, A, wseed):
    from flowmol_b200.vector_field import CTMCVectorFieldB200
    key = (cfg_name, A, wseed)
    if key not in _models:
        cfg = ModelConfig.named(cfg_name, A)
        _models[key] = (cfg, CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, wseed), device="cuda:0"))
    return _models[key]


def model_for(gd):
    return cuda_model(str(gd["config"]), int(gd["n_atom_types"]), int(gd["weight_seed"]))


def check_pred(d, want, tag):
    d = {k: v.cpu().numpy() for k, v in d.items()}
    assert np.isfinite(d["x"]).all(), tag
    np.testing.assert_allclose(d["x"], want["x"], rtol=0, atol=TOL_X, err_msg=f"{tag} x")
    for k in "ace":
        np.testing.assert_allclose(d[k], want[k], rtol=0, atol=TOL_P, err_msg=f"{tag} {k}")
        assert np.array_equal(d[k].argmax(-1), want[k].argmax(-1)), f"{tag} argmax {k}"


@pytest.mark.parametrize("name", FWD)
def test_forward_vs_reference_golden(name):
    gd = load_golden(name)
    cfg, vf = model_for(gd)
    n_atoms = gd["n_atoms"]
    d0 = vf.forward_tokens(n_atoms, t(gd["c0.x_t"]), t(gd["c0.a"]), t(gd["c0.c"]), t(gd["c0.e"]), float(gd["c0.t"]), None)
    check_pred(d0, {k: gd[f"c0.out.{k}"] for k in "xace"}, f"{name} first step")
    prev = {k: t(gd[f"c0.out.{k}"]) for k in "xace"}
    d1 = vf.forward_tokens(n_atoms, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]), prev)
    check_pred(d1, {k: gd[f"c1.out.{k}"] for k in "xace"}, f"{name} mid trajectory")
    assert vf.last_launches > 0


@pytest.mark.parametrize("name", ["fwd_dev_taps", "fwd_flowmol3_taps"])
def test_layerwise_hidden_state_vs_reference_golden(name):
    gd = load_golden(name)
    cfg, vf = model_for(gd)
    n_atoms = gd["n_atoms"]
    perm = torch.from_numpy(G.ref_edge_to_internal(n_atoms))
    prev = {k: t(gd[f"c0.out.{k}"]) for k in "xace"}
    args = (n_atoms, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]), prev)
    N = int(np.sum(n_atoms))
    for l in range(cfg.n_convs):
        vf.forward_tokens(*args, stop_after_conv=l)
        s = vf.workspace_tensor("s").view(N, cfg.n_hidden_scalars).cpu().numpy()
        v = vf.workspace_tensor("v").view(N, 3, cfg.n_vec_channels).permute(0, 2, 1).cpu().numpy()
        np.testing.assert_allclose(s, gd[f"c1.tap.conv{l}.s"], rtol=TOL_H, atol=ATOL_H, err_msg=f"conv{l} s")
        np.testing.assert_allclose(v, gd[f"c1.tap.conv{l}.v"], rtol=TOL_H, atol=ATOL_H, err_msg=f"conv{l} v")
        if l >= 1:
            x = vf.workspace_tensor("x").view(N, 3).cpu().numpy()
            ef = vf.workspace_tensor("ef").view(-1, cfg.n_hidden_edge_feats).cpu()[perm].numpy()
            np.testing.assert_allclose(x, gd[f"c1.tap.pos{l}"], rtol=0, atol=TOL_X, err_msg=f"pos{l}")
            np.testing.assert_allclose(ef, gd[f"c1.tap.eupd{l}"], rtol=TOL_H, atol=ATOL_H, err_msg=f"eupd{l}")


@pytest.mark.parametrize("name", ITG)
def test_trajectory_vs_reference_golden(name):
    gd = load_golden(name)
    cfg, vf = model_for(gd)
    n_atoms = gd["n_atoms"]
    N, U = int(n_atoms.sum()), int((n_atoms * (n_atoms - 1) // 2).sum())
    A = cfg.n_atom_types
    out = vf.integrate_tokens(n_atoms, t(gd["x_0"]), torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4),
                              int(gd["T"]), seed=int(gd["noise_seed"]))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert np.array_equal(out["a"], gd["a_1"]), name
    assert np.array_equal(out["c"], gd["c_1"]), name
    assert np.array_equal(out["e"], gd["e_1"]), name
    np.testing.assert_allclose(out["x"], gd["x_1"], rtol=0, atol=1e-4)


def _oracle_vs_cuda_forward(cfg_name, A, n_atoms, wseed, seed):
    cfg, vf = cuda_model(cfg_name, A, wseed)
    om = O.OracleModel(cfg, WT.init_state_dict(cfg, wseed))
    bt = O.make_batch(n_atoms)
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(bt.N, 3, generator=gen)
    a = torch.randint(0, A + 1, (bt.N,), generator=gen)
    c = torch.randint(0, 7, (bt.N,), generator=gen)
    e = torch.randint(0, 5, (bt.U,), generator=gen)
    with torch.no_grad():
        w0 = om.forward(bt, x, torch.full_like(a, A), torch.full_like(c, 6), torch.full_like(e, 4), 0.0, None)
        w1 = om.forward(bt, x, a, c, e, 0.5, w0)
    d0 = vf.forward_tokens(n_atoms, x, torch.full_like(a, A), torch.full_like(c, 6), torch.full_like(e, 4), 0.0, None)
    check_pred(d0, {k: v.numpy() for k, v in w0.items()}, f"{cfg_name} {n_atoms} first")
    d1 = vf.forward_tokens(n_atoms, x, a, c, e, 0.5, w0)
    check_pred(d1, {k: v.numpy() for k, v in w1.items()}, f"{cfg_name} {n_atoms} mid")


def test_forward_edge_cases_vs_oracle_small_and_ragged():
    # n = 2 (one edge pair), n = 3, one 64-edge-exact molecule boundary (n = 9: 72 edges), ragged mix
    _oracle_vs_cuda_forward("dev", 6, [2, 3, 9, 2, 17, 5], wseed=21, seed=1)
    _oracle_vs_cuda_forward("flowmol3", 11, [2, 3, 9, 12], wseed=22, seed=2)


def test_forward_large_molecules_multi_tile_segments_vs_oracle():
    # n - 1 > 64: every dst's in-edge list spans 2..3 tiles (partial-sum path); 181 = GEOM maximum
    _oracle_vs_cuda_forward("dev", 6, [70, 130, 4], wseed=21, seed=3)
    _oracle_vs_cuda_forward("flowmol3", 11, [66, 181], wseed=22, seed=4)


def test_trajectory_vs_oracle_fresh_inputs():
    for cfg_name, A, n_atoms, T in (("dev", 6, [6, 2, 29, 18, 3], 30), ("flowmol3", 11, [5, 24, 9], 12)):
        cfg, vf = cuda_model(cfg_name, A, 31)
        om = O.OracleModel(cfg, WT.init_state_dict(cfg, 31))
        bt = O.make_batch(n_atoms)
        x0 = torch.randn(bt.N, 3, generator=torch.Generator().manual_seed(9))
        a0, c0, e0 = torch.full((bt.N,), A), torch.full((bt.N,), 6), torch.full((bt.U,), 4)
        with torch.no_grad():
            want = O.integrate(om, bt, x0, a0, c0, e0, T, seed=4242, mol_id_offset=100)
        got = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, T, seed=4242, mol_id_offset=100)
        for k in "ace":
            assert torch.equal(got[k].cpu().long(), want[k]), (cfg_name, k)
        assert (got["x"].cpu() - want["x"]).abs().max() <= 1e-4
        assert (got["a"] != A).all() and (got["c"] != 6).all() and (got["e"] != 4).all()


def test_trajectory_frames_vs_oracle_every_step():
    """fm_integrate_traj: the frames the step kernel writes (state after every step, predicted endpoint positions, endpoint
    tokens sampled by campbell_step) against the oracle's per-step record; the capture must not change the result."""
    cfg_name, A, n_atoms, T = "dev", 6, [6, 2, 17, 3], 16
    cfg, vf = cuda_model(cfg_name, A, 33)
    om = O.OracleModel(cfg, WT.init_state_dict(cfg, 33))
    bt = O.make_batch(n_atoms)
    x0 = torch.randn(bt.N, 3, generator=torch.Generator().manual_seed(11))
    a0, c0, e0 = torch.full((bt.N,), A), torch.full((bt.N,), 6), torch.full((bt.U,), 4)
    rec = []
    with torch.no_grad():
        want = O.integrate(om, bt, x0, a0, c0, e0, T, seed=77, record=rec)
    got = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, T, seed=77, traj=True)
    plain = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, T, seed=77)
    assert all(torch.equal(got[k], plain[k]) for k in "xace")
    fr = {k: v.cpu() for k, v in got["traj"].items()}
    assert fr["x"].shape == (T, bt.N, 3) and fr["e"].shape == (T, bt.U) and fr["x_1_pred"].shape == (T - 1, bt.N, 3)
    assert torch.equal(fr["x"][0], x0) and (fr["a"][0] == A).all() and (fr["e"][0] == 4).all()
    for k, r in enumerate(rec):
        for f in "ace":
            assert torch.equal(fr[f][k + 1].long(), r[f]), (k, f)
            assert torch.equal(fr[f + "_1_pred"][k].long(), r[f + "1"]), (k, f, "endpoint")
        assert (fr["x"][k + 1] - r["x"]).abs().max() <= 1e-4 and (fr["x_1_pred"][k] - r["x1"]).abs().max() <= 1e-4
    assert torch.equal(fr["x"][-1], got["x"].cpu()) and torch.equal(fr["a"][-1], got["a"].cpu())


def test_gat_sampler_and_schedules_vs_oracle():
    """dfm_type='gat' (ctmc_vector_field.py:463-510) with the reference's 'beta' forward-weight schedule, a decaying categorical
    temperature (:71-95) and an inverse-temperature factor on the positions (:334) -- the oracle side of this case is pinned
    bit-exactly to the verbatim reference in tests/test_oracle_vs_reference.py."""
    from flowmol_b200.vector_field import build_cat_temp_schedule, build_fw_schedule
    ctf, fwf, itf = build_cat_temp_schedule('decay', 0.8, 2), build_fw_schedule('beta', 0.25, 0.25, 10.0), (lambda t: 1.0 + 0.5 * t)
    for cfg_name, A, n_atoms, T in (("dev", 6, [6, 2, 17, 3], 14), ("flowmol3", 11, [5, 12], 8)):
        cfg, vf = cuda_model(cfg_name, A, 35)
        om = O.OracleModel(cfg, WT.init_state_dict(cfg, 35))
        bt = O.make_batch(n_atoms)
        x0 = torch.randn(bt.N, 3, generator=torch.Generator().manual_seed(12))
        a0, c0, e0 = torch.full((bt.N,), A), torch.full((bt.N,), 6), torch.full((bt.U,), 4)
        rec = []
        with torch.no_grad():
            want = O.integrate(om, bt, x0, a0, c0, e0, T, seed=99, record=rec, dfm_type='gat', cat_temp_func=ctf,
                               forward_weight_func=fwf, inv_temp_func=itf)
        got = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, T, seed=99, traj=True, dfm_type='gat', cat_temp_func=ctf,
                                  forward_weight_func=fwf, inv_temp_func=itf)
        fr = {k: v.cpu() for k, v in got["traj"].items()}
        for k, r in enumerate(rec):
            for f in "ace":
                assert torch.equal(fr[f][k + 1].long(), r[f]), (cfg_name, k, f)
                assert torch.equal(fr[f + "_1_pred"][k].long(), r[f + "1"]), (cfg_name, k, f, "endpoint")
            assert (fr["x"][k + 1] - r["x"]).abs().max() <= 1e-4
        for k in "ace":
            assert torch.equal(got[k].cpu().long(), want[k]), (cfg_name, k)
        assert (got["x"].cpu() - want["x"]).abs().max() <= 1e-4
        # the campbell sampler with a custom temperature schedule only
        with torch.no_grad():
            want = O.integrate(om, bt, x0, a0, c0, e0, T, seed=5, cat_temp_func=ctf)
        got = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, T, seed=5, cat_temp_func=ctf)
        for k in "ace":
            assert torch.equal(got[k].cpu().long(), want[k]), (cfg_name, k, "campbell + decay temperature")


def test_reference_api_trajectories_xt_ep():
    """model.sample(..., xt_traj=True, ep_traj=True): per-molecule frames in the reference's layout (ctmc_vector_field.py:268-283)
    and decoded frame molecules (molecule_builder.py:75-84,156-214)."""
    import flowmol_b200 as flowmol
    model = flowmol.FlowMolB200.from_config("dev", dataset="qm9", seed=3).cuda().eval()
    torch.manual_seed(1)
    n_atoms, T = torch.tensor([5, 9, 3]), 12
    mols = model.sample(n_atoms, n_timesteps=T, xt_traj=True, ep_traj=True)
    torch.manual_seed(1)
    plain = model.sample(n_atoms, n_timesteps=T)
    for m, q, n in zip(mols, plain, n_atoms.tolist()):
        assert torch.equal(m.positions, q.positions) and m.atom_types == q.atom_types       # capture changes nothing
        tf = m.traj_frames
        assert tf["x"].shape == (T, n, 3) and tf["a"].shape == (T, n, model.n_atom_types + 1) and tf["c"].shape == (T, n, 7)
        assert tf["e"].shape == (T, n * (n - 1), 5) and tf["x_1_pred"].shape == (T - 1, n, 3)
        assert tf["e_1_pred"].shape == (T - 1, n * (n - 1), 5)
        assert torch.equal(tf["e"][:, :n * (n - 1) // 2], tf["e"][:, n * (n - 1) // 2:])       # both triangles agree
        assert (tf["a"][0].argmax(-1) == model.n_atom_types).all()                             # frame 0 = the all-mask prior
        assert len(m.traj_mols) == T and len(m.ep_traj_mols) == T - 1
        assert m.traj_mols[-1].num_atoms == n                                                    # fake atoms are shown in frames
    only_xt = model.sample(n_atoms, n_timesteps=T, xt_traj=True)
    assert only_xt[0].traj_mols is not None and only_xt[0].ep_traj_mols is None


def test_results_do_not_depend_on_batch_composition_or_sharding():
    """Per-molecule Philox noise + molecule-aligned tiles => bit-identical molecules however the batch is cut."""
    cfg, vf = cuda_model("flowmol3", 11, 41)
    n_atoms = np.array([7, 30, 3, 70, 12, 19])
    bt = O.make_batch(n_atoms)
    x0 = torch.randn(bt.N, 3, generator=torch.Generator().manual_seed(3))
    A = 11
    full = vf.integrate_tokens(n_atoms, x0, torch.full((bt.N,), A), torch.full((bt.N,), 6), torch.full((bt.U,), 4), 8, seed=7)
    full = {k: v.cpu() for k, v in full.items()}
    noff = np.concatenate([[0], np.cumsum(n_atoms)])
    uoff = np.concatenate([[0], np.cumsum(n_atoms * (n_atoms - 1) // 2)])
    for lo, hi in ((0, 2), (2, 3), (3, 6)):          # three "ranks"
        sub = n_atoms[lo:hi]
        N, U = int(sub.sum()), int((sub * (sub - 1) // 2).sum())
        part = vf.integrate_tokens(sub, x0[noff[lo]:noff[hi]], torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4), 8,
                                   seed=7, mol_id_offset=lo)
        assert torch.equal(part["x"].cpu(), full["x"][noff[lo]:noff[hi]])
        assert torch.equal(part["a"].cpu(), full["a"][noff[lo]:noff[hi]])
        assert torch.equal(part["c"].cpu(), full["c"][noff[lo]:noff[hi]])
        assert torch.equal(part["e"].cpu(), full["e"][uoff[lo]:uoff[hi]])


def test_fp32_and_tensor_core_message_kernels_agree():
    """conv_impl 0 (fp32 CUDA cores), 1 (fused tcgen05) and 2 (wide tcgen05 pipeline, default) on the same inputs;
    with conv_impl 2 both node-update implementations (node_impl 1: pipeline around k_egemm_tc, default; 0: fused fp32 kernel)
    and both operand formats (tc_prec 1: scaled fp16 hi/lo, default; 0: 3xTF32 -- variant 5); variant 6 runs the fp16 operands
    through the one-tile-per-CTA kernel k_egemm_tc instead of the persistent k_egemm_p (same MMA order: bit-identical); variant 7
    is k_egemm_p with fp32 hand-over between the linears instead of the default fp16 (hi, lo) operand images (same split of the
    same values on the other side of HBM: bit-identical)."""
    cfg, vf = cuda_model("flowmol3", 11, 61)
    assert vf.get_option("conv_impl") == 2 and vf.get_option("node_impl") == 1 and vf.get_option("tc_prec") == 1
    n_atoms = np.array([9, 70, 3, 33])
    bt = O.make_batch(n_atoms)
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(bt.N, 3, generator=gen)
    a, c, e = torch.randint(0, 12, (bt.N,), generator=gen), torch.randint(0, 7, (bt.N,), generator=gen), torch.randint(0, 5, (bt.U,), generator=gen)
    outs = {}
    try:
        assert vf.get_option("eg_img") == 1
        for impl in (0, 1, 2, 3, 4, 5, 6, 7):
            vf.set_option("conv_impl", min(impl, 2))
            vf.set_option("tc_prec", 0 if impl == 5 else 1)
            vf.set_option("eg_persist", 0 if impl == 6 else 1)
            vf.set_option("eg_img", 0 if impl == 7 else 1)
            vf.set_option("node_impl", 0 if impl == 3 else 1)
            vf.set_option("fuse_agg", 0 if impl == 4 else 1)      # 4: scalar segment-sum in k_vec_c instead of the egemm epilogue
            d0 = vf.forward_tokens(n_atoms, x, torch.full_like(a, 11), torch.full_like(c, 6), torch.full_like(e, 4), 0.0, None)
            outs[impl] = {k: v.cpu() for k, v in vf.forward_tokens(n_atoms, x, a, c, e, 0.4, d0).items()}
    finally:
        vf.set_option("conv_impl", 2)
        vf.set_option("tc_prec", 1)
        vf.set_option("eg_persist", 1)
        vf.set_option("eg_img", 1)
        vf.set_option("node_impl", 1)
        vf.set_option("fuse_agg", 1)
    assert all(torch.equal(outs[4][k], outs[2][k]) for k in "xace")      # same sums in the same order: bit-identical
    assert all(torch.equal(outs[6][k], outs[2][k]) for k in "xace")
    assert all(torch.equal(outs[7][k], outs[2][k]) for k in "xace")
    for impl in (1, 2, 3, 5):
        assert (outs[impl]["x"] - outs[0]["x"]).abs().max() <= TOL_X
        for k in "ace":
            assert (outs[impl][k] - outs[0][k]).abs().max() <= TOL_P
            assert torch.equal(outs[impl][k].argmax(-1), outs[0][k].argmax(-1))


def test_host_entry_point_and_cuda_graph_match_device_entry_point():
    cfg, vf = cuda_model("dev", 6, 51)
    n_atoms = np.array([9, 4, 21])
    bt = O.make_batch(n_atoms)
    x0 = torch.randn(bt.N, 3, generator=torch.Generator().manual_seed(5))
    a0, c0, e0 = torch.full((bt.N,), 6), torch.full((bt.N,), 6), torch.full((bt.U,), 4)
    ref = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, 15, seed=11)
    gr = vf.integrate_tokens(n_atoms, x0, a0, c0, e0, 15, seed=11, cuda_graph=True)
    hx, ha, hc, he = x0.clone().pin_memory(), a0.to(torch.uint8).pin_memory(), c0.to(torch.uint8).pin_memory(), e0.to(torch.uint8).pin_memory()
    vf.sample_host(n_atoms, hx, ha, hc, he, 15, seed=11)
    for k, hv in zip("xace", (hx, ha, hc, he)):
        assert torch.equal(ref[k].cpu(), gr[k].cpu()), k
        assert torch.equal(ref[k].cpu(), hv), k


def test_reference_api_surface_sample_random_sizes():
    import flowmol_b200 as flowmol
    model = flowmol.FlowMolB200.from_config("dev", dataset="qm9", seed=3).cuda().eval()
    torch.manual_seed(0)
    mols = model.sample_random_sizes(6, n_timesteps=20)
    assert len(mols) == 6
    for m in mols:
        assert m.positions.shape[1] == 3 and len(m.atom_types) == m.num_atoms == m.positions.shape[0]
        assert (m.bond_src_idxs < m.bond_dst_idxs).all() and (m.bond_types > 0).all()
        assert "Se" not in m.atom_types                       # no masked atom survives t = 1
    torch.manual_seed(0)
    again = model.sample_random_sizes(6, n_timesteps=20)
    assert all(torch.equal(p.positions, q.positions) and p.atom_types == q.atom_types for p, q in zip(mols, again))


@pytest.mark.parametrize("cfg_name,dataset,B,T", [("flowmol3", "qm9", 1024, 250), ("flowmol3", "geom", 512, 20)])
def test_full_size_batches_size_independent_properties(cfg_name, dataset, B, T):
    """BASELINE configs 2 and 3 at full batch size: properties that do not need the (hours-long) CPU oracle."""
    from flowmol_b200.api import n_atoms_histogram
    A = 6 if dataset == "qm9" else 11
    cfg, vf = cuda_model(cfg_name, A, 0)
    nmap, counts = n_atoms_histogram(dataset)
    gen = torch.Generator().manual_seed(1234)
    n_atoms = nmap[torch.multinomial(counts / counts.sum(), B, replacement=True, generator=gen)].numpy()
    N, U = int(n_atoms.sum()), int((n_atoms * (n_atoms - 1) // 2).sum())
    x0 = torch.randn(N, 3, generator=gen)
    nbi = torch.arange(B).repeat_interleave(torch.from_numpy(n_atoms))
    x0 = x0 - (torch.zeros(B, 3).index_add_(0, nbi, x0) / torch.from_numpy(n_atoms)[:, None].float())[nbi]
    out = vf.integrate_tokens(n_atoms, x0, torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4), T, seed=99)
    x = out["x"].cpu()
    assert torch.isfinite(x).all()
    assert (out["a"] != A).all() and (out["c"] != 6).all() and (out["e"] != 4).all()      # every mask resolved at t = 1
    com = torch.zeros(B, 3).index_add_(0, nbi, x) / torch.from_numpy(n_atoms)[:, None].float()
    assert com.abs().max() < 1e-3                                                           # trajectories stay COM-free
    # re-running the first 8 molecules alone reproduces them bit for bit
    k = 8
    Nk, Uk = int(n_atoms[:k].sum()), int((n_atoms[:k] * (n_atoms[:k] - 1) // 2).sum())
    sub = vf.integrate_tokens(n_atoms[:k], x0[:Nk], torch.full((Nk,), A), torch.full((Nk,), 6), torch.full((Uk,), 4), T, seed=99)
    assert torch.equal(sub["x"].cpu(), x[:Nk]) and torch.equal(sub["a"].cpu(), out["a"].cpu()[:Nk])
    assert torch.equal(sub["e"].cpu(), out["e"].cpu()[:Uk])


def test_tcgen05_3xtf32_gemm_building_block():
    """tcgen05.mma kind::tf32 with hand-built SW128 K-major descriptors: 1xTF32 ~1e-3, 3xTF32 at fp32 accuracy."""
    import ctypes as C
    from flowmol_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for K in (32, 96, 128):
        W = rng.standard_normal((128, K)).astype(np.float32)
        X = rng.standard_normal((64, K)).astype(np.float32)
        want = W.astype(np.float64) @ X.astype(np.float64).T
        errs = {}
        for passes in (1, 3):
            out = np.zeros((128, 64), np.float32)
            _lib.check(lib.fm_debug_tc_gemm(W.ctypes.data, X.ctypes.data, K, out.ctypes.data, passes, 0))
            errs[passes] = float(np.abs(out - want).max() / np.abs(want).max())
        fp32 = float(np.abs((W @ X.T) - want).max() / np.abs(want).max())
        print(f"K={K} rel err 1xTF32 {errs[1]:.2e} 3xTF32 {errs[3]:.2e} fp32 {fp32:.2e}")
        assert errs[1] < 5e-3, errs
        assert errs[3] < 2e-6, errs


def test_tcgen05_fp16x3_gemm_building_block():
    """tcgen05.mma kind::f16 on scaled fp16 hi/lo operands (64 k values per SW128 row): fp32-level accuracy, also for inputs
    with a wide dynamic range (small values keep their low part thanks to the power-of-two operand scales)."""
    from flowmol_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for K in (64, 128):
        for kind in ("normal", "wide"):
            W = (rng.uniform(-1, 1, (128, K)) / np.sqrt(K)).astype(np.float32)
            X = rng.standard_normal((64, K)).astype(np.float32)
            if kind == "wide":
                X *= np.where(rng.random((64, K)) < 0.5, 1e-3, 30.0).astype(np.float32)
                W *= np.where(rng.random((128, K)) < 0.5, 1e-2, 1.0).astype(np.float32)
            want = W.astype(np.float64) @ X.astype(np.float64).T
            out = np.zeros((128, 64), np.float32)
            _lib.check(lib.fm_debug_tc_gemm(W.ctypes.data, X.ctypes.data, K, out.ctypes.data, 16, 0))
            out3 = np.zeros((128, 64), np.float32)
            _lib.check(lib.fm_debug_tc_gemm(W.ctypes.data, X.ctypes.data, K, out3.ctypes.data, 3, 0))
            err, err3 = (float(np.abs(o - want).max() / np.abs(want).max()) for o in (out, out3))
            print(f"K={K} {kind}: rel err fp16x3 {err:.2e} 3xTF32 {err3:.2e}")
            assert err < 2e-6, (K, kind, err)


def test_fp16_operand_overflow_is_reported_not_silent():
    """Activations beyond the fp16 operand range (|x| >= 65504) must raise, and the 3xTF32 operands must still work."""
    from flowmol_b200.vector_field import CTMCVectorFieldB200
    cfg = ModelConfig.named("flowmol3", 11)
    sd = WT.init_state_dict(cfg, 71)
    sd["scalar_embedding.4.weight"] = sd["scalar_embedding.4.weight"] * 1e5        # LayerNorm gain: node scalars ~1e5
    vf = CTMCVectorFieldB200(cfg, sd, device="cuda:0")
    n_atoms = np.array([5, 8])
    bt = O.make_batch(n_atoms)
    x = torch.randn(bt.N, 3, generator=torch.Generator().manual_seed(0))
    args = (n_atoms, x, torch.full((bt.N,), 11), torch.full((bt.N,), 6), torch.full((bt.U,), 4), 0.0, None)
    with pytest.raises(RuntimeError, match="fp16 operand range"):
        vf.forward_tokens(*args)
    vf.set_option("tc_prec", 0)
    d = vf.forward_tokens(*args)
    assert torch.isfinite(d["a"]).all()


def test_gat_trajectory_vs_reference_golden():
    """dfm_type='gat' + 'decay' temperature + 'beta' forward weights against the golden the verbatim reference produced
    (oracle/make_golden.py:gen_integrate_gat): final state, and molecule 0's state / recorded endpoint at every step."""
    gd = load_golden("itg_dev_gat_T12")
    cfg, vf = model_for(gd)
    n_atoms = gd["n_atoms"]
    N, U = int(n_atoms.sum()), int((n_atoms * (n_atoms - 1) // 2).sum())
    A = cfg.n_atom_types
    ctf, fwf = golden_schedules(gd)
    out = vf.integrate_tokens(n_atoms, t(gd["x_0"]), torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4),
                              int(gd["T"]), seed=int(gd["noise_seed"]), traj=True, dfm_type='gat', cat_temp_func=ctf,
                              forward_weight_func=fwf)
    assert np.array_equal(out["a"].cpu().numpy(), gd["a_1"])
    assert np.array_equal(out["c"].cpu().numpy(), gd["c_1"])
    assert np.array_equal(out["e"].cpu().numpy(), gd["e_1"])
    np.testing.assert_allclose(out["x"].cpu().numpy(), gd["x_1"], rtol=0, atol=1e-4)
    n0 = int(n_atoms[0])
    fr = {k: v.cpu().numpy() for k, v in out["traj"].items()}
    assert np.array_equal(fr["a"][:, :n0], gd["traj0.a"])
    assert np.array_equal(fr["a_1_pred"][:, :n0], gd["traj0.a_1_pred"])
    np.testing.assert_allclose(fr["x"][:, :n0], gd["traj0.x"], rtol=0, atol=1e-4)
