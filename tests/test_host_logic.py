"""CPU tests of the host-side logic: weight layout, C ABI surface, graph index algebra, decode, config validation."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from flowmol_b200 import graph as G
from flowmol_b200 import weight_layout as WL
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig, NAMED_VECTOR_FIELDS
from oracle import flowmol_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_weight_id_header_is_in_sync():
    with open(WL.HEADER_PATH) as f:
        assert f.read() == WL.header_text(), "run `python -m flowmol_b200.weight_layout --write`"


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from flowmol_b200 import _lib
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "flowmol_b200.h")) as f:
        declared = set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", f.read()))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fm_abi_version() == 2
    # the ctypes mirrors list the header's struct fields in the header's order
    hdr = open(os.path.join(ROOT, "include", "flowmol_b200.h")).read()
    for cname, cls in (("FmSampleOpts", _lib.FmSampleOpts), ("FmTraj", _lib.FmTraj), ("FmPred", _lib.FmPred)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [re.split(r"[\s\*]+", d.strip())[-1] for d in body.split(";") if d.strip()]
        assert names == [f[0] for f in cls._fields_], (cname, names)
    # host-only helper (no GPU needed): fallback time grid is within 1 ulp of torch.linspace
    for n in (2, 50, 250):
        out = np.zeros(n, np.float32)
        lib.fm_debug_time_grid(n, out.ctypes.data)
        ref = torch.linspace(0, 1, n).numpy()
        assert np.abs(out - ref).max() <= np.spacing(np.float32(1.0))


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from flowmol_b200.vector_field import CTMCVectorFieldB200
    cfg = ModelConfig.named("dev", 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0), device="cuda:0")
    with pytest.raises(RuntimeError):
        CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0), device="cpu")


@pytest.mark.parametrize("n_atoms", [[2], [3, 2, 9], [70, 5, 130]])
def test_graph_contract_and_internal_edge_order(n_atoms):
    g = G.MolGraphBatch(n_atoms)
    bt = O.make_batch(n_atoms)
    src, dst = g.edges()
    assert torch.equal(src, bt.src) and torch.equal(dst, bt.dst) and torch.equal(g.upper_edge_mask(), bt.upper)
    assert torch.equal(g.node_batch_idx(), bt.node_mol)
    perm = G.ref_edge_to_internal(n_atoms)
    sizes = G.batch_sizes(n_atoms)
    assert len(np.unique(perm)) == sizes["E"] and perm.max() < sizes["EP"]
    # internal order is dst-major inside each molecule: slots of one dst are contiguous and sorted by src
    ebase, nb = 0, 0
    for n in n_atoms:
        m = (bt.dst >= nb) & (bt.dst < nb + n)
        slots = perm[m.numpy()] - ebase
        d, s = (bt.dst[m] - nb).numpy(), (bt.src[m] - nb).numpy()
        order = np.argsort(slots)
        assert np.array_equal(slots[order], np.arange(n * (n - 1)))
        assert np.array_equal(d[order], np.repeat(np.arange(n), n - 1))
        assert all(np.array_equal(s[order][j * (n - 1):(j + 1) * (n - 1)], np.delete(np.arange(n), j)) for j in range(n))
        ebase += (n * (n - 1) + 63) // 64 * 64
        nb += n
    assert torch.equal(G.n_atoms_of(g), torch.tensor(n_atoms))


def test_segment_sum_scheme_of_conv_edge_kernel_emulated():
    """Emulates k_conv_edge's tile-local segment sums + k_node_update's in-order assembly (direct / partL / partF)."""
    rng = np.random.default_rng(0)
    for n in (2, 5, 33, 65, 66, 130, 181):
        ecount, TM = n * (n - 1), 64
        msg = rng.standard_normal(ecount)
        ntile = (ecount + TM - 1) // TM
        M, partF, partL = np.full(n, np.nan), np.full(ntile, np.nan), np.full(ntile, np.nan)
        for t in range(ntile):
            le0, acc, seg_first = t * TM, 0.0, t * TM
            for row in range(TM):
                le = le0 + row
                if le >= ecount:
                    break
                j = le // (n - 1)
                acc += msg[le]
                last = row == TM - 1 or le + 1 >= ecount or (le + 1) // (n - 1) != j
                if last:
                    head, tail = seg_first == j * (n - 1), le == j * (n - 1) + n - 2
                    if head and tail:
                        M[j] = acc
                    elif head:
                        partL[t] = acc
                    else:
                        partF[t] = acc
                    acc, seg_first = 0.0, le + 1
        for j in range(n):
            first, last = j * (n - 1), j * (n - 1) + n - 2
            t0, t1 = first // TM, last // TM
            got = M[j] if t0 == t1 else partL[t0] + sum(partF[t] for t in range(t0 + 1, t1 + 1))
            assert abs(got - msg[first:last + 1].sum()) < 1e-9, (n, j)


@pytest.mark.parametrize("name,A", [("dev", 6), ("flowmol3", 11)])
def test_packed_weights_node_side_folding_is_algebraically_exact(name, A):
    """P[src] (+ Q[dst]) + W_edge . [d | ef | sh] == W . cat[s_src, d, ef, (s_dst_msg), sh] + b   (fp64 check)."""
    cfg = ModelConfig.named(name, A)
    sd = WT.init_state_dict(cfg, 3)
    blob, off = WT.pack(cfg, sd)
    S, V, F, R, cp = cfg.n_hidden_scalars, cfg.n_vec_channels, cfg.n_hidden_edge_feats, cfg.rbf_dim, cfg.n_cp_feats
    sdst, h0 = cfg.s_dst, max(V + 1 + cfg.v_dst, V)
    pad4, pad32 = (lambda x: (x + 3) // 4 * 4), (lambda x: (x + 31) // 32 * 32)

    def mat(idx, K, N):
        o = off[idx]
        return blob[o:o + pad4(K) * pad32(N)].reshape(pad4(K), pad32(N))[:K, :N].astype(np.float64)
    rng = np.random.default_rng(1)
    s_src, d, ef, sd_, sh = (rng.standard_normal(k) for k in (S, R, F, sdst, h0 + cp))
    W = sd["conv_layers.1.edge_message.0.to_feats_out.0.weight"].double().numpy()
    b = sd["conv_layers.1.edge_message.0.to_feats_out.0.bias"].double().numpy()
    want = W @ np.concatenate([s_src, d, ef, sd_, sh]) + b
    P = s_src @ mat(WL.cid(1, "WSRC"), S, S) + blob[off[WL.cid(1, "BSRC")]:][:S]
    got = P + np.concatenate([d, ef, sh]) @ mat(WL.cid(1, "MSG0_W"), R + F + h0 + cp, S)
    if sdst:
        got = got + sd_ @ mat(WL.cid(1, "WDST"), sdst, S)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    # EdgeUpdate: EA[src] + EB[dst] + W_e . [ef | d]
    W1 = sd["edge_updaters.1.edge_update_fn.0.weight"].double().numpy()
    b1 = sd["edge_updaters.1.edge_update_fn.0.bias"].double().numpy()
    s_dst = rng.standard_normal(S)
    want = W1 @ np.concatenate([s_src, s_dst, ef, d]) + b1
    WN = mat(WL.uid(cfg.n_convs, 1, "EUPD_WN"), S, 2 * F)
    BN = blob[off[WL.uid(cfg.n_convs, 1, "EUPD_BN")]:][:2 * F]
    got = (s_src @ WN + BN)[:F] + (s_dst @ WN + BN)[F:] + np.concatenate([ef, d]) @ mat(WL.uid(cfg.n_convs, 1, "EUPD_WE"), F + R, F)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


def test_config_rejects_options_off_the_live_path():
    base = dict(NAMED_VECTOR_FIELDS["flowmol3"])
    for bad in (dict(attention=True), dict(s_message_dim=64), dict(n_recycles=2), dict(a_token_dim=0), dict(n_message_gvps=2)):
        with pytest.raises((NotImplementedError, ValueError)):
            ModelConfig.from_vector_field_block({**base, **bad}, 11)
    cfg = ModelConfig.named("flowmol3", 11)
    assert (cfg.n_convs, cfg.n_updaters, cfg.s_dst) == (6, 6, 0)
    assert sum(int(np.prod(s)) for _, s in WT.expected_tensors(cfg)) == 5854185          # BASELINE.md parameter count


def test_lightning_checkpoint_ingestion_and_molecule_decode(tmp_path):
    from flowmol_b200.api import FlowMolB200, SampledMolecule, load_pretrained
    cfg = ModelConfig.named("dev", 11)
    sd = WT.init_state_dict(cfg, 0)
    ck = {"state_dict": {"vector_field." + k: v for k, v in sd.items()} | {"loss_fn.weight": torch.zeros(3)},
          "hyper_parameters": {"atom_type_map": ['C', 'H', 'N', 'O', 'F', 'P', 'S', 'Cl', 'Br', 'I'], "fake_atom_p": 0.3,
                               "vector_field_config": NAMED_VECTOR_FIELDS["dev"], "default_n_timesteps": 250,
                               "explicit_aromaticity": False, "n_atoms_hist_file": "data/geom_full_kekulized/x.pt"}}
    d = tmp_path / "flowmol3" / "checkpoints"
    d.mkdir(parents=True)
    torch.save(ck, d / "last.ckpt")
    model = load_pretrained("flowmol3", models_dir=tmp_path) if not torch.cuda.is_available() else FlowMolB200.from_checkpoint(d / "last.ckpt")
    assert model.n_atom_types == 11 and model.cfg.n_hidden_scalars == 64
    assert all(torch.equal(model._state_dict[k], sd[k]) for k in sd)
    sizes = model.sample_n_atoms(1000)
    assert sizes.min() >= 3 and sizes.max() <= 181
    with pytest.raises(ValueError):
        load_pretrained("nope")
    # options the kernels do not implement must be refused at load time, not sampled with the wrong coefficients
    for bad in ({"parameterization": "dirichlet"}, {"n_atom_charges": 5}, {"exclude_charges": True},
                {"interpolant_scheduler_config": {"schedule_type": {"x": "cosine", "a": "linear", "c": "linear", "e": "linear"}}}):
        ck2 = dict(ck, hyper_parameters={**ck["hyper_parameters"], **bad})
        torch.save(ck2, d / "bad.ckpt")
        with pytest.raises(NotImplementedError):
            FlowMolB200.from_checkpoint(d / "bad.ckpt", device="cpu")
    ok = dict(ck, hyper_parameters={**ck["hyper_parameters"], "parameterization": "ctmc", "n_atom_charges": 6,
                                    "interpolant_scheduler_config": {"schedule_type": {k: "linear" for k in "xace"}}})
    torch.save(ok, d / "ok.ckpt")
    assert FlowMolB200.from_checkpoint(d / "ok.ckpt", device="cpu").n_atom_types == 11
    # a checkpoint that pickles a class is not read by the safe loader unless the caller opts in
    import argparse
    torch.save(dict(ck, callbacks={"x": argparse.Namespace(a=1)}), d / "cls.ckpt")
    with pytest.raises(RuntimeError, match="weights_only"):
        WT.state_dict_from_checkpoint(d / "cls.ckpt")
    sd2, _ = WT.state_dict_from_checkpoint(d / "cls.ckpt", allow_unsafe_pickle=True)
    assert all(torch.equal(sd2[k], sd[k]) for k in sd)
    # decode: 4 atoms, atom 2 is a fake atom (index 10), bonds 0-1 single, 0-2 (to the fake atom) double, 1-3 masked
    x = np.arange(12, dtype=np.float32).reshape(4, 3)
    a = np.array([0, 3, 10, 1])
    c = np.array([2, 3, 2, 1])
    e = np.array([1, 2, 0, 0, 4, 3])            # upper edges (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    m = SampledMolecule(x, a, c, e, model.atom_type_map, fake_atoms=True)
    assert m.atom_types == ['C', 'O', 'H'] and m.atom_charges.tolist() == [0, 1, -1] and m.num_atoms == 3
    assert m.bond_types.tolist() == [1] and m.bond_src_idxs.tolist() == [0] and m.bond_dst_idxs.tolist() == [1]
    assert torch.equal(m.positions, torch.from_numpy(x[[0, 1, 3]]))


def test_cli_batch_planner_is_cost_aware_and_complete():
    """flowmol_b200/cli.py:plan_batches (the batch loop of the reference's test.py:99-133, made size-aware)."""
    from flowmol_b200.cli import plan_batches
    rng = np.random.default_rng(0)
    n = rng.integers(3, 182, size=500)
    for mbs, mbe in ((128, 4_000_000), (32, 60_000), (500, 10 ** 9), (7, 100)):
        batches = plan_batches(n, mbs, mbe)
        flat = np.concatenate(batches)
        assert sorted(flat.tolist()) == list(range(len(n)))                      # every molecule exactly once
        for b in batches:
            edges = int((n[b] * (n[b] - 1)).sum())
            assert len(b) <= mbs and (edges <= mbe or len(b) == 1)                 # an oversized molecule gets its own batch
        sizes = [int(n[b].max()) for b in batches]
        assert sizes == sorted(sizes, reverse=True)                                # large molecules first, similar sizes together
    assert len(plan_batches(n, 500, 10 ** 9)) == 1


def test_cli_sdf_writer_and_trajectory_frame_decode(tmp_path):
    """V2000 mol blocks from the decoded arrays (no rdkit needed), the reference's trajectory-frame layout decoded per frame
    (molecule_builder.py:156-214) and its Kabsch alignment (priors.py:128-169)."""
    from flowmol_b200.api import SampledMolecule, rigid_alignment
    from flowmol_b200.cli import mol_block, write_sdf
    amap = ['C', 'H', 'N', 'O', 'F']
    x = np.array([[0.0, 0.0, 0.0], [1.1, 0.0, 0.0], [0.0, 1.2, 0.0], [5.0, 5.0, 5.0]], np.float32)
    a = np.array([0, 3, 2, 5])                 # last atom: the fake-atom token -> dropped
    c = np.array([2, 1, 3, 2])                 # charges 0, -1, +1
    e = np.array([2, 1, 0, 0, 4, 0])           # upper edges (0,1) double, (0,2) single, (1,2) still masked -> no bond
    m = SampledMolecule(x, a, c, e, amap, fake_atoms=True)
    assert m.atom_types == ['C', 'O', 'N'] and m.atom_charges.tolist() == [0, -1, 1]
    assert m.bond_types.tolist() == [2, 1] and m.bond_src_idxs.tolist() == [0, 0] and m.bond_dst_idxs.tolist() == [1, 2]
    blk = mol_block(m, "t").split("\n")
    assert blk[3].startswith("  3  2") and blk[3].endswith("V2000")
    assert blk[4].split()[3] == 'C' and blk[7].split() == ['1', '2', '2', '0'] and blk[8].split() == ['1', '3', '1', '0']
    assert blk[9].split() == ['M', 'CHG', '2', '2', '-1', '3', '1'] and blk[-1] == "M  END"
    write_sdf(tmp_path / "o.sdf", [m, m])
    assert (tmp_path / "o.sdf").read_text().count("$$$$") == 2
    # alignment: a rotated + shifted copy comes back onto the target
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    y = x @ R.T + np.array([1.0, -2.0, 0.5])
    assert np.abs(rigid_alignment(y, x) - x).max() < 1e-5
    # trajectory frames in the reference layout: one-hot floats, edges in both directions
    T, n = 3, 4
    oh = lambda idx, k: torch.nn.functional.one_hot(torch.as_tensor(idx), k).float()
    frames = {'x': torch.from_numpy(np.stack([x + k for k in range(T)])), 'a': torch.stack([oh(a, 7)] * T),
              'c': torch.stack([oh(c, 7)] * T), 'e': torch.stack([torch.cat([oh(e, 5), oh(e, 5)])] * T),
              'x_1_pred': torch.from_numpy(np.stack([x] * (T - 1))), 'a_1_pred': torch.stack([oh(a, 7)] * (T - 1)),
              'c_1_pred': torch.stack([oh(c, 7)] * (T - 1)), 'e_1_pred': torch.stack([torch.cat([oh(e, 5), oh(e, 5)])] * (T - 1))}
    mt = SampledMolecule(x, a, c, e, amap, fake_atoms=True, traj_frames=frames)
    assert len(mt.traj_mols) == T and len(mt.ep_traj_mols) == T - 1
    assert mt.traj_mols[0].atom_types == ['C', 'O', 'N', 'Sn']                    # fake atoms are shown in trajectory frames
    assert np.abs(mt.traj_mols[0].positions.numpy() - (x + T - 1)).max() < 1e-4   # every frame is aligned onto the last one


def test_operand_image_layout_contract_between_producer_and_consumer():
    """The fp16 (hi, lo) operand images handed from one tensor-core linear to the next (csrc/egemm_p.cuh): the producer epilogue's
    address arithmetic (thread = feature, lane pairs swap packed words, rows of a 128-row tile) must write exactly the SW128
    K-major tile image the consumer's UMMA descriptor reads -- the same image format weights.py builds for the weight operands."""
    from flowmol_b200 import weights as W
    rng = np.random.default_rng(0)
    T, OW = 128, 256
    tile = rng.standard_normal((T, OW)).astype(np.float32)
    hi, lo = W.split_h16(tile)
    XSTAGE = 32768
    img = np.zeros(OW // 64 * XSTAGE, np.uint8)
    hb, lb = hi.view(np.uint16), lo.view(np.uint16)
    # producer (egemm_p.cuh, IMG_OUT block): warp (q, hf) handles features f = mt*128 + q*32 + lane and chunks c = 2*hf + ci
    for mt in range(OW // 128):
        for q in range(4):
            for c in range(4):
                for lane in range(32):
                    odd, kk = lane & 1, (q & 1) * 32 + (lane & ~1)
                    ob = (mt * 2 + (q >> 1)) * XSTAGE + (c * 32 + odd) * 128 + (kk & 7) * 2
                    c0x = (kk >> 3) ^ odd
                    f_even = mt * 128 + q * 32 + (lane & ~1)
                    for t in range(16):
                        row = c * 32 + 2 * t + odd                       # even lane stores row 2t, odd lane row 2t + 1
                        dst = ob + (2 * t) * 128 + ((c0x ^ ((2 * t) & 7)) << 4)
                        for half, src in ((0, hb), (16384, lb)):
                            word = np.array([src[row, f_even], src[row, f_even + 1]], np.uint16).view(np.uint8)   # (k, k + 1), k low
                            img[dst + half:dst + half + 4] = word
    # consumer: slab j = k values [64 j, 64 j + 64) as [hi image | lo image], each a K-major SW128 tile (weights.sw128_image_h16)
    for j in range(OW // 64):
        want_hi = W.sw128_image_h16(hi[:, 64 * j:64 * j + 64]).view(np.uint8)
        want_lo = W.sw128_image_h16(lo[:, 64 * j:64 * j + 64]).view(np.uint8)
        assert np.array_equal(img[j * XSTAGE:j * XSTAGE + 16384], want_hi), j
        assert np.array_equal(img[j * XSTAGE + 16384:(j + 1) * XSTAGE], want_lo), j


def test_sampler_options_schedules_are_evaluated_per_step_like_the_reference():
    """CTMCVectorFieldB200._opts (host side of fm_integrate's options): the categorical-temperature / forward-weight /
    inverse-temperature callables are evaluated once per step on the fp32 time grid exactly as the reference's step() would
    (ctmc_vector_field.py:71-95,334,353,385,491) and handed over as per-step fp32 scalars; 'gat' needs both weight arrays."""
    import ctypes as C
    from types import SimpleNamespace
    from flowmol_b200.vector_field import CTMCVectorFieldB200, build_cat_temp_schedule, build_fw_schedule
    me = SimpleNamespace(eta=30.0, hc_thresh=0.9, cat_temperature=0.05, dfm_type='campbell',
                         cat_temp_func=build_cat_temp_schedule(0.05), forward_weight_func=build_fw_schedule('beta'))
    T = 9
    o, keep = CTMCVectorFieldB200._opts(me, T, None, None, seed=7, mol_id_offset=3, tspan=None, cuda_graph=False)
    assert o.n_timesteps == T and o.dfm_type == 0 and o.stochasticity == 30.0 and o.mol_id_offset == 3
    assert not o.fw_host and not o.bw_host and not o.inv_temp_host
    assert np.allclose(np.ctypeslib.as_array(o.tau_host, (T - 1,)), 0.05)
    assert np.array_equal(np.ctypeslib.as_array(o.tspan_host, (T,)), torch.linspace(0, 1, T).numpy())
    ctf, fwf, itf = build_cat_temp_schedule('decay', 0.8, 2), build_fw_schedule('beta', 0.25, 0.25, 10.0), (lambda t: 1.0 + 0.5 * t)
    o, keep = CTMCVectorFieldB200._opts(me, T, 5.0, 0.0, 7, 0, None, False, 'gat', ctf, fwf, itf)
    tt = torch.linspace(0, 1, T)
    assert o.dfm_type == 1 and o.stochasticity == 5.0 and o.high_confidence_threshold == 0.0
    tau = np.ctypeslib.as_array(o.tau_host, (T - 1,))
    fw, bw = np.ctypeslib.as_array(o.fw_host, (T - 1,)), np.ctypeslib.as_array(o.bw_host, (T - 1,))
    it = np.ctypeslib.as_array(o.inv_temp_host, (T - 1,))
    for k in range(T - 1):
        assert tau[k] == np.float32(0.8 * torch.pow(1 - tt[k], 2)) and it[k] == np.float32(1.0 + 0.5 * tt[k])
        f = 1 + 10.0 * torch.pow(tt[k], 0.25) * torch.pow(1 - tt[k], 0.25)
        assert fw[k] == np.float32(f) and bw[k] == np.float32(f - 1)
    assert fw[0] == 1.0 and bw[0] == 0.0                      # t = 0: pure forward velocity
    with pytest.raises(ValueError):
        CTMCVectorFieldB200._opts(me, T, None, None, 7, 0, None, False, 'bogus')
    if RL_available():
        from oracle import ref_loader as RL
        vf_cfg, sc_cfg = RL.read_vector_field_cfg("dev")
        m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=6)
        for k in range(T - 1):                                 # the reference's own default schedules (constructor defaults)
            assert float(m.forward_weight_func(tt[k])) == float(fwf(tt[k]))
            assert float(torch.as_tensor(m.cat_temp_func(tt[k]))) == float(torch.as_tensor(me.cat_temp_func(tt[k])))


def RL_available():
    from oracle import ref_loader as RL
    return RL.available()


def test_cli_prior_is_drawn_once_per_batch_and_sliced_per_rank():
    """flowmol_b200/cli.py: noise that does not depend on the GPU count -- the COM-free prior positions are drawn for the whole
    batch from one generator and every rank takes its contiguous slice (FlowMolB200.centered_normal / prior_from_x0)."""
    from types import SimpleNamespace
    from flowmol_b200.api import FlowMolB200
    from flowmol_b200 import sharding as SH
    n = np.array([5, 9, 3, 12, 7, 4])
    x0 = FlowMolB200.centered_normal(n, torch.Generator().manual_seed(3))
    assert torch.equal(x0, FlowMolB200.centered_normal(n, torch.Generator().manual_seed(3)))
    noff = np.concatenate([[0], np.cumsum(n)])
    for b in range(len(n)):                                   # every molecule is COM-free (priors.py:27-35)
        assert x0[noff[b]:noff[b + 1]].mean(0).abs().max() < 1e-6
    me = SimpleNamespace(n_atom_types=6, n_bond_types=4, fake_atoms=True)
    parts = []
    for lo, hi in SH.partition(n, 3):
        pr = FlowMolB200.prior_from_x0(me, n[lo:hi], x0[noff[lo]:noff[hi]])
        N, E = int(n[lo:hi].sum()), int((n[lo:hi] * (n[lo:hi] - 1)).sum())
        assert pr['x_0'].shape == (N, 3) and pr['a_0'].shape == (N, 7) and pr['c_0'].shape == (N, 7) and pr['e_0'].shape == (E, 5)
        assert (pr['a_0'].argmax(-1) == 6).all() and (pr['e_0'].argmax(-1) == 4).all() and pr['fake_atoms'] is True
        parts.append(pr['x_0'])
    assert torch.equal(torch.cat(parts), x0)


def test_permuted_feature_layouts_match_the_epilogue_index_algebra():
    """weights.quad_perm / feature_perm (packed entries EUPD_TC1_HP / EUPD_TC2_HP, MSG1_TCW_HP / MSG1_TCG_HP / MSG1_BP) against the
    register <-> accumulator-column maps of the tcgen05.ld / st shapes the epilogues use (csrc/egemm_c.cuh QL, csrc/egemm_h.cuh PERM;
    the maps themselves were measured on the GPU: tools/gpu_tmem_shapes.py)."""
    qp, fp = WT.quad_perm(128), WT.feature_perm(256)
    assert sorted(qp) == list(range(128)) and sorted(fp) == list(range(256))
    # 16x256b.x4 load of column group j: thread t of a row's quad receives accumulator columns 32 j + 8 n + 2 t + c (n < 4, c < 2) -- under
    # quad_perm these are 8 CONSECUTIVE features, the 32 bytes of EA[src] / EB[dst] / the residual row the thread fetches with one load
    for j in range(4):
        for t in range(4):
            cols = [32 * j + 8 * n + 2 * t + c for n in range(4) for c in range(2)]
            assert [int(qp[L]) for L in cols] == list(range(32 * j + 8 * t, 32 * j + 8 * t + 8))
    # epilogue 2: warp half hf, group gj, register i of thread t = accumulator column 64 hf + 8 (4 gj + i // 2) + 2 t + i % 2 is stored
    # to feature 64 hf + 32 gj + 8 t + i (fp32 row) = fp16 elements of 16-byte piece 4 gj + t of k-slab hf (operand image)
    for hf in range(2):
        for gj in range(2):
            for t in range(4):
                for i in range(8):
                    M = 64 * hf + 8 * (4 * gj + i // 2) + 2 * t + i % 2
                    f = 64 * hf + 32 * gj + 8 * t + i
                    assert int(qp[M]) == f and (f % 64) // 8 == 4 * gj + t
    # k_egemm_h<MSG, PERM>: the packed (hi | lo) words of a 32-feature chunk are re-read with 16x128b.x4 -- thread t receives word
    # 4 n + t (n < 4) = accumulator columns 8 n + 2 t, 8 n + 2 t + 1 -- and must hold the physical fp16 pairs 4 t .. 4 t + 3 in order
    for chunk in range(8):
        for t in range(4):
            feats = [int(fp[32 * chunk + 2 * (4 * n + t) + c]) for n in range(4) for c in range(2)]
            assert feats == list(range(32 * chunk + 8 * t, 32 * chunk + 8 * t + 8))
    # the algebra the packed entries rely on: permuting the hidden features of linear 1 and BOTH sides of linear 2 yields the permuted output
    rng = np.random.default_rng(3)
    W1, W2 = rng.standard_normal((128, 160)), rng.standard_normal((128, 128))
    x, ea, b2 = rng.standard_normal(160), rng.standard_normal(128), rng.standard_normal(128)
    silu = lambda z: z / (1.0 + np.exp(-z))
    y = W2 @ silu(W1 @ x + ea) + b2
    h_log = silu(W1[qp] @ x + ea[qp])                       # accumulator column L holds hidden feature qp[L]; the thread adds EA at qp[L]
    y_log = W2[qp][:, qp] @ h_log + b2[qp]                  # k order of linear 2 = logical hidden order (in-place write-back), rows permuted
    assert np.allclose(y_log, y[qp], rtol=1e-12, atol=1e-12)
    # message GVP 1: permuted output features, the gate linear contracts over them in the permuted k order
    Wm, Wg, bm = rng.standard_normal((256, 292)), rng.standard_normal((32, 256)), rng.standard_normal(256)
    xm = rng.standard_normal(292)
    s1 = silu(Wm @ xm + bm)
    s1_log = silu(Wm[fp] @ xm + bm[fp])
    assert np.allclose(s1_log, s1[fp]) and np.allclose(Wg[:, fp] @ s1_log, Wg @ s1)


def test_packer_builds_the_permuted_entries_from_the_natural_ones():
    """The permuted twins the default kernels read are the natural entries with rows / k columns permuted: same unit builder, same
    scale rule (checked on the bias vector and on the first hi unit of the EdgeUpdate linears)."""
    cfg = ModelConfig.named("flowmol3", 11)
    sd = WT.init_state_dict(cfg, 5)
    blob, off = WT.pack(cfg, sd)
    fp, qp = WT.feature_perm(256), WT.quad_perm(128)
    b = sd["conv_layers.2.edge_message.1.to_feats_out.0.bias"].numpy()
    o = off[WL.cid(2, "MSG1_BP")]
    assert o >= 0 and np.array_equal(blob[o:o + 256], b[fp])
    S, F = cfg.n_hidden_scalars, cfg.n_hidden_edge_feats
    w1 = sd["edge_updaters.0.edge_update_fn.0.weight"].numpy().T         # [in, out]
    w2 = sd["edge_updaters.0.edge_update_fn.2.weight"].numpy()           # [out, in]
    for name, w in (("EUPD_TC1_HP", w1[2 * S:].T[qp, :]), ("EUPD_TC2_HP", w2[qp, :][:, qp])):
        want = WT.tc_units_h16(w, 128)
        o = off[WL.uid(cfg.n_convs, 0, name)]
        assert o >= 0 and np.array_equal(blob[o:o + want.size].view(np.uint32), want.view(np.uint32)), name


def test_node_embed_mma_tile_gemm_index_algebra_emulated():
    """csrc/kernels.cuh:tile_gemm_h16 in numpy: weight chunks as packed (k, k + 1) half2 words [16 k pairs][264], the m16n8k16 fragment
    ownership (A: rows g / g + 8, k 2t / 2t + 8; B: k pairs t / t + 4, column g; C: rows g / g + 8, columns 2t, 2t + 1), warp w = output
    columns [32 w, 32 w + 32), the result tile Ys [64][264] and its read-back into the row-per-warp register layout (ColMap) --
    with K = 308 (ragged last chunk, as the self-conditioning linear) the emulation must reproduce the fp16x3 product sum."""
    rng = np.random.default_rng(11)
    K, S, LDW, LDA = 308, 256, 264, 316
    X = np.zeros((64, LDA), np.float32); X[:, :K] = rng.standard_normal((64, K)).astype(np.float32)
    W = (rng.standard_normal((K, S)) * 0.06).astype(np.float32)
    xh, xl = WT.split_h16(X)
    wh, wl = WT.split_h16(W * np.float32(1024.0))
    xh, xl, wh, wl = (a.astype(np.float64) for a in (xh, xl, wh, wl))
    Ys = np.zeros((64, 264))
    for ch in range((K + 31) // 32):
        hi, lo = np.zeros((16, LDW, 2)), np.zeros((16, LDW, 2))          # converted chunk: word (kp, n) = (W[k][n], W[k + 1][n])
        for item in range(1024):
            kp, n4 = item >> 6, item & 63
            for r in range(2):
                k = ch * 32 + 2 * kp + r
                if k < K:
                    hi[kp, 4 * n4:4 * n4 + 4, r], lo[kp, 4 * n4:4 * n4 + 4, r] = wh[k, 4 * n4:4 * n4 + 4], wl[k, 4 * n4:4 * n4 + 4]
        for ks in range(2):
            k0 = ch * 32 + 16 * ks
            if k0 >= K:
                continue
            # operand coverage: A registers a_i <-> (row g + 8 (i & 1), k 2t + 8 (i >> 1) + {0, 1}); B b_i <-> (k pair t + 4 i, column g)
            A_h = np.zeros((64, 16)); A_l = np.zeros((64, 16))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for mt in range(4):
                    for i in range(4):
                        col = k0 + 2 * t + (i >> 1) * 8
                        row = 16 * mt + g + (i & 1) * 8
                        if col < K:
                            A_h[row, col - k0:col - k0 + 2], A_l[row, col - k0:col - k0 + 2] = xh[row, col:col + 2], xl[row, col:col + 2]
            B_h = np.zeros((16, S)); B_l = np.zeros((16, S))
            for warp in range(8):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for nt in range(4):
                        for i in range(2):
                            kp, n = 8 * ks + t + 4 * i, 32 * warp + 8 * nt + g
                            B_h[2 * (t + 4 * i):2 * (t + 4 * i) + 2, n], B_l[2 * (t + 4 * i):2 * (t + 4 * i) + 2, n] = hi[kp, n], lo[kp, n]
            Ys[:, :S] += A_l @ B_h + A_h @ B_l + A_h @ B_h
    Ys[:, :S] /= 1024.0
    want = (xl[:, :K] @ wh + xh[:, :K] @ wl + xh[:, :K] @ wh) / 1024.0
    assert np.allclose(Ys[:, :S], want, rtol=0, atol=1e-9)
    assert np.abs(want - X[:, :K].astype(np.float64) @ W.astype(np.float64)).max() < 2e-5 * np.abs(want).max()      # fp16x3 ~ 2^-22 relative per product
    # read-back: warp w, row r, register c <- Ys[8 w + r][ColMap<8>::col(lane, c)], col = (c // 4) * 128 + lane * 4 + c % 4: every column once
    cols = sorted((c // 4) * 128 + lane * 4 + c % 4 for lane in range(32) for c in range(8))
    assert cols == list(range(256))
