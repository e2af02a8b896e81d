"""CPU: the oracle restatement reproduces the golden vectors produced by the reference's own code (verbatim)."""
import numpy as np
import pytest
import torch

from oracle import flowmol_oracle as O
from oracle import philox
from tests.helpers import load_golden, model_from_golden, t

FWD = ["fwd_dev_taps", "fwd_flowmol3_taps", "fwd_flowmol3_geom", "fwd_dev_qm9"]
ITG = ["itg_dev_T10", "itg_dev_T50", "itg_flowmol3_T10", "itg_flowmol3_T25"]

# same torch CPU kernels on the same shapes: the restatement is expected to agree to the last bit or two
TOL = dict(rtol=0, atol=2e-6)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = tuple(int(x) for x in philox.philox4x32_10(*c, *k))
        assert got == want
    u = philox.uniforms(np.arange(1000, dtype=np.uint32), 3, 7, 1, 12345)
    for x in u:
        assert x.dtype == np.float32 and x.min() >= 0.0 and x.max() < 1.0


@pytest.mark.parametrize("name", FWD)
def test_forward_matches_reference(name):
    gd = load_golden(name)
    cfg, sd, om = model_from_golden(gd)
    bt = O.make_batch(gd["n_atoms"])
    with torch.no_grad():
        taps0 = {}
        d0 = om.forward(bt, t(gd["c0.x_t"]), t(gd["c0.a"]), t(gd["c0.c"]), t(gd["c0.e"]), float(gd["c0.t"]), None, taps0)
        taps1 = {}
        prev = {k: t(gd[f"c0.out.{k}"]) for k in "xace"}
        d1 = om.forward(bt, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]), prev, taps1)
    for tag, d, taps in (("c0", d0, taps0), ("c1", d1, taps1)):
        for k in "xace":
            np.testing.assert_allclose(d[k].numpy(), gd[f"{tag}.out.{k}"], **TOL, err_msg=f"{name} {tag} {k}")
        for key in [k for k in gd if k.startswith(f"{tag}.tap.conv") and k.endswith(".s")]:
            l = int(key.split("conv")[1].split(".")[0])
            np.testing.assert_allclose(taps[f"conv{l}.s"].numpy(), gd[key], **TOL)
            np.testing.assert_allclose(taps[f"conv{l}.v"].numpy(), gd[key[:-2] + ".v"], **TOL)
        for key in [k for k in gd if k.startswith(f"{tag}.tap.eupd")]:
            u = int(key.split("eupd")[1])
            np.testing.assert_allclose(taps[f"upd{u}.ef"].numpy(), gd[key], **TOL)
            np.testing.assert_allclose(taps[f"upd{u}.x"].numpy(), gd[f"{tag}.tap.pos{u}"], **TOL)


@pytest.mark.parametrize("name", ITG)
def test_integrate_matches_reference(name):
    gd = load_golden(name)
    cfg, sd, om = model_from_golden(gd)
    bt = O.make_batch(gd["n_atoms"])
    A = cfg.n_atom_types
    rec = []
    with torch.no_grad():
        out = O.integrate(om, bt, t(gd["x_0"]), torch.full((bt.N,), A), torch.full((bt.N,), 6), torch.full((bt.U,), 4),
                          int(gd["T"]), seed=int(gd["noise_seed"]), record=rec)
    assert np.array_equal(out["a"].numpy(), gd["a_1"])
    assert np.array_equal(out["c"].numpy(), gd["c_1"])
    assert np.array_equal(out["e"].numpy(), gd["e_1"])
    assert np.array_equal(gd["e_1"], gd["e_1_lower"])
    np.testing.assert_allclose(out["x"].numpy(), gd["x_1"], rtol=0, atol=1e-5)
    n0 = int(gd["n_atoms"][0])
    for s, r in enumerate(rec):                       # molecule 0, every step
        np.testing.assert_allclose(r["x"][:n0].numpy(), gd["traj0.x"][s + 1], rtol=0, atol=1e-5)
        assert np.array_equal(r["a"][:n0].numpy(), gd["traj0.a"][s + 1])
    assert (out["a"] != A).all() and (out["c"] != 6).all() and (out["e"] != 4).all()   # every mask resolved at t = 1


def test_campbell_step_crafted_cases():
    gd = load_golden("ctmc_cases")
    n_per = torch.from_numpy(gd["n_per"])
    item_mol = torch.arange(len(n_per)).repeat_interleave(n_per)
    for ci in range(3):
        p, xt, u = t(gd[f"k{ci}.p"]), t(gd[f"k{ci}.xt"]), t(gd[f"k{ci}.u"])
        t_i, s_i = torch.tensor(0.4), torch.tensor(0.45)
        xt_new, x1 = O.campbell_step(p, xt, float(gd[f"k{ci}.eta"]), 0.9, t_i, torch.tensor(1.0), s_i - t_i, n_per,
                                     item_mol, 4, bool(gd[f"k{ci}.last"]), (u[0], u[1], u[2]))
        assert np.array_equal(xt_new.numpy(), gd[f"k{ci}.xt_new"])
        assert np.array_equal(x1.numpy(), gd[f"k{ci}.x1"])


def test_fp64_yardstick():
    """fp32 restatement vs fp64 restatement on one forward: the reference's own noise floor."""
    gd = load_golden("fwd_flowmol3_geom")
    cfg, sd, om32 = model_from_golden(gd)
    om64 = O.OracleModel(cfg, sd, dtype=torch.float64)
    bt = O.make_batch(gd["n_atoms"])
    prev = {k: t(gd[f"c0.out.{k}"]) for k in "xace"}
    args = (bt, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]))
    with torch.no_grad():
        d32 = om32.forward(*args, prev)
        d64 = om64.forward(*args, {k: v.double() for k, v in prev.items()})
    assert (d32["x"].double() - d64["x"]).abs().max() < 5e-6
    for k in "ace":
        assert (d32[k].double() - d64[k]).abs().max() < 2e-6
        assert torch.equal(d32[k].argmax(-1), d64[k].argmax(-1))
