"""Generate tests/golden/*.npz by running the REFERENCE'S OWN MODULES VERBATIM (build container only).

    python -m oracle.make_golden            # needs /root/reference; rewrites tests/golden/

The reference ships no tests or golden vectors, so these fixtures are the pin for the oracle
restatement (oracle/flowmol_oracle.py) and, through it, for the CUDA path.  Weights are NOT stored: they are
re-created from `flowmol_b200.weights.init_state_dict(cfg, seed)` (CPU torch.Generator => identical on every box)
and guarded by `weights_checksum`.  Outputs come from `CTMCVectorField` in /root/reference/flowmol/models
(forward hooks capture the per-layer intermediates), with noise injected as described in oracle/ref_loader.py.
"""
import os
import sys

import numpy as np
import torch

from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from . import ref_loader as RL

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES_FORWARD = [
    # name, config, n_atom_types, weight seed, n_atoms, keep per-layer taps
    ("fwd_dev_taps", "dev", 6, 11, [3, 7, 18], True),
    ("fwd_flowmol3_taps", "flowmol3", 11, 12, [3, 7, 20], True),
    ("fwd_flowmol3_geom", "flowmol3", 11, 12, [3, 18, 46], False),
    ("fwd_dev_qm9", "dev", 6, 11, [29, 3, 18, 9, 17, 4], False),
]
CASES_INTEGRATE = [
    # name, config, A, weight seed, n_atoms, T, noise seed
    ("itg_dev_T10", "dev", 6, 11, [3, 7, 18, 9], 10, 2024),
    ("itg_dev_T50", "dev", 6, 11, [5, 18, 12, 29, 3, 16, 21, 9], 50, 77),
    ("itg_flowmol3_T10", "flowmol3", 11, 12, [3, 12, 20], 10, 99),
    ("itg_flowmol3_T25", "flowmol3", 11, 12, [4, 30, 17], 25, 31337),
    # SURVEY section 8c probe scales (round 2): GEOM-sized flowmol3 molecules, QM9-sized dev batches of 32 / 64, the largest GEOM
    # molecule (181 atoms: several 128-row tiles per destination segment), and one full-length T = 250 trajectory
    ("itg_flowmol3_geom8_T50", "flowmol3", 11, 12, [46, 44, 51, 38, 47, 46, 59, 42], 50, 4242),
    ("itg_dev_qm9x32_T50", "dev", 6, 11, None, 50, 3232),
    ("itg_dev_qm9x64_T100", "dev", 6, 11, None, 100, 6464),
    ("itg_flowmol3_big_T25", "flowmol3", 11, 12, [64, 181], 25, 181),
    ("itg_flowmol3_geom4_T250", "flowmol3", 11, 12, [46, 39, 52, 45], 250, 250250),
]


def qm9_sizes(count, seed):
    """QM9-sized molecule sizes for the `None` entries above: the reference's own histogram file, a seeded draw."""
    n_map, counts = torch.load("/root/reference/data/qm9/train_data_n_atoms_histogram.pt")
    gen = torch.Generator().manual_seed(seed)
    idx = torch.multinomial(counts.float() / counts.sum().float(), count, replacement=True, generator=gen)
    return [int(v) for v in n_map[idx]]


def _model(cfg_name, A, wseed):
    vf_cfg, sc_cfg = RL.read_vector_field_cfg(cfg_name)
    cfg = ModelConfig.from_vector_field_block(vf_cfg, n_atom_types=A)
    m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=A, seed=0)
    sd = WT.init_state_dict(cfg, seed=wseed)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("dummy_param" in k for k in missing.missing_keys), missing
    return cfg, m, sd


def _one_hot_state(R, g, uem, a_idx, c_idx, e_up, A):
    from torch.nn.functional import one_hot
    g.ndata['a_t'] = one_hot(a_idx, A + 1).float()
    g.ndata['c_t'] = one_hot(c_idx, 7).float()
    e = torch.zeros(uem.shape[0], 5)
    oh = one_hot(e_up, 5).float()
    e[uem] = oh
    e[~uem] = oh
    g.edata['e_t'] = e


def gen_forward(name, cfg_name, A, wseed, n_atoms, taps):
    R = RL.load()
    cfg, m, sd = _model(cfg_name, A, wseed)
    gen = torch.Generator().manual_seed(1000 + wseed)
    g, nbi, ebi, uem = RL.build_reference_graph(n_atoms, generator=gen)
    N, U = g.num_nodes(), int(uem.sum())
    out = {"config": cfg_name, "n_atom_types": A, "weight_seed": wseed, "n_atoms": np.array(n_atoms),
           "weights_checksum": WT.weights_checksum(sd)}
    cap = {}
    hooks = []
    if taps:
        for l, conv in enumerate(m.conv_layers):
            hooks.append(conv.register_forward_hook(lambda mod, i, o, l=l: cap.__setitem__(f"conv{l}", o)))
        for u, (pu, eu) in enumerate(zip(m.node_position_updaters, m.edge_updaters)):
            hooks.append(pu.register_forward_hook(lambda mod, i, o, u=u: cap.__setitem__(f"pos{u}", o)))
            hooks.append(eu.register_forward_hook(lambda mod, i, o, u=u: cap.__setitem__(f"eupd{u}", o)))
    # call 1: first sampling step: all-mask state, t = 0, prev None  (runs the self-conditioning pre-pass)
    x_t = g.ndata['x_0'].clone()
    g.ndata['x_t'] = x_t
    a_idx, c_idx, e_up = torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4)
    _one_hot_state(R, g, uem, a_idx, c_idx, e_up, A)
    with torch.no_grad():
        d0 = m(g, t=torch.full((g.batch_size,), 0.0), node_batch_idx=nbi, upper_edge_mask=uem, apply_softmax=True,
               remove_com=True, prev_dst_dict=None)
    out.update({"c0.x_t": x_t.numpy(), "c0.a": a_idx.numpy(), "c0.c": c_idx.numpy(), "c0.e": e_up.numpy(), "c0.t": 0.0})
    out.update({f"c0.out.{k}": v.numpy() for k, v in d0.items()})
    if taps:   # the hooks saw the pre-pass first and were then overwritten by the main pass: keep the main pass
        for k, v in cap.items():
            if isinstance(v, tuple):
                out[f"c0.tap.{k}.s"], out[f"c0.tap.{k}.v"] = v[0].numpy(), v[1].numpy()
            else:
                out[f"c0.tap.{k}"] = v.numpy()
    # call 2: mid-trajectory: partially unmasked random state, t = 0.37, prev = output of call 1
    cap.clear()
    a_idx = torch.randint(0, A + 1, (N,), generator=gen)
    c_idx = torch.randint(0, 7, (N,), generator=gen)
    e_up = torch.randint(0, 5, (U,), generator=gen)
    x_t2 = x_t + 0.3 * torch.randn(N, 3, generator=gen)
    g.ndata['x_t'] = x_t2
    _one_hot_state(R, g, uem, a_idx, c_idx, e_up, A)
    tval = float(torch.tensor(0.37, dtype=torch.float32))
    with torch.no_grad():
        d1 = m(g, t=torch.full((g.batch_size,), tval), node_batch_idx=nbi, upper_edge_mask=uem, apply_softmax=True,
               remove_com=True, prev_dst_dict=d0)
    out.update({"c1.x_t": x_t2.numpy(), "c1.a": a_idx.numpy(), "c1.c": c_idx.numpy(), "c1.e": e_up.numpy(), "c1.t": tval})
    out.update({f"c1.out.{k}": v.numpy() for k, v in d1.items()})
    if taps:
        for k, v in cap.items():
            if isinstance(v, tuple):
                out[f"c1.tap.{k}.s"], out[f"c1.tap.{k}.v"] = v[0].numpy(), v[1].numpy()
            else:
                out[f"c1.tap.{k}"] = v.numpy()
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", N, "U", U, {k: tuple(v.shape) for k, v in d1.items()})


def gen_integrate(name, cfg_name, A, wseed, n_atoms, T, nseed):
    R = RL.load()
    if n_atoms is None:
        n_atoms = qm9_sizes(64 if "x64" in name else 32, nseed)
    cfg, m, sd = _model(cfg_name, A, wseed)
    gen = torch.Generator().manual_seed(2000 + nseed)
    g, nbi, ebi, uem = RL.build_reference_graph(n_atoms, generator=gen)
    N, U = g.num_nodes(), int(uem.sum())
    g.ndata['a_0'] = R.priors.ctmc_masked_prior(N, A)
    g.ndata['c_0'] = R.priors.ctmc_masked_prior(N, 6)
    ep = R.priors.ctmc_masked_prior(U, 4)
    e0 = torch.zeros(uem.shape[0], 5)
    e0[uem] = ep
    e0[~uem] = ep
    g.edata['e_0'] = e0
    x0 = g.ndata['x_0'].clone()
    with torch.no_grad(), RL.injected_noise(m, n_atoms, seed=nseed):
        g2, traj = m.integrate(g, nbi, upper_edge_mask=uem, n_timesteps=T, visualize=True,
                               stochasticity=None, high_confidence_threshold=None)
    out = {"config": cfg_name, "n_atom_types": A, "weight_seed": wseed, "n_atoms": np.array(n_atoms), "T": T,
           "noise_seed": nseed, "weights_checksum": WT.weights_checksum(sd), "x_0": x0.numpy(),
           "x_1": g2.ndata['x_1'].numpy(), "a_1": g2.ndata['a_1'].argmax(-1).numpy(),
           "c_1": g2.ndata['c_1'].argmax(-1).numpy(), "e_1": g2.edata['e_1'][uem].argmax(-1).numpy(),
           # lower triangle must mirror the upper one (ctmc_vector_field.py:397-406)
           "e_1_lower": g2.edata['e_1'][~uem].argmax(-1).numpy()}
    # per-step trajectory of molecule 0 (states after each step) for step-level checks
    out["traj0.x"] = traj[0]['x'].numpy()
    out["traj0.a"] = traj[0]['a'].argmax(-1).numpy()
    out["traj0.x_1_pred"] = traj[0]['x_1_pred'].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", N, "final a", out["a_1"][:10], "masked left", int((out["a_1"] == A).sum()))


# dfm_type='gat' with non-default schedules (ctmc_vector_field.py:71-95,463-510): name, config, A, weight seed, sizes, T, noise seed
CASES_GAT = [("itg_dev_gat_T12", "dev", 6, 11, [4, 9, 15], 12, 555)]
GAT_SCHEDULES = dict(cat_temperature_schedule="decay", cat_temp_decay_max=0.8, cat_temp_decay_a=2.0,
                     forward_weight_schedule="beta", fw_beta_a=0.25, fw_beta_b=0.25, fw_beta_max=10.0)


def gen_integrate_gat(name, cfg_name, A, wseed, n_atoms, T, nseed):
    """The verbatim reference with dfm_type='gat', its own 'decay' temperature and 'beta' forward-weight schedule builders."""
    R = RL.load()
    cfg, m, sd = _model(cfg_name, A, wseed)
    gen = torch.Generator().manual_seed(2000 + nseed)
    g, nbi, ebi, uem = RL.build_reference_graph(n_atoms, generator=gen)
    N, U = g.num_nodes(), int(uem.sum())
    g.ndata['a_0'] = R.priors.ctmc_masked_prior(N, A)
    g.ndata['c_0'] = R.priors.ctmc_masked_prior(N, 6)
    ep = R.priors.ctmc_masked_prior(U, 4)
    e0 = torch.zeros(uem.shape[0], 5)
    e0[uem] = ep
    e0[~uem] = ep
    g.edata['e_0'] = e0
    x0 = g.ndata['x_0'].clone()
    S = GAT_SCHEDULES
    ctf = m.build_cat_temp_schedule(S["cat_temperature_schedule"], S["cat_temp_decay_max"], S["cat_temp_decay_a"])
    fwf = m.build_fw_schedule(S["forward_weight_schedule"], S["fw_beta_a"], S["fw_beta_b"], S["fw_beta_max"])
    with torch.no_grad(), RL.injected_noise(m, n_atoms, seed=nseed):
        g2, traj = m.integrate(g, nbi, upper_edge_mask=uem, n_timesteps=T, visualize=True, dfm_type='gat',
                               stochasticity=None, high_confidence_threshold=None, cat_temp_func=ctf, forward_weight_func=fwf)
    out = {"config": cfg_name, "n_atom_types": A, "weight_seed": wseed, "n_atoms": np.array(n_atoms), "T": T,
           "noise_seed": nseed, "weights_checksum": WT.weights_checksum(sd), "x_0": x0.numpy(), "dfm_type": "gat",
           "x_1": g2.ndata['x_1'].numpy(), "a_1": g2.ndata['a_1'].argmax(-1).numpy(),
           "c_1": g2.ndata['c_1'].argmax(-1).numpy(), "e_1": g2.edata['e_1'][uem].argmax(-1).numpy(),
           "e_1_lower": g2.edata['e_1'][~uem].argmax(-1).numpy()}
    out.update({k: np.array(v) for k, v in S.items()})
    out["traj0.x"] = traj[0]['x'].numpy()
    out["traj0.a"] = traj[0]['a'].argmax(-1).numpy()
    out["traj0.a_1_pred"] = traj[0]['a_1_pred'].argmax(-1).numpy()          # the reference records p itself here: its argmax
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", N, "final a", out["a_1"][:10], "masked left", int((out["a_1"] == A).sum()))


def gen_ctmc_cases():
    """campbell_step / purity_sampling on crafted states (h = 0, h = m, nothing masked, last step)."""
    R = RL.load()
    vf_cfg, sc_cfg = RL.read_vector_field_cfg("dev")
    m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=6, seed=0)
    gen = torch.Generator().manual_seed(5)
    out = {}
    n_per = [4, 6, 5, 3]
    B = len(n_per)
    n = sum(n_per)
    batch_idx = torch.arange(B).repeat_interleave(torch.tensor(n_per))
    K = 4
    for ci, (last, eta) in enumerate(((False, 20.0), (True, 20.0), (False, 0.0))):
        p = torch.softmax(3.0 * torch.randn(n, K, generator=gen), -1)
        p[0:4] = torch.tensor([0.25, 0.25, 0.25, 0.25])                 # mol 0: nothing high-confidence (h = 0)
        p[4:10] = torch.tensor([0.97, 0.01, 0.01, 0.01])                # mol 1: everything high-confidence (h = m)
        xt = torch.full((n,), K)
        xt[10:15] = torch.tensor([0, 1, K, 2, K])                        # mol 2: mixed
        xt[15:18] = torch.tensor([1, 2, 3])                              # mol 3: nothing masked (m = 0)
        u = [torch.rand(n, generator=gen) for _ in range(3)]

        class Tape:
            k = 0

        def fake_rand(nn, device=None):
            Tape.k += 1
            return u[Tape.k]

        class FakeCat:
            def __init__(self, probs):
                self.probs = probs

            def sample(self):
                c = torch.cumsum(self.probs, -1)
                return torch.clamp((c <= (u[0] * c[:, -1]).unsqueeze(-1)).sum(-1), max=K - 1)

        class Proxy:
            rand = staticmethod(fake_rand)

            def __getattr__(self, nm):
                return getattr(torch, nm)
        saved = (R.ctmc.Categorical, R.ctmc.torch, R.ctmc_utils.torch)
        R.ctmc.Categorical, R.ctmc.torch, R.ctmc_utils.torch = FakeCat, Proxy(), Proxy()
        try:
            t_i, s_i = torch.tensor(0.4), torch.tensor(0.45)
            xt_new, x1 = m.campbell_step(p_1_given_t=p, xt=xt.clone(), stochasticity=eta, hc_thresh=0.9,
                                         alpha_t=t_i, alpha_t_prime=torch.tensor(1.0), dt=s_i - t_i, batch_size=B,
                                         batch_num_nodes=torch.tensor(n_per), n_classes=K + 1, mask_index=K,
                                         last_step=last, batch_idx=batch_idx)
        finally:
            R.ctmc.Categorical, R.ctmc.torch, R.ctmc_utils.torch = saved
        out.update({f"k{ci}.p": p.numpy(), f"k{ci}.xt": xt.numpy(), f"k{ci}.u": torch.stack(u).numpy(),
                    f"k{ci}.last": last, f"k{ci}.eta": eta, f"k{ci}.xt_new": xt_new.argmax(-1).numpy(),
                    f"k{ci}.x1": x1.argmax(-1).numpy()})
    out["n_per"] = np.array(n_per)
    np.savez_compressed(os.path.join(OUT, "ctmc_cases.npz"), **out)
    print("ctmc_cases", {k: v.tolist() for k, v in out.items() if k.endswith("xt_new")})


def main():
    """`python -m oracle.make_golden [name ...]`: all fixtures, or only the named ones."""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = set(sys.argv[1:])
    for c in CASES_FORWARD:
        if not only or c[0] in only:
            gen_forward(*c)
    for c in CASES_INTEGRATE:
        if not only or c[0] in only:
            gen_integrate(*c)
    for c in CASES_GAT:
        if not only or c[0] in only:
            gen_integrate_gat(*c)
    if not only or "ctmc_cases" in only:
        gen_ctmc_cases()


if __name__ == "__main__":
    sys.exit(main())
