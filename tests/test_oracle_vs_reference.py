"""CPU, build container only: restatement vs the reference's modules run verbatim (skipped where /root/reference is absent)."""
import pytest
import torch

from oracle import flowmol_oracle as O
from oracle import ref_loader as RL
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig, NAMED_VECTOR_FIELDS

pytestmark = pytest.mark.skipif(not RL.available(), reason="reference tree not present on this box")


@pytest.mark.parametrize("name,A", [("dev", 6), ("flowmol3", 11)])
def test_embedded_config_and_state_dict_layout_match_reference(name, A):
    vf_cfg, sc_cfg = RL.read_vector_field_cfg(name)
    assert vf_cfg == NAMED_VECTOR_FIELDS[name]
    assert all(v == 'linear' for v in sc_cfg['schedule_type'].values())
    cfg = ModelConfig.from_vector_field_block(vf_cfg, n_atom_types=A)
    m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=A)
    ref = {k: tuple(v.shape) for k, v in m.state_dict().items() if "dummy_param" not in k}
    assert list(ref.items()) == [(k, tuple(s)) for k, s in WT.expected_tensors(cfg)]
    assert m.eta == cfg.stochasticity and m.hc_thresh == cfg.high_confidence_threshold
    assert m.cat_temp_func(0.3) == cfg.cat_temperature


@pytest.mark.parametrize("name,A,n_atoms,T", [("dev", 6, [4, 11, 7], 8), ("flowmol3", 11, [3, 9], 6)])
def test_integrate_bitwise_against_verbatim_reference(name, A, n_atoms, T):
    R = RL.load()
    vf_cfg, sc_cfg = RL.read_vector_field_cfg(name)
    cfg = ModelConfig.from_vector_field_block(vf_cfg, n_atom_types=A)
    m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=A)
    sd = WT.init_state_dict(cfg, seed=3)
    m.load_state_dict(sd, strict=False)
    g, nbi, ebi, uem = RL.build_reference_graph(n_atoms, generator=torch.Generator().manual_seed(4))
    N, U = g.num_nodes(), int(uem.sum())
    g.ndata['a_0'] = R.priors.ctmc_masked_prior(N, A)
    g.ndata['c_0'] = R.priors.ctmc_masked_prior(N, 6)
    ep = R.priors.ctmc_masked_prior(U, 4)
    e0 = torch.zeros(uem.shape[0], 5)
    e0[uem] = ep
    e0[~uem] = ep
    g.edata['e_0'] = e0
    x0 = g.ndata['x_0'].clone()
    with torch.no_grad(), RL.injected_noise(m, n_atoms, seed=17):
        g2 = m.integrate(g, nbi, upper_edge_mask=uem, n_timesteps=T, stochasticity=None, high_confidence_threshold=None)
    bt = O.make_batch(n_atoms)
    assert torch.equal(bt.src, g.edges()[0]) and torch.equal(bt.dst, g.edges()[1]) and torch.equal(bt.upper, uem)
    with torch.no_grad():
        out = O.integrate(O.OracleModel(cfg, sd), bt, x0, torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4),
                          T, seed=17)
    assert torch.equal(out['a'], g2.ndata['a_1'].argmax(-1))
    assert torch.equal(out['c'], g2.ndata['c_1'].argmax(-1))
    assert torch.equal(out['e'], g2.edata['e_1'][uem].argmax(-1))
    assert (out['x'] - g2.ndata['x_1']).abs().max() <= 1e-6


def test_gat_sampler_and_schedules_bitwise_against_verbatim_reference():
    """dfm_type='gat' (ctmc_vector_field.py:463-510) with the reference's default 'beta' forward-weight schedule and a decaying
    categorical temperature (:71-95), plus an inverse-temperature factor on the position update (:334)."""
    R = RL.load()
    name, A, n_atoms, T = "dev", 6, [4, 9, 3], 9
    vf_cfg, sc_cfg = RL.read_vector_field_cfg(name)
    cfg = ModelConfig.from_vector_field_block(vf_cfg, n_atom_types=A)
    m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=A)
    sd = WT.init_state_dict(cfg, seed=5)
    m.load_state_dict(sd, strict=False)
    g, nbi, ebi, uem = RL.build_reference_graph(n_atoms, generator=torch.Generator().manual_seed(6))
    N, U = g.num_nodes(), int(uem.sum())
    g.ndata['a_0'] = R.priors.ctmc_masked_prior(N, A)
    g.ndata['c_0'] = R.priors.ctmc_masked_prior(N, 6)
    ep = R.priors.ctmc_masked_prior(U, 4)
    e0 = torch.zeros(uem.shape[0], 5)
    e0[uem] = ep
    e0[~uem] = ep
    g.edata['e_0'] = e0
    x0 = g.ndata['x_0'].clone()
    ctf = m.build_cat_temp_schedule('decay', 0.8, 2)
    fwf = m.build_fw_schedule('beta', 0.25, 0.25, 10.0)
    itf = lambda t: 1.0 + 0.5 * t
    with torch.no_grad(), RL.injected_noise(m, n_atoms, seed=23):
        g2 = m.integrate(g, nbi, upper_edge_mask=uem, n_timesteps=T, dfm_type='gat', stochasticity=None,
                         high_confidence_threshold=None, cat_temp_func=ctf, forward_weight_func=fwf, inv_temp_func=itf)
    bt = O.make_batch(n_atoms)
    with torch.no_grad():
        out = O.integrate(O.OracleModel(cfg, sd), bt, x0, torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4),
                          T, seed=23, dfm_type='gat', cat_temp_func=ctf, forward_weight_func=fwf, inv_temp_func=itf)
    assert torch.equal(out['a'], g2.ndata['a_1'].argmax(-1))
    assert torch.equal(out['c'], g2.ndata['c_1'].argmax(-1))
    assert torch.equal(out['e'], g2.edata['e_1'][uem].argmax(-1))
    assert (out['x'] - g2.ndata['x_1']).abs().max() <= 1e-6
