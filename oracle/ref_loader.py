"""Run the reference's own sampling-path modules VERBATIM from /root/reference (TEST INFRASTRUCTURE ONLY).

Works only in the build container (the GPU box has no /root/reference); used by oracle/make_golden.py to
produce tests/golden/*.npz and by tests/test_oracle_vs_reference.py (skipped when the tree is absent).

Recipe:
  1. register `flowmol`, `flowmol.models`, `flowmol.utils`, `flowmol.data_processing` as empty namespace
     packages pointing into /root/reference/flowmol, so flowmol/__init__.py (needs Lightning, rdkit) is skipped;
  2. put oracle/refshim (stand-ins for dgl, torch_scatter) on sys.path;
  3. import flowmol.models.ctmc_vector_field etc. unmodified.

Noise injection ("identical noise seeds"): inside `injected_noise(...)` the names `Categorical` and `torch`
of flowmol.models.ctmc_vector_field / flowmol.utils.ctmc_utils are swapped for an inverse-CDF sampler and a
proxy whose `rand` reads the Philox stream of oracle/philox.py.  The reference's control flow and arithmetic
are untouched.
"""
import contextlib
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import philox

REF_ROOT = os.environ.get("FLOWMOL_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "flowmol", "models"))


_loaded = {}


def load():
    """Import the verbatim reference modules; returns a namespace of them."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    for name, sub in (("flowmol", ""), ("flowmol.models", "models"), ("flowmol.utils", "utils"),
                      ("flowmol.data_processing", "data_processing")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, "flowmol", sub)]
            sys.modules[name] = m
    mods = {
        "ctmc": "flowmol.models.ctmc_vector_field",
        "vf": "flowmol.models.vector_field",
        "gvp": "flowmol.models.gvp",
        "sc": "flowmol.models.self_conditioning",
        "sched": "flowmol.models.interpolant_scheduler",
        "emb": "flowmol.utils.embedding",
        "ctmc_utils": "flowmol.utils.ctmc_utils",
        "dutils": "flowmol.data_processing.utils",
        "priors": "flowmol.data_processing.priors",
    }
    for k, v in mods.items():
        _loaded[k] = importlib.import_module(v)
    import dgl  # the shim
    _loaded["dgl"] = dgl
    return types.SimpleNamespace(**_loaded)


def load_function_verbatim(relpath, func_name, namespace):
    """Compile ONE top-level function of a reference file from its own source text, unmodified, into `namespace` -- for modules
    whose import needs packages this image lacks (flowmol/analysis/molecule_builder.py imports rdkit at the top; its
    `extract_moldata_from_graph` needs only torch and the graph container)."""
    import ast
    path = os.path.join(REF_ROOT, relpath)
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == func_name:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)
            return namespace[func_name]
    raise KeyError(f"{func_name} not found in {path}")


def read_vector_field_cfg(name):
    """`vector_field:` and `interpolant_scheduler:` blocks of configs/{dev,flowmol3}.yml (reference file)."""
    import yaml
    with open(os.path.join(REF_ROOT, "configs", f"{name}.yml")) as f:
        cfg = yaml.safe_load(f)
    return cfg["vector_field"], cfg["interpolant_scheduler"]


def build_reference_model(vf_cfg, sched_cfg, n_atom_types, seed=0):
    """Mirror of flowmol/models/flowmol.py:130-153 (CTMC branch)."""
    R = load()
    torch.manual_seed(seed)
    order = ['x', 'a', 'c', 'e']
    sched = R.sched.InterpolantScheduler(canonical_feat_order=order, **sched_cfg)
    m = R.ctmc.CTMCVectorField(n_atom_types=n_atom_types, canonical_feat_order=order,
                               interpolant_scheduler=sched, n_charges=6, n_bond_types=4,
                               exclude_charges=False, fake_atoms=True, **vf_cfg)
    return m.eval()


def build_reference_graph(n_atoms, x0=None, generator=None):
    """Graph + prior exactly as flowmol/models/flowmol.py:509-545; x0 optional (COM-free positions)."""
    R = load()
    dgl = R.dgl
    gs = []
    for n in n_atoms:
        e = R.dutils.build_edge_idxs(int(n))
        gs.append(dgl.graph((e[0], e[1]), num_nodes=int(n), device="cpu"))
    g = dgl.batch(gs)
    uem = R.dutils.get_upper_edge_mask(g)
    nbi, ebi = R.dutils.get_batch_idxs(g)
    N = g.num_nodes()
    if x0 is None:
        x0 = torch.randn(N, 3, generator=generator)
        g.ndata['tmp'] = x0
        x0 = x0 - dgl.readout_nodes(g, feat='tmp', op='mean')[nbi]
        del g.ndata['tmp']
    g.ndata['x_0'] = x0
    return g, nbi, ebi, uem


class _InverseCdfCategorical:
    """Replacement for torch.distributions.Categorical inside the reference's campbell_step."""

    def __init__(self, noise, probs):
        self.noise, self.probs = noise, probs

    def sample(self):
        u = self.noise.next_uniform(self.probs.shape[0], kind=0)
        c = torch.cumsum(self.probs, dim=-1)          # fp32 sequential prefix sum
        thr = (u * c[:, -1]).unsqueeze(-1)
        k = (c <= thr).sum(-1)
        return torch.clamp(k, max=self.probs.shape[1] - 1)


class NoiseTape:
    """Philox stream addressed by (item, molecule, step, modality); hands out uniforms in reference call order."""

    def __init__(self, n_atoms, seed, mol_id_offset=0):
        n_atoms = [int(n) for n in n_atoms]
        self.seed = int(seed)
        node_mol, node_item, edge_mol, edge_item = [], [], [], []
        for b, n in enumerate(n_atoms):
            node_mol += [b + mol_id_offset] * n
            node_item += list(range(n))
            u = n * (n - 1) // 2
            edge_mol += [b + mol_id_offset] * u
            edge_item += list(range(u))
        self.items = {
            'node': (np.array(node_item, dtype=np.uint32), np.array(node_mol, dtype=np.uint32)),
            'edge': (np.array(edge_item, dtype=np.uint32), np.array(edge_mol, dtype=np.uint32)),
        }
        self.step = 0
        self.n_cat_calls = 0

    def begin_step(self, step):
        self.step = int(step)
        self.n_cat_calls = 0

    def next_uniform(self, n, kind):
        if kind == 0:
            self.n_cat_calls += 1
        modality = self.n_cat_calls - 1               # 0=a, 1=c, 2=e (canonical_feat_order minus 'x')
        item, mol = self.items['edge' if modality == 2 else 'node']
        assert item.shape[0] == n, (item.shape, n, modality)
        u = philox.uniforms(item, mol, self.step, modality, self.seed)[kind]
        return torch.from_numpy(u)


class _TorchProxy:
    """`torch` as seen by the patched reference modules: only `rand` differs."""

    def __init__(self, noise):
        self._noise = noise
        self._rand_calls_in_modality = 0
        self._last_cat = -1

    def rand(self, n, device=None, **kw):
        if self._noise.n_cat_calls != self._last_cat:
            self._last_cat = self._noise.n_cat_calls
            self._rand_calls_in_modality = 0
        self._rand_calls_in_modality += 1
        return self._noise.next_uniform(int(n), kind=self._rand_calls_in_modality)

    def __getattr__(self, name):
        return getattr(torch, name)


@contextlib.contextmanager
def injected_noise(model, n_atoms, seed, mol_id_offset=0):
    """Within this context `model.integrate(...)` consumes the Philox stream instead of the global torch RNG."""
    R = load()
    noise = NoiseTape(n_atoms, seed, mol_id_offset)
    proxy = _TorchProxy(noise)
    saved = (R.ctmc.Categorical, R.ctmc.torch, R.ctmc_utils.torch, type(model).step)
    counter = {'s': 0}
    orig_step = type(model).step

    def step_with_counter(self, *a, **k):
        counter['s'] += 1
        noise.begin_step(counter['s'])
        return orig_step(self, *a, **k)

    R.ctmc.Categorical = lambda probs: _InverseCdfCategorical(noise, probs)
    R.ctmc.torch = proxy
    R.ctmc_utils.torch = proxy
    type(model).step = step_with_counter
    try:
        yield noise
    finally:
        R.ctmc.Categorical, R.ctmc.torch, R.ctmc_utils.torch = saved[:3]
        type(model).step = saved[3]
