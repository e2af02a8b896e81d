"""User-facing surface of the reference kept for the sampling path.

    flowmol.load_pretrained(model_name)                       flowmol/__init__.py:30-56
    model.sample_random_sizes(n_molecules, n_timesteps=...)   flowmol/models/flowmol.py:473-486
    model.sample(n_atoms, n_timesteps, prior=...)             flowmol/models/flowmol.py:489-589
    SampledMolecule                                           flowmol/analysis/molecule_builder.py:17-85,217-297

Out of scope: training, losses, metrics, dataset processing.
"""
import json
import os
from pathlib import Path

import numpy as np
import torch
from torch.nn.functional import one_hot

from . import weights as WT
from .config import GEOM_ATOM_MAP, QM9_ATOM_MAP, ModelConfig
from .graph import MolGraphBatch
from .vector_field import CTMCVectorFieldB200

# flowmol/__init__.py:5-28
pretrained_model_names = [
    'flowmol3', 'fm3_nodistort', 'fm3_none', 'fm3_ahigh', 'fm3_alow', 'fm3_chigh', 'fm3_clow', 'fm3_distort_extreme',
    'fm3_distort_highp', 'fm3_distort_hight', 'fm3_distort_lowp', 'fm3_distort_lowt', 'fm3_ehigh', 'fm3_elow',
    'fm3_fa_highp', 'fm3_fa_highstd', 'fm3_fa_lowp', 'fm3_fa_lowstd', 'fm3_scprop_high', 'fm3_scprop_low', 'fm3_xhigh',
    'fm3_xlow']

_HERE = Path(__file__).parent


def n_atoms_histogram(dataset):
    """(n_atoms int64[K], counts float64[K]) of the training set (data/<dataset>/train_data_n_atoms_histogram.pt)."""
    with open(_HERE / "data" / "n_atoms_hist.json") as f:
        h = json.load(f)[dataset]
    return torch.tensor(h["n_atoms"], dtype=torch.int64), torch.tensor(h["counts"], dtype=torch.float64)


class SampledMolecule:
    """A sampled molecule: decoded arrays always; an rdkit Mol when rdkit is importable (molecule_builder.py:17-85).

    Decode rules (molecule_builder.py:217-265): atom type = argmax a_1 with fake atoms dropped, charge = argmax c_1 - 2,
    bond order = argmax e_1 on upper-triangle edges with the mask token mapped to 0, zero-order bonds dropped."""

    def __init__(self, x, a_idx, c_idx, e_idx_upper, atom_type_map, fake_atoms=True, explicit_aromaticity=False,
                 traj_frames=None, build_xt_traj=True, build_ep_traj=True, align_traj=True, show_fake_atoms=False):
        atom_type_map = list(atom_type_map)
        n = int(x.shape[0])
        keep = np.ones(n, dtype=bool)
        if fake_atoms and not show_fake_atoms:
            keep = a_idx != len(atom_type_map)              # the fake-atom token sits right after the real types
        symbols_all = atom_type_map + (['Sn'] if fake_atoms else []) + ['Se']     # 'Se' marks a still-masked atom
        new_index = np.cumsum(keep) - 1
        iu = np.triu_indices(n, k=1)
        bond = e_idx_upper.copy()
        bond[bond == (5 if explicit_aromaticity else 4)] = 0
        sel = (bond != 0) & keep[iu[0]] & keep[iu[1]]
        self.positions = torch.from_numpy(np.ascontiguousarray(x[keep]))
        self.atom_types = [symbols_all[int(t)] for t in a_idx[keep]]
        self.atom_charges = torch.from_numpy((c_idx[keep].astype(np.int64) - 2))
        self.bond_types = torch.from_numpy(bond[sel].astype(np.int64))
        self.bond_src_idxs = torch.from_numpy(new_index[iu[0][sel]].astype(np.int64))
        self.bond_dst_idxs = torch.from_numpy(new_index[iu[1][sel]].astype(np.int64))
        self.num_atoms = int(keep.sum())
        self.atom_type_map = atom_type_map
        self.fake_atoms, self.explicit_aromaticity, self.align_traj = fake_atoms, explicit_aromaticity, align_traj
        self.rdkit_mol = self.build_molecule()
        # trajectories (molecule_builder.py:75-84): one decoded frame molecule per step, aligned to the last frame
        self.traj_frames = traj_frames
        self.traj_mols, self.ep_traj_mols = None, None
        if traj_frames is not None:
            if build_xt_traj:
                self.traj_mols = self.process_traj_frames(traj_frames)
            if build_ep_traj and 'x_1_pred' in traj_frames:
                self.ep_traj_mols = self.process_traj_frames(traj_frames, ep_traj=True)

    @classmethod
    def from_decoded(cls, positions, atom_tokens, charges, bond_types, bond_src, bond_dst, atom_type_map, fake_atoms=True,
                     explicit_aromaticity=False):
        """A molecule from the arrays the device-side decode (fm_decode) produced: surviving atoms only, renumbered bonds."""
        self = cls.__new__(cls)
        atom_type_map = list(atom_type_map)
        symbols_all = atom_type_map + (['Sn'] if fake_atoms else []) + ['Se']
        self.positions = positions
        self.atom_types = [symbols_all[int(t)] for t in atom_tokens]
        self.atom_charges = charges.to(torch.int64)
        self.bond_types = bond_types.to(torch.int64)
        self.bond_src_idxs = bond_src.to(torch.int64)
        self.bond_dst_idxs = bond_dst.to(torch.int64)
        self.num_atoms = int(positions.shape[0])
        self.atom_type_map = atom_type_map
        self.fake_atoms, self.explicit_aromaticity, self.align_traj = fake_atoms, explicit_aromaticity, True
        self.rdkit_mol = self.build_molecule()
        self.traj_frames, self.traj_mols, self.ep_traj_mols = None, None, None
        return self

    def process_traj_frames(self, traj_frames, ep_traj=False):
        """molecule_builder.py:156-214: every frame decoded like a final molecule, fake atoms shown (as 'Sn'), positions
        rigidly aligned to the last frame.  Returns SampledMolecule objects (their `.rdkit_mol` is set when rdkit is present)."""
        sfx = '_1_pred' if ep_traj else ''
        xs = traj_frames['x' + sfx]
        x_final = xs[-1].numpy()
        n = x_final.shape[0]
        n_up = n * (n - 1) // 2
        mols = []
        for k in range(xs.shape[0]):
            pos = xs[k].numpy()
            if self.align_traj:
                pos = rigid_alignment(pos, x_final)
            a = traj_frames['a' + sfx][k].argmax(-1).numpy()
            c = traj_frames['c' + sfx][k].argmax(-1).numpy()
            e = traj_frames['e' + sfx][k][:n_up].argmax(-1).numpy()
            mols.append(SampledMolecule(pos.astype(np.float32), a, c, e, self.atom_type_map, fake_atoms=self.fake_atoms,
                                        explicit_aromaticity=self.explicit_aromaticity, show_fake_atoms=True))
        return mols

    def build_molecule(self):
        try:
            from rdkit import Chem
            from rdkit.Geometry import Point3D
        except ImportError:
            return None
        order = [None, Chem.rdchem.BondType.SINGLE, Chem.rdchem.BondType.DOUBLE, Chem.rdchem.BondType.TRIPLE,
                 Chem.rdchem.BondType.AROMATIC]
        mol = Chem.RWMol()
        for sym, ch in zip(self.atom_types, self.atom_charges.tolist()):
            at = Chem.Atom(sym)
            if ch != 0:
                at.SetFormalCharge(int(ch))
            mol.AddAtom(at)
        for bt, s, d in zip(self.bond_types.tolist(), self.bond_src_idxs.tolist(), self.bond_dst_idxs.tolist()):
            mol.AddBond(int(s), int(d), order[bt])
        try:
            mol = mol.GetMol()
        except Exception:
            return None
        conf = Chem.Conformer(mol.GetNumAtoms())
        for i, p in enumerate(self.positions.tolist()):
            conf.SetAtomPosition(i, Point3D(*map(float, p)))
        mol.AddConformer(conf)
        return mol


def rigid_alignment(x_0, x_1):
    """Kabsch alignment of x_0 onto x_1 (flowmol/data_processing/priors.py:128-169: R = V U^T from the SVD of x_0c^T x_1c, no
    reflection correction), numpy [n,3] -> [n,3].  The reference adds `x_0_mean - R x_0_mean` on top, which vanishes for the
    COM-free frames it is applied to; here the aligned cloud is simply centred on x_1's centroid."""
    x_0, x_1 = np.asarray(x_0, dtype=np.float64), np.asarray(x_1, dtype=np.float64)
    m0, m1 = x_0.mean(0, keepdims=True), x_1.mean(0, keepdims=True)
    a, b = x_0 - m0, x_1 - m1
    u, _, vt = np.linalg.svd(a.T @ b)
    rot = vt.T @ u.T
    return (a @ rot.T + m1).astype(np.float32)


class FlowMolB200:
    """Sampling-only counterpart of `FlowMol` (flowmol/models/flowmol.py:23)."""
    canonical_feat_order = ['x', 'a', 'c', 'e']

    def __init__(self, atom_type_map, vector_field_config, state_dict, n_atoms_hist=None, dataset="geom",
                 default_n_timesteps=250, fake_atom_p=0.3, explicit_aromaticity=False, device="cuda:0", vector_field=None):
        self.atom_type_map = list(atom_type_map)
        self.fake_atoms = fake_atom_p > 0
        self.explicit_aromaticity = explicit_aromaticity
        self.n_atom_types = len(self.atom_type_map) + int(self.fake_atoms)          # flowmol.py:58,78-80
        self.n_bond_types = 5 if explicit_aromaticity else 4
        self.default_n_timesteps = default_n_timesteps
        self.cfg = ModelConfig.from_vector_field_block(vector_field_config, self.n_atom_types, 6, self.n_bond_types)
        self._state_dict = state_dict
        self._device = device
        self.vector_field = None
        self.n_atoms_map, counts = n_atoms_hist if n_atoms_hist is not None else n_atoms_histogram(dataset)
        self.n_atoms_dist = torch.distributions.Categorical(probs=counts / counts.sum())       # flowmol.py:461-466
        if vector_field is not None:                      # adopt an already built device model (same config and weights)
            self.vector_field = vector_field
            self._device = str(vector_field.device)
        elif torch.cuda.is_available():
            self._materialise(device)

    # -- construction ------------------------------------------------------------------------------------------------------
    @classmethod
    def from_checkpoint(cls, ckpt_path, device="cuda:0", **load_kwargs):
        sd, hp = WT.state_dict_from_checkpoint(ckpt_path)
        # the kernels hard-wire what every shipped config uses: the CTMC parameterisation, the linear
        # interpolant (alpha = t, alpha' = 1) for all four modalities and 6 charge classes.  Anything else must not load silently.
        param = hp.get("parameterization", "ctmc")
        if param != "ctmc":
            raise NotImplementedError(f"checkpoint parameterization {param!r}: only 'ctmc' is built (flowmol.py:190-193)")
        sched = (hp.get("interpolant_scheduler_config") or {}).get("schedule_type", "linear")
        kinds = set(sched.values()) if isinstance(sched, dict) else {sched}
        if kinds != {"linear"}:
            raise NotImplementedError(f"checkpoint interpolant schedule {sched!r}: only the linear schedule is built "
                                      "(interpolant_scheduler.py:148-154)")
        if int(hp.get("n_atom_charges", 6)) != 6 or hp.get("exclude_charges", False):
            raise NotImplementedError("checkpoint needs n_atom_charges == 6 and exclude_charges == False")
        hist = None
        if load_kwargs.get("n_atoms_hist_file"):
            n, c = torch.load(load_kwargs["n_atoms_hist_file"])
            hist = (n, c.double())
        dataset = "qm9" if "qm9" in str(hp.get("n_atoms_hist_file", "")) else "geom"
        return cls(hp["atom_type_map"], hp["vector_field_config"], sd, n_atoms_hist=hist, dataset=dataset,
                   default_n_timesteps=hp.get("default_n_timesteps", 250), fake_atom_p=hp.get("fake_atom_p", 0.0),
                   explicit_aromaticity=hp.get("explicit_aromaticity", False), device=device)

    @classmethod
    def from_config(cls, name="flowmol3", dataset="geom", seed=0, device="cuda:0", vector_field=None):
        """Random-init weights of a named reference config (no checkpoint is reachable offline)."""
        from .config import NAMED_VECTOR_FIELDS
        amap = GEOM_ATOM_MAP if dataset == "geom" else QM9_ATOM_MAP
        cfg = ModelConfig.named(name, len(amap) + 1)
        return cls(amap, NAMED_VECTOR_FIELDS[name], WT.init_state_dict(cfg, seed), dataset=dataset, device=device,
                   vector_field=vector_field)

    def _materialise(self, device):
        self.vector_field = CTMCVectorFieldB200(self.cfg, self._state_dict, device=device)
        self._device = str(self.vector_field.device)

    def cuda(self, device=None):
        dev = "cuda:0" if device is None else (f"cuda:{device}" if isinstance(device, int) else str(device))
        if self.vector_field is None or str(self.vector_field.device) != str(torch.device(dev)):
            self._materialise(dev)
        return self

    def to(self, device):
        return self.cuda(device)

    def eval(self):
        return self

    # -- sampling ----------------------------------------------------------------------------------------------------------------
    def sample_n_atoms(self, n_molecules, **kwargs):
        return self.n_atoms_map[self.n_atoms_dist.sample((n_molecules,), **kwargs)]            # flowmol.py:468-471

    def sample_random_sizes(self, n_molecules, device="cuda:0", stochasticity=None, high_confidence_threshold=None,
                            xt_traj=False, ep_traj=False, **kwargs):
        return self.sample(self.sample_n_atoms(n_molecules), device=device, stochasticity=stochasticity,
                           high_confidence_threshold=high_confidence_threshold, xt_traj=xt_traj, ep_traj=ep_traj, **kwargs)

    def sample_prior(self, g):
        """flowmol.py:417-448 with the CTMC priors: COM-free N(0,1) positions (priors.py:27-35, `std` ignored there too),
        all-mask categorical state (priors.py:101-107,305-316)."""
        N, E = g.num_nodes(), g.num_edges()
        nbi = g.node_batch_idx().cpu()
        x = torch.randn(N, 3)
        com = torch.zeros(g.batch_size, 3).index_add_(0, nbi, x) / g.n_atoms[:, None].float()
        g.ndata['x_0'] = (x - com[nbi]).to(g.device)
        g.ndata['a_0'] = one_hot(torch.full((N,), self.n_atom_types), self.n_atom_types + 1).float().to(g.device)
        g.ndata['c_0'] = one_hot(torch.full((N,), 6), 7).float().to(g.device)
        g.edata['e_0'] = one_hot(torch.full((E,), self.n_bond_types), self.n_bond_types + 1).float().to(g.device)
        return g

    def prior_from_x0(self, n_atoms, x_0):
        """The `prior=` dict of `sample` (flowmol.py:532-545) for given COM-free positions and the all-mask CTMC state --
        lets a caller that shards one batch over several GPUs draw the positions ONCE for the whole batch and hand every rank
        its slice (flowmol_b200/cli.py)."""
        n = torch.as_tensor(n_atoms).long()
        N, E = int(n.sum()), int((n * (n - 1)).sum())
        return {'x_0': x_0.float(), 'fake_atoms': self.fake_atoms,
                'a_0': one_hot(torch.full((N,), self.n_atom_types), self.n_atom_types + 1).float(),
                'c_0': one_hot(torch.full((N,), 6), 7).float(),
                'e_0': one_hot(torch.full((E,), self.n_bond_types), self.n_bond_types + 1).float()}

    @staticmethod
    def centered_normal(n_atoms, generator=None):
        """COM-free N(0,1) positions for molecules of the given sizes (priors.py:27-35), [sum n, 3] on the CPU."""
        n = torch.as_tensor(n_atoms).long()
        x = torch.randn(int(n.sum()), 3, generator=generator)
        nbi = torch.arange(len(n)).repeat_interleave(n)
        return x - (torch.zeros(len(n), 3).index_add_(0, nbi, x) / n[:, None].float())[nbi]

    device_decode = True       # sample(): token-level integrate + fm_decode on the device (False: the graph-level seam + host decode)

    def _sample_tokens(self, g, n_timesteps, stochasticity, high_confidence_threshold, seed=None, mol_id_offset=0, **kwargs):
        """The body of `sample` without one-hot round trips: the prior's tokens (argmax of the one-hots, on the device), the
        token-level trajectory, the decode kernel, one D2H of compact arrays, then one SampledMolecule per molecule by slicing."""
        vf = self.vector_field
        uem = g.upper_edge_mask()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # like vector_field.integrate: torch's global RNG decides
        e0 = g.edata['e_0']
        out = vf.integrate_tokens(g.n_atoms.numpy(), g.ndata['x_0'], g.ndata['a_0'].argmax(-1), g.ndata['c_0'].argmax(-1),
                                  e0[uem.to(e0.device)].argmax(-1), n_timesteps, seed, stochasticity, high_confidence_threshold,
                                  mol_id_offset, kwargs.pop('tspan', None), kwargs.pop('cuda_graph', False),
                                  dfm_type=kwargs.pop('dfm_type', None), cat_temp_func=kwargs.pop('cat_temp_func', None),
                                  forward_weight_func=kwargs.pop('forward_weight_func', None), inv_temp_func=kwargs.pop('inv_temp_func', None))
        fake_token = len(self.atom_type_map) if self.fake_atoms else -1
        d = vf.decode_tokens(g.n_atoms.numpy(), out['x'], out['a'], out['c'], out['e'], fake_token)
        n = g.n_atoms.numpy()
        noff = np.concatenate([[0], np.cumsum(n)])
        uoff = np.concatenate([[0], np.cumsum(n * (n - 1) // 2)])
        keep = d['atom_new'] >= 0
        x_kept, a_kept, ch_kept = d['x'][keep], d['a'][keep], d['charge'][keep]
        koff = np.concatenate([[0], np.cumsum(d['mol_kept'].numpy())])
        nbonds = d['mol_bonds'].numpy()
        mols = []
        for b in range(len(n)):
            ks, bs = slice(koff[b], koff[b + 1]), slice(uoff[b], uoff[b] + nbonds[b])
            mols.append(SampledMolecule.from_decoded(x_kept[ks], a_kept[ks].tolist(), ch_kept[ks], d['bond_type'][bs], d['bond_src'][bs],
                                                     d['bond_dst'][bs], self.atom_type_map, fake_atoms=self.fake_atoms,
                                                     explicit_aromaticity=self.explicit_aromaticity))
        return mols

    @torch.no_grad()
    def sample(self, n_atoms, n_timesteps=None, device="cuda:0", stochasticity=None, high_confidence_threshold=None,
               xt_traj=False, ep_traj=False, prior=None, **kwargs):
        if self.vector_field is None:
            self._materialise(device)
        if n_timesteps is None:
            n_timesteps = self.default_n_timesteps
        visualize = bool(xt_traj or ep_traj)                                                   # flowmol.py:497
        dev = self.vector_field.device
        g = MolGraphBatch(torch.as_tensor(n_atoms).cpu(), device=dev)
        if prior is None:
            g = self.sample_prior(g)
        else:                                                                                  # flowmol.py:532-545
            g.ndata['x_0'], g.ndata['c_0'] = prior['x_0'].to(dev), prior['c_0'].to(dev)
            g.edata['e_0'] = prior['e_0'].to(dev)
            a0 = prior['a_0'].to(dev)
            if prior['fake_atoms'] and not self.fake_atoms:
                a0 = a0[:, 1:]
            elif not prior['fake_atoms'] and self.fake_atoms:
                a0 = torch.cat([torch.zeros(a0.shape[0], 1, device=dev), a0], dim=-1)
            g.ndata['a_0'] = a0
        if not visualize and self.device_decode:
            return self._sample_tokens(g, n_timesteps, stochasticity, high_confidence_threshold, **kwargs)
        itg = self.vector_field.integrate(g, g.node_batch_idx(), upper_edge_mask=g.upper_edge_mask(), n_timesteps=n_timesteps,
                                          visualize=visualize, stochasticity=stochasticity,
                                          high_confidence_threshold=high_confidence_threshold, **kwargs)
        g, traj_frames = itg if visualize else (itg, None)                                     # flowmol.py:559-562
        g.edata['ue_mask'] = g.upper_edge_mask()
        g = g.to('cpu')                                                                        # flowmol.py:564 (device -> host)
        uem = g.edata['ue_mask']
        mols, no, eo = [], 0, 0
        a_idx = g.ndata['a_1'].argmax(-1).numpy()
        c_idx = g.ndata['c_1'].argmax(-1).numpy()
        e_idx = g.edata['e_1'].argmax(-1).numpy()
        x = g.ndata['x_1'].numpy()
        uem = uem.numpy()
        for mi, n in enumerate(g.n_atoms.tolist()):
            e = n * (n - 1)
            mols.append(SampledMolecule(x[no:no + n], a_idx[no:no + n], c_idx[no:no + n], e_idx[eo:eo + e][uem[eo:eo + e]],
                                        self.atom_type_map, fake_atoms=self.fake_atoms,
                                        explicit_aromaticity=self.explicit_aromaticity,
                                        traj_frames=traj_frames[mi] if visualize else None,
                                        build_xt_traj=xt_traj, build_ep_traj=ep_traj))
            no += n
            eo += e
        return mols


def load_pretrained(model_name="flowmol3", device="cuda:0", models_dir=None):
    """flowmol.load_pretrained: the checkpoint must already be on disk (`<models_dir>/<name>/checkpoints/last.ckpt`,
    default models_dir = $FLOWMOL_B200_MODELS or flowmol_b200/trained_models); this build has no downloader."""
    if model_name not in pretrained_model_names:
        raise ValueError(f"Model {model_name} not found. Supported models: {pretrained_model_names}")
    root = Path(models_dir or os.environ.get("FLOWMOL_B200_MODELS", _HERE / "trained_models"))
    ckpt = root / model_name / "checkpoints" / "last.ckpt"
    if not ckpt.exists():
        raise FileNotFoundError(f"{ckpt} not found. Copy the reference's trained_models/{model_name}/ directory there "
                                "(the reference fetches it with wget, flowmol/__init__.py:58-77; no network here).")
    return FlowMolB200.from_checkpoint(ckpt, device=device)
