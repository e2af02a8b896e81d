"""Model / sampler hyper-parameters of the sampling hot path.

`ModelConfig.from_vector_field_block` accepts the `vector_field:` block of a reference YAML
(configs/flowmol3.yml:80-106, configs/dev.yml:78-108) or the `vector_field_config` entry of a Lightning
checkpoint's `hyper_parameters` (flowmol/models/flowmol.py:29-53,146-153,169).  The two shipped blocks are
embedded below as plain data so the package works without the reference tree.
"""
from dataclasses import dataclass, asdict, field
from typing import Union

# configs/flowmol3.yml:80-106 (values only)
FLOWMOL3_VECTOR_FIELD = dict(
    self_conditioning=True, stochasticity=30.0, high_confidence_threshold=0.9, n_vec_channels=32,
    update_edge_w_distance=True, n_hidden_scalars=256, n_hidden_edge_feats=128, s_message_dim=None,
    v_message_dim=None, n_expansion_gvps=3, attention=False, n_heads=32, n_recycles=1,
    separate_mol_updaters=True, n_molecule_updates=6, convs_per_update=1, n_cp_feats=4, n_message_gvps=3,
    n_update_gvps=3, message_norm='sum', rbf_dmax=10, rbf_dim=32, time_embedding_dim=64, a_token_dim=64,
    c_token_dim=64, e_token_dim=64)

# configs/dev.yml:78-108 (values only)
DEV_VECTOR_FIELD = dict(
    self_conditioning=True, stochasticity=20.0, high_confidence_threshold=0.9, update_edge_w_distance=True,
    n_vec_channels=16, n_hidden_scalars=64, n_hidden_edge_feats=64, s_message_dim=None, v_message_dim=None,
    n_expansion_gvps=2, use_dst_feats=True, dst_feat_msg_reduction_factor=4, attention=False, n_heads=4,
    n_recycles=1, n_molecule_updates=3, separate_mol_updaters=True, convs_per_update=1, n_cp_feats=4,
    n_message_gvps=3, n_update_gvps=3, message_norm='sum', rbf_dmax=10, rbf_dim=32, time_embedding_dim=64,
    a_token_dim=64, c_token_dim=64, e_token_dim=64, dropout=0.0)

# dataset.atom_map of the YAMLs (configs/flowmol3.yml:40); QM9's has no YAML in the tree
GEOM_ATOM_MAP = ['C', 'H', 'N', 'O', 'F', 'P', 'S', 'Cl', 'Br', 'I']
QM9_ATOM_MAP = ['C', 'H', 'N', 'O', 'F']

NAMED_VECTOR_FIELDS = {'flowmol3': FLOWMOL3_VECTOR_FIELD, 'dev': DEV_VECTOR_FIELD}


@dataclass
class ModelConfig:
    n_atom_types: int                 # incl. the fake-atom type when fake atoms are on; mask token index == n_atom_types
    n_charges: int = 6
    n_bond_types: int = 4
    n_vec_channels: int = 16
    n_hidden_scalars: int = 64
    n_hidden_edge_feats: int = 64
    n_cp_feats: int = 0
    n_molecule_updates: int = 2
    convs_per_update: int = 2
    separate_mol_updaters: bool = False
    n_message_gvps: int = 3
    n_update_gvps: int = 3
    message_norm: Union[float, str] = 100
    update_edge_w_distance: bool = False
    rbf_dmax: float = 20
    rbf_dim: int = 16
    time_embedding_dim: int = 1
    a_token_dim: int = 0
    c_token_dim: int = 0
    e_token_dim: int = 0
    self_conditioning: bool = False
    use_dst_feats: bool = False
    dst_feat_msg_reduction_factor: float = 4
    stochasticity: float = 0.0
    high_confidence_threshold: float = 0.0
    cat_temperature: float = 0.05     # CTMCVectorField default cat_temperature_schedule (ctmc_vector_field.py:27)
    extra: dict = field(default_factory=dict)

    @property
    def n_convs(self):
        return self.convs_per_update * self.n_molecule_updates

    @property
    def n_updaters(self):
        return self.n_molecule_updates if self.separate_mol_updaters else 1

    @property
    def s_dst(self):
        return int(self.n_hidden_scalars / self.dst_feat_msg_reduction_factor) if self.use_dst_feats else 0

    @property
    def v_dst(self):
        return int(self.n_vec_channels / self.dst_feat_msg_reduction_factor) if self.use_dst_feats else 0

    @classmethod
    def from_vector_field_block(cls, block, n_atom_types, n_charges=6, n_bond_types=4):
        known = {k for k in cls.__dataclass_fields__ if k != 'extra'}
        kw = {k: v for k, v in block.items() if k in known}
        extra = {k: v for k, v in block.items() if k not in known}
        cfg = cls(n_atom_types=n_atom_types, n_charges=n_charges, n_bond_types=n_bond_types, extra=extra, **kw)
        cfg.validate()
        return cfg

    @classmethod
    def named(cls, name, n_atom_types):
        return cls.from_vector_field_block(NAMED_VECTOR_FIELDS[name], n_atom_types)

    def validate(self):
        """Reject reference options that are off the live sampling path (SURVEY.md headline fact 3, section 8f-4)."""
        ex = self.extra
        if ex.get('attention', False):
            raise NotImplementedError("attention=True is not used by any shipped config")
        if ex.get('s_message_dim') is not None or ex.get('v_message_dim') is not None:
            raise NotImplementedError("compressed messaging (s_message_dim/v_message_dim) is not supported")
        if ex.get('n_recycles', 1) != 1:
            raise NotImplementedError("n_recycles != 1 is not supported")
        if self.a_token_dim == 0 or self.c_token_dim == 0 or self.e_token_dim == 0:
            raise NotImplementedError("CTMC sampling needs token embeddings (a/c/e_token_dim > 0)")
        if self.time_embedding_dim < 4 or self.time_embedding_dim % 2:
            raise NotImplementedError("time_embedding_dim must be even and >= 4")
        if isinstance(self.message_norm, str) and self.message_norm not in ('sum', 'mean'):
            raise ValueError(f"message_norm must be 'mean', 'sum' or a number, got {self.message_norm}")
        if self.n_vec_channels < 3:
            raise ValueError("n_vec_channels must be >= 3")          # vector_field.py:88
        if self.n_message_gvps != 3 or self.n_update_gvps != 3:
            raise NotImplementedError("kernels are specialised for 3 message / 3 update GVPs (all shipped configs)")
        if not self.update_edge_w_distance:
            raise NotImplementedError("update_edge_w_distance=False is not used by any shipped config")

    def to_dict(self):
        return asdict(self)
