"""Weights of the vector field: reference-named state_dict <-> one packed fp32 blob for the CUDA library.

* `expected_tensors(cfg)` -- the (name, shape) list of `CTMCVectorField(**cfg).state_dict()`
  (flowmol/models/vector_field.py:16-197, gvp.py:30-88,188-433, self_conditioning.py:9-35); `nn.Linear.weight` is
  [out, in], GVP `Wh/Wcp/Wu` are [in, out].
* `init_state_dict(cfg, seed)` -- random weights with the reference constructors' distributions (GVP uniform init
  gvp.py:53-70, nn.Linear / nn.Embedding / nn.LayerNorm defaults), drawn from a CPU torch.Generator so the same seed
  gives the same weights on every box.  No pretrained checkpoint is reachable offline, so tests and
  benchmarks use these.
* `state_dict_from_checkpoint(path)` -- `vector_field.*` tensors + hyper-parameters of a Lightning `.ckpt`.
* `pack(cfg, state_dict)` -- the device layout documented in DESIGN.md ("packed weights"): every tensor the kernels read,
  transposed to [in, out] (K-major rows, out contiguous), out padded to a multiple of 4 floats, with the node-side
  slices of the first message / edge-update linears split out (they are applied per node, not per edge).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from .config import ModelConfig


def _gvp_tensors(prefix, v_in, v_out, s_in, s_out, n_cp):
    h = max(v_in, v_out)
    out = [(f"{prefix}.Wh", (v_in, h))]
    if n_cp > 0:
        out.append((f"{prefix}.Wcp", (v_in, 2 * n_cp)))
    out += [(f"{prefix}.Wu", (h + n_cp, v_out)),
            (f"{prefix}.to_feats_out.0.weight", (s_out, h + n_cp + s_in)),
            (f"{prefix}.to_feats_out.0.bias", (s_out,)),
            (f"{prefix}.scalar_to_vector_gates.weight", (v_out, s_out)),
            (f"{prefix}.scalar_to_vector_gates.bias", (v_out,))]
    return out


def _linear(prefix, n_in, n_out):
    return [(f"{prefix}.weight", (n_out, n_in)), (f"{prefix}.bias", (n_out,))]


def _norm(prefix, n):
    return [(f"{prefix}.weight", (n,)), (f"{prefix}.bias", (n,))]


def expected_tensors(cfg: ModelConfig):
    S, V, F, R, cp = cfg.n_hidden_scalars, cfg.n_vec_channels, cfg.n_hidden_edge_feats, cfg.rbf_dim, cfg.n_cp_feats
    A, C, E = cfg.n_atom_types, cfg.n_charges, cfg.n_bond_types
    t = [("token_embeddings.a.weight", (A + 1, cfg.a_token_dim)),
         ("token_embeddings.c.weight", (C + 1, cfg.c_token_dim)),
         ("token_embeddings.e.weight", (E + 1, cfg.e_token_dim))]
    t += _linear("scalar_embedding.0", cfg.a_token_dim + cfg.c_token_dim + cfg.time_embedding_dim, S)
    t += _linear("scalar_embedding.2", S, S) + _norm("scalar_embedding.4", S)
    t += _linear("edge_embedding.0", cfg.e_token_dim, F) + _linear("edge_embedding.2", F, F) + _norm("edge_embedding.4", F)
    for l in range(cfg.n_convs):
        p = f"conv_layers.{l}"
        if cfg.use_dst_feats:
            t += _gvp_tensors(f"{p}.dst_feat_msg_projection", V, cfg.v_dst, S, cfg.s_dst, 0)
        t += _gvp_tensors(f"{p}.edge_message.0", V + 1 + cfg.v_dst, V, S + R + F + cfg.s_dst, S, cp)
        t += _gvp_tensors(f"{p}.edge_message.1", V, V, S, S, cp)
        t += _gvp_tensors(f"{p}.edge_message.2", V, V, S, S, cp)
        for i in range(3):
            t += _gvp_tensors(f"{p}.node_update.{i}", V, V, S, S, cp)
        t += _norm(f"{p}.message_layer_norm.feat_norm", S) + _norm(f"{p}.update_layer_norm.feat_norm", S)
    for u in range(cfg.n_updaters):
        p = f"node_position_updaters.{u}.gvps"
        t += _gvp_tensors(f"{p}.0", V, V, S, S, cp) + _gvp_tensors(f"{p}.1", V, V, S, S, cp)
        t += _gvp_tensors(f"{p}.2", V, 1, S, S, cp)
    for u in range(cfg.n_updaters):
        p = f"edge_updaters.{u}"
        t += _linear(f"{p}.edge_update_fn.0", 2 * S + F + (R if cfg.update_edge_w_distance else 0), F)
        t += _linear(f"{p}.edge_update_fn.2", F, F) + _norm(f"{p}.edge_norm", F)
    t += _linear("node_output_head.0", S, S) + _linear("node_output_head.2", S, A + C)
    t += _linear("to_edge_logits.0", F, F) + _linear("to_edge_logits.2", F, E)
    if cfg.self_conditioning:
        p = "self_conditioning_residual_layer"
        t += _linear(f"{p}.node_residual_mlp.0", S + A + C + R, S) + _linear(f"{p}.node_residual_mlp.2", S, S)
        t += _linear(f"{p}.edge_residual_mlp.0", F + E + R, F) + _linear(f"{p}.edge_residual_mlp.2", F, F)
    return t


def init_state_dict(cfg: ModelConfig, seed: int = 0):
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    sd = OrderedDict()
    shapes = dict(expected_tensors(cfg))
    for name, shape in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf in ("Wh", "Wcp", "Wu"):
            k = 1.0 / math.sqrt(shape[0])
            w = torch.empty(shape).uniform_(-k, k, generator=g)
        elif name.startswith("token_embeddings"):
            w = torch.empty(shape).normal_(0.0, 1.0, generator=g)
        elif "norm" in name or name.startswith(("scalar_embedding.4", "edge_embedding.4")):
            w = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        elif leaf == "weight":
            k = 1.0 / math.sqrt(shape[1])
            w = torch.empty(shape).uniform_(-k, k, generator=g)
        else:  # bias of a Linear: bound 1/sqrt(fan_in) of the matching weight
            fan_in = shapes[name[:-4] + "weight"][1]
            k = 1.0 / math.sqrt(fan_in)
            w = torch.empty(shape).uniform_(-k, k, generator=g)
        sd[name] = w
    return sd


def check_state_dict(cfg: ModelConfig, sd):
    exp = expected_tensors(cfg)
    missing = [n for n, _ in exp if n not in sd]
    if missing:
        raise KeyError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
    for n, shape in exp:
        if tuple(sd[n].shape) != tuple(shape):
            raise ValueError(f"{n}: expected shape {tuple(shape)}, got {tuple(sd[n].shape)}")


def weights_checksum(sd):
    """Order-independent fingerprint used by the golden fixtures to make sure both sides hold the same weights."""
    acc = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        acc += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64).reshape(v.shape).remainder(7.0).add(1.0)).sum())
    return acc


class _Stub:
    """Placeholder for classes pickled into a Lightning checkpoint whose packages are not installed."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


def state_dict_from_checkpoint(path):
    """Read a Lightning `.ckpt` ({'state_dict': {'vector_field.<name>': tensor}, 'hyper_parameters': {...}}).

    Returns (state_dict without the `vector_field.` prefix, hyper_parameters dict)."""
    import pickle

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                return _Stub

    class _Pickle:
        Unpickler = _Unpickler
        __name__ = "pickle"
        load = staticmethod(lambda f, **k: _Unpickler(f, **k).load())

    for attr in dir(pickle):
        if not hasattr(_Pickle, attr):
            setattr(_Pickle, attr, getattr(pickle, attr))
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Pickle)
    sd = OrderedDict((k[len("vector_field."):], v) for k, v in ckpt["state_dict"].items() if k.startswith("vector_field."))
    hp = dict(ckpt.get("hyper_parameters", {}))
    return sd, hp


# ----------------------------------------------------------------------------------------------------------------
# packing for the CUDA library
# ----------------------------------------------------------------------------------------------------------------
def _pad4(n):
    return (n + 3) // 4 * 4


class Packer:
    """Appends [K, Npad] row-major fp32 matrices / vectors to one blob; every entry starts 16-byte aligned."""

    def __init__(self):
        self.chunks = []
        self.size = 0
        self.index = OrderedDict()

    def add(self, name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        pad = (-self.size) % 4
        if pad:
            self.chunks.append(np.zeros(pad, np.float32))
            self.size += pad
        self.index[name] = (self.size, arr.size)
        self.chunks.append(arr)
        self.size += arr.size

    def add_matrix_kn(self, name, w_kn):
        """w_kn: [K, N] -> stored [K, pad4(N)] (zero padded columns)."""
        w = np.asarray(w_kn, dtype=np.float32)
        K, N = w.shape
        out = np.zeros((K, _pad4(N)), np.float32)
        out[:, :N] = w
        self.add(name, out)

    def blob(self):
        return np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.float32)


def _gvp_pack(P, name, sd, p, n_node_rows=0):
    """One GVP: [Wh | Wcp] fused as a single [v_in, h+2cp] matrix, Wu, W (transposed, optional node-side rows split
    off: rows [0, n_node_rows) of the transposed to_feats_out weight multiply per-node inputs), b, Wg (transposed), bg."""
    wh = sd[p + ".Wh"].numpy()
    parts = [wh]
    if (p + ".Wcp") in sd:
        parts.append(sd[p + ".Wcp"].numpy())
    P.add_matrix_kn(name + ".whcp", np.concatenate(parts, axis=1))
    P.add_matrix_kn(name + ".wu", sd[p + ".Wu"].numpy())
    w = sd[p + ".to_feats_out.0.weight"].numpy().T            # [in, out]
    P.add_matrix_kn(name + ".w", w)
    P.add(name + ".b", sd[p + ".to_feats_out.0.bias"].numpy())
    P.add_matrix_kn(name + ".wg", sd[p + ".scalar_to_vector_gates.weight"].numpy().T)
    P.add(name + ".bg", sd[p + ".scalar_to_vector_gates.bias"].numpy())


def _lin_pack(P, name, sd, p):
    P.add_matrix_kn(name + ".w", sd[p + ".weight"].numpy().T)
    P.add(name + ".b", sd[p + ".bias"].numpy())


def pack(cfg: ModelConfig, sd):
    """Returns (blob float32[n], index {name: (offset, size)}) in the fixed order the C library expects
    (flowmol_b200/csrc/weights.h walks the same sequence; a layout hash guards against drift)."""
    check_state_dict(cfg, sd)
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    P = Packer()
    P.add("emb.a", sd["token_embeddings.a.weight"].numpy())
    P.add("emb.c", sd["token_embeddings.c.weight"].numpy())
    P.add("emb.e", sd["token_embeddings.e.weight"].numpy())
    _lin_pack(P, "semb.0", sd, "scalar_embedding.0")
    _lin_pack(P, "semb.2", sd, "scalar_embedding.2")
    P.add("semb.ln.w", sd["scalar_embedding.4.weight"].numpy()); P.add("semb.ln.b", sd["scalar_embedding.4.bias"].numpy())
    _lin_pack(P, "eemb.0", sd, "edge_embedding.0")
    _lin_pack(P, "eemb.2", sd, "edge_embedding.2")
    P.add("eemb.ln.w", sd["edge_embedding.4.weight"].numpy()); P.add("eemb.ln.b", sd["edge_embedding.4.bias"].numpy())
    if cfg.self_conditioning:
        p = "self_conditioning_residual_layer"
        _lin_pack(P, "sc.n0", sd, p + ".node_residual_mlp.0"); _lin_pack(P, "sc.n2", sd, p + ".node_residual_mlp.2")
        _lin_pack(P, "sc.e0", sd, p + ".edge_residual_mlp.0"); _lin_pack(P, "sc.e2", sd, p + ".edge_residual_mlp.2")
    for l in range(cfg.n_convs):
        p = f"conv_layers.{l}"
        if cfg.use_dst_feats:
            _gvp_pack(P, f"conv{l}.dst", sd, p + ".dst_feat_msg_projection")
        for i in range(3):
            _gvp_pack(P, f"conv{l}.msg{i}", sd, f"{p}.edge_message.{i}")
        for i in range(3):
            _gvp_pack(P, f"conv{l}.upd{i}", sd, f"{p}.node_update.{i}")
        for nm, q in (("ln_msg", "message_layer_norm"), ("ln_upd", "update_layer_norm")):
            P.add(f"conv{l}.{nm}.w", sd[f"{p}.{q}.feat_norm.weight"].numpy())
            P.add(f"conv{l}.{nm}.b", sd[f"{p}.{q}.feat_norm.bias"].numpy())
    for u in range(cfg.n_updaters):
        for i in range(3):
            _gvp_pack(P, f"pos{u}.gvp{i}", sd, f"node_position_updaters.{u}.gvps.{i}")
    for u in range(cfg.n_updaters):
        p = f"edge_updaters.{u}"
        _lin_pack(P, f"eupd{u}.0", sd, p + ".edge_update_fn.0")
        _lin_pack(P, f"eupd{u}.2", sd, p + ".edge_update_fn.2")
        P.add(f"eupd{u}.ln.w", sd[p + ".edge_norm.weight"].numpy()); P.add(f"eupd{u}.ln.b", sd[p + ".edge_norm.bias"].numpy())
    _lin_pack(P, "nhead.0", sd, "node_output_head.0"); _lin_pack(P, "nhead.2", sd, "node_output_head.2")
    _lin_pack(P, "ehead.0", sd, "to_edge_logits.0"); _lin_pack(P, "ehead.2", sd, "to_edge_logits.2")
    return P.blob(), P.index
