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


def state_dict_from_checkpoint(path, allow_unsafe_pickle=None):
    """Read a Lightning `.ckpt` ({'state_dict': {'vector_field.<name>': tensor}, 'hyper_parameters': {...}}).

    Returns (state_dict without the `vector_field.` prefix, hyper_parameters dict)."""
    import os
    import pickle
    if allow_unsafe_pickle is None:
        allow_unsafe_pickle = os.environ.get("FLOWMOL_B200_UNSAFE_CKPT", "0") == "1"

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
    # First the safe loader (tensors + plain containers + the few pathlib types Lightning hyper-parameters carry).  A checkpoint
    # that pickles other classes (e.g. from packages this image lacks) is only read with the stubbing unpickler -- which, like any
    # pickle load, executes code from the file -- when the caller opts in; stubbed classes are reported.
    import pathlib
    import warnings
    try:
        with torch.serialization.safe_globals([pathlib.PosixPath, pathlib.Path, pathlib.PurePosixPath, pathlib.PurePath]):
            ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:                    # noqa: BLE001 -- torch raises UnpicklingError / RuntimeError depending on the cause
        if not allow_unsafe_pickle:
            raise RuntimeError(f"{path}: not loadable with torch.load(weights_only=True) ({type(exc).__name__}: {exc}); pass "
                               "allow_unsafe_pickle=True (or set FLOWMOL_B200_UNSAFE_CKPT=1) only for checkpoints you trust") from exc
        stubbed = []

        class _Reporting(_Unpickler):
            def find_class(self, module, name):
                cls = super().find_class(module, name)
                if cls is _Stub:
                    stubbed.append(f"{module}.{name}")
                return cls
        _Pickle.Unpickler = _Reporting
        _Pickle.load = staticmethod(lambda f, **k: _Reporting(f, **k).load())
        ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_Pickle)
        if stubbed:
            warnings.warn(f"{path}: classes not importable here were replaced by empty stubs: {sorted(set(stubbed))}")
    sd = OrderedDict((k[len("vector_field."):], v) for k, v in ckpt["state_dict"].items() if k.startswith("vector_field."))
    hp = dict(ckpt.get("hyper_parameters", {}))
    for key, kinds in (("atom_type_map", (list, tuple)), ("vector_field_config", (dict,))):
        if key in hp and not isinstance(hp[key], kinds):
            raise TypeError(f"{path}: hyper-parameter {key!r} has type {type(hp[key]).__name__}, expected {kinds[0].__name__}")
    if "n_atoms_hist_file" in hp and not isinstance(hp["n_atoms_hist_file"], (str, os.PathLike)):
        raise TypeError(f"{path}: hyper-parameter 'n_atoms_hist_file' is not a path")
    return sd, hp


# ----------------------------------------------------------------------------------------------------------------
# packing for the CUDA library  (layout: flowmol_b200/weight_layout.py)
# ----------------------------------------------------------------------------------------------------------------
from . import weight_layout as WL


def _pad4(n):
    return (n + 3) // 4 * 4


def _pad32(n):
    return (n + 31) // 32 * 32


class Packer:
    """Collects entries of the packed blob.  `tc(idx, idx_h, w_fk, rows)` stores both operand-image formats of one matrix."""

    def tc(self, idx, idx_h, w_fk, rows_per_unit):
        self.raw(idx, tc_units(w_fk, rows_per_unit))
        self.raw(idx_h, tc_units_h16(w_fk, rows_per_unit))

    def __init__(self, n_entries):
        self.chunks = []
        self.size = 0
        self.offsets = np.full(n_entries, -1, dtype=np.int64)

    def _append(self, idx, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        pad = (-self.size) % 4
        if pad:
            self.chunks.append(np.zeros(pad, np.float32))
            self.size += pad
        assert self.offsets[idx] == -1, f"entry {idx} packed twice"
        self.offsets[idx] = self.size
        self.chunks.append(arr)
        self.size += arr.size

    def raw(self, idx, arr):
        self._append(idx, arr)

    def gemm(self, idx, w_kn):
        """w_kn [K, N] -> [pad4(K), pad32(N)] zero padded."""
        w = np.asarray(w_kn, dtype=np.float32)
        K, N = w.shape
        out = np.zeros((_pad4(K), _pad32(N)), np.float32)
        out[:K, :N] = w
        self._append(idx, out)

    def vec(self, idx, b):
        b = np.asarray(b, dtype=np.float32).reshape(-1)
        out = np.zeros(_pad32(b.size), np.float32)
        out[:b.size] = b
        self._append(idx, out)

    def blob(self):
        return np.concatenate(self.chunks)


def _np(sd, k):
    return sd[k].detach().float().cpu().numpy()


# ---- tensor-core operand images ------------------------------------------------------------------------------------------
def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) to 10 explicit mantissa bits, result in an fp32 container."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(x):
    x = np.asarray(x, dtype=np.float32)
    hi = tf32_rna(x)
    return hi, tf32_rna(x - hi)


def sw128_image(tile):
    """[rows, 32] fp32 -> the byte image of a K-major SWIZZLE_128B UMMA operand tile: row r occupies bytes
    [128 r, 128 r + 128), its 16-byte chunk c is stored at chunk position c ^ (r & 7)  (csrc/tc.cuh:sw128_off)."""
    rows = tile.shape[0]
    t = np.asarray(tile, dtype=np.float32).reshape(rows, 8, 4)
    out = np.empty_like(t)
    r = np.arange(rows)
    for c in range(8):
        out[r, c ^ (r & 7)] = t[r, c]
    return out.reshape(-1)


def tc_units(w_fk, rows_per_unit):
    """w_fk [features, K] fp32 -> concatenated units in the order the kernel consumes them:
    for k-slab s (32 columns), for m-tile m (rows_per_unit feature rows): [hi unit | lo unit], each rows_per_unit*128 B."""
    w = np.asarray(w_fk, dtype=np.float32)
    Fdim, K = w.shape
    ns, nm = (K + 31) // 32, (Fdim + rows_per_unit - 1) // rows_per_unit
    wp = np.zeros((nm * rows_per_unit, ns * 32), np.float32)
    wp[:Fdim, :K] = w
    hi, lo = split_tf32(wp)
    chunks = []
    for s_ in range(ns):
        for m in range(nm):
            sl = (slice(m * rows_per_unit, (m + 1) * rows_per_unit), slice(s_ * 32, s_ * 32 + 32))
            chunks.append(sw128_image(hi[sl]))
            chunks.append(sw128_image(lo[sl]))
    return np.concatenate(chunks)


ACT_SCALE_H16 = 1.0      # == fm::tc::ACT_SCALE_H16 (csrc/tc.cuh); fm_create refuses a blob built for another value


def sw128_image_h16(tile):
    """[rows, 64] fp16 -> byte image of a K-major SWIZZLE_128B UMMA operand tile (same 128-byte rows / 16-byte chunk
    swizzle as `sw128_image`; a chunk holds 8 halves), returned as float32 words for the blob (csrc/tc.cuh:sw128_off_h)."""
    rows = tile.shape[0]
    t = np.asarray(tile, dtype=np.float16).reshape(rows, 8, 8)
    out = np.empty_like(t)
    r = np.arange(rows)
    for c in range(8):
        out[r, c ^ (r & 7)] = t[r, c]
    return np.ascontiguousarray(out).reshape(-1).view(np.float32)


def h16_weight_scale(w):
    """Power of two s with max|w| * s in [2^13, 2^14): keeps the fp16 `lo` parts of all significant weights normal."""
    m = float(np.abs(w).max())
    return 1.0 if m == 0.0 or not np.isfinite(m) else float(2.0 ** (13 - math.floor(math.log2(m))))


def split_h16(x):
    x = np.asarray(x, dtype=np.float32)
    hi = x.astype(np.float16)
    return hi, (x - hi.astype(np.float32)).astype(np.float16)


def tc_units_h16(w_fk, rows_per_unit):
    """fp16x3 twin of `tc_units`: w_fk [features, K] fp32 -> for k-slab s (64 columns), for m-tile m: [hi unit | lo unit] of
    w * scale in fp16, each rows_per_unit*128 B, followed by 4 floats [1 / (ACT_SCALE_H16 * scale), scale, 0, 0] -- the exact
    power-of-two factor the kernel's epilogue multiplies the accumulator with."""
    w = np.asarray(w_fk, dtype=np.float32)
    Fdim, K = w.shape
    ns, nm = (K + 63) // 64, (Fdim + rows_per_unit - 1) // rows_per_unit
    scale = h16_weight_scale(w)
    wp = np.zeros((nm * rows_per_unit, ns * 64), np.float32)
    wp[:Fdim, :K] = w * np.float32(scale)
    hi, lo = split_h16(wp)
    assert np.isfinite(hi.astype(np.float32)).all()
    chunks = []
    for s_ in range(ns):
        for m in range(nm):
            sl = (slice(m * rows_per_unit, (m + 1) * rows_per_unit), slice(s_ * 64, s_ * 64 + 64))
            chunks.append(sw128_image_h16(hi[sl]))
            chunks.append(sw128_image_h16(lo[sl]))
    chunks.append(np.array([1.0 / (ACT_SCALE_H16 * scale), scale, 0.0, 0.0], np.float32))
    return np.concatenate(chunks)


def feature_perm(n):
    """perm[lf] = physical feature held by accumulator column lf.  Inside every chunk of 32: lf = 8 n + 2 t + c  ->  8 t + 2 n + c
    (n, t in 0..3): read back from tensor memory with the 16x128b shape (thread t of a quad: packed column 4 n + t), a thread's four
    words are then the physical pairs 4 t .. 4 t + 3 = one 16-byte piece of the operand image (csrc/egemm_h.cuh)."""
    lf = np.arange(n)
    c32, r = lf // 32 * 32, lf % 32
    nn, t, c = r // 8, (r % 8) // 2, r % 2
    return c32 + 8 * t + 2 * nn + c


def quad_perm(n):
    """perm[L] = physical feature held by accumulator column L when an epilogue reads tensor memory with the 16x256b shape (thread t of
    a row's quad: columns 8 n + 2 t + c): 32 (n // 4) + 8 t + 2 (n % 4) + c -- the thread then owns 8 CONSECUTIVE features of every
    group of 32 (csrc/egemm_c.cuh, QL)."""
    L = np.arange(n)
    nn, t, c = L // 8, (L % 8) // 2, L % 2
    return 32 * (nn // 4) + 8 * t + 2 * (nn % 4) + c


def _pack_gvp(P, base, sd, p, w_rows=None):
    """GVP under state_dict prefix `p` -> 6 consecutive entries starting at id `base`.
    [Wh | Wcp] are fused into one operand (they multiply the same input); `w_rows` selects/reorders the rows of the
    transposed to_feats_out weight that stay on the per-row path."""
    parts = [_np(sd, p + ".Wh")]
    if (p + ".Wcp") in sd:
        parts.append(_np(sd, p + ".Wcp"))
    P.gemm(base + 0, np.concatenate(parts, axis=1))
    P.gemm(base + 1, _np(sd, p + ".Wu"))
    w = _np(sd, p + ".to_feats_out.0.weight").T                     # [in, out]
    P.gemm(base + 2, w if w_rows is None else w[w_rows])
    P.vec(base + 3, _np(sd, p + ".to_feats_out.0.bias"))
    P.gemm(base + 4, _np(sd, p + ".scalar_to_vector_gates.weight").T)
    P.vec(base + 5, _np(sd, p + ".scalar_to_vector_gates.bias"))


def _pack_linear(P, iw, ib, sd, p):
    P.gemm(iw, _np(sd, p + ".weight").T)
    P.vec(ib, _np(sd, p + ".bias"))


def pack(cfg: ModelConfig, sd):
    """state_dict -> (blob float32[n], offsets int64[n_entries]).

    Algebraic restructuring done here (DESIGN.md "node-side folding"): the rows of the first message linear that
    multiply s_src (and, with use_dst_feats, s_dst_msg) and the rows of the first edge-update linear that multiply
    s_src / s_dst are split off; the kernels apply them once per NODE and gather the result per edge."""
    check_state_dict(cfg, sd)
    S, V, F, R, cp = cfg.n_hidden_scalars, cfg.n_vec_channels, cfg.n_hidden_edge_feats, cfg.rbf_dim, cfg.n_cp_feats
    L, NU = cfg.n_convs, cfg.n_updaters
    P = Packer(WL.n_entries(L, NU))
    g = WL.gid
    P.raw(g("EMB_A"), _np(sd, "token_embeddings.a.weight"))
    P.raw(g("EMB_C"), _np(sd, "token_embeddings.c.weight"))
    P.raw(g("EMB_E"), _np(sd, "token_embeddings.e.weight"))
    # constant tables evaluated with torch so they are bit-identical to flowmol/utils/embedding.py:5-34
    P.raw(g("RBF_MU"), torch.linspace(0.0, float(cfg.rbf_dmax), R).numpy())
    half = cfg.time_embedding_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(1000) / (half - 1)))
    P.raw(g("TIME_FREQ"), freq.numpy())
    _pack_linear(P, g("SEMB0_W"), g("SEMB0_B"), sd, "scalar_embedding.0")
    _pack_linear(P, g("SEMB2_W"), g("SEMB2_B"), sd, "scalar_embedding.2")
    P.vec(g("SEMB_LN_W"), _np(sd, "scalar_embedding.4.weight")); P.vec(g("SEMB_LN_B"), _np(sd, "scalar_embedding.4.bias"))
    _pack_linear(P, g("EEMB0_W"), g("EEMB0_B"), sd, "edge_embedding.0")
    _pack_linear(P, g("EEMB2_W"), g("EEMB2_B"), sd, "edge_embedding.2")
    P.vec(g("EEMB_LN_W"), _np(sd, "edge_embedding.4.weight")); P.vec(g("EEMB_LN_B"), _np(sd, "edge_embedding.4.bias"))
    if cfg.self_conditioning:
        p = "self_conditioning_residual_layer"
        _pack_linear(P, g("SCN0_W"), g("SCN0_B"), sd, p + ".node_residual_mlp.0")
        _pack_linear(P, g("SCN2_W"), g("SCN2_B"), sd, p + ".node_residual_mlp.2")
        _pack_linear(P, g("SCE0_W"), g("SCE0_B"), sd, p + ".edge_residual_mlp.0")
        _pack_linear(P, g("SCE2_W"), g("SCE2_B"), sd, p + ".edge_residual_mlp.2")
    _pack_linear(P, g("NHEAD0_W"), g("NHEAD0_B"), sd, "node_output_head.0")
    _pack_linear(P, g("NHEAD2_W"), g("NHEAD2_B"), sd, "node_output_head.2")
    _pack_linear(P, g("EHEAD0_W"), g("EHEAD0_B"), sd, "to_edge_logits.0")
    _pack_linear(P, g("EHEAD2_W"), g("EHEAD2_B"), sd, "to_edge_logits.2")
    if S % 128 == 0 and V <= 32:
        P.raw(g("TC_INFO"), np.array([ACT_SCALE_H16, 0.0, 0.0, 0.0], np.float32))
    sdst, vdst = cfg.s_dst, cfg.v_dst
    for l in range(L):
        p = f"conv_layers.{l}"
        c = lambda n: WL.cid(l, n)
        if cfg.use_dst_feats:
            _pack_gvp(P, c("DST_WHCP"), sd, p + ".dst_feat_msg_projection")
        # message GVP 0: input scalars are cat[s_src(S) | d(R) | ef(F) | s_dst_msg(sdst) | sh(H0+cp)]  (gvp.py:523-539)
        h0 = max(V + 1 + vdst, V)
        w0 = _np(sd, f"{p}.edge_message.0.to_feats_out.0.weight").T                       # [in, S]
        rows_edge = list(range(S, S + R + F)) + list(range(S + R + F + sdst, S + R + F + sdst + h0 + cp))
        _pack_gvp(P, c("MSG0_WHCP"), sd, f"{p}.edge_message.0", w_rows=rows_edge)
        P.gemm(c("WSRC"), w0[:S])
        P.vec(c("BSRC"), _np(sd, f"{p}.edge_message.0.to_feats_out.0.bias"))
        if cfg.use_dst_feats:
            P.gemm(c("WDST"), w0[S + R + F:S + R + F + sdst])
        _pack_gvp(P, c("MSG1_WHCP"), sd, f"{p}.edge_message.1")
        _pack_gvp(P, c("MSG2_WHCP"), sd, f"{p}.edge_message.2")
        if S % 128 == 0 and V <= 32:     # tensor-core images: features on the UMMA M axis (128-row tiles), gates 32-row units
            for i in range(3):
                wt = _np(sd, f"{p}.edge_message.{i}.to_feats_out.0.weight")            # [S, in]
                if i == 0:
                    # fp16 images: k order ef(F) | d(R) | sh, so that the edge features are whole 64-wide k-slabs (they arrive
                    # as ready-made operand images by bulk TMA, csrc/egemm_p.cuh); the 3xTF32 images keep d | ef | sh
                    rows_edge_h = rows_edge[R:R + F] + rows_edge[:R] + rows_edge[R + F:]
                    P.raw(c("MSG0_TCW"), tc_units(wt[:, rows_edge], 128))
                    P.raw(c("MSG0_TCW_H"), tc_units_h16(wt[:, rows_edge_h], 128))
                else:
                    P.tc(c(f"MSG{i}_TCW"), c(f"MSG{i}_TCW_H"), wt, 128)
                wg = _np(sd, f"{p}.edge_message.{i}.scalar_to_vector_gates.weight")
                P.tc(c(f"MSG{i}_TCG"), c(f"MSG{i}_TCG_H"), wg, 32)
                if i == 1:
                    perm = feature_perm(S)
                    P.raw(c("MSG1_TCW_HP"), tc_units_h16(wt[perm, :], 128))
                    P.raw(c("MSG1_TCG_HP"), tc_units_h16(wg[:, perm], 32))
                    P.vec(c("MSG1_BP"), _np(sd, f"{p}.edge_message.1.to_feats_out.0.bias")[perm])
        for i in range(3):
            _pack_gvp(P, c(f"UPD{i}_WHCP"), sd, f"{p}.node_update.{i}")
        if S % 128 == 0 and V <= 32 and 2 * F == S:      # node pipeline on the tensor cores (node rows through k_egemm_tc)
            for i in range(3):
                P.tc(c(f"UPD{i}_TCW"), c(f"UPD{i}_TCW_H"), _np(sd, f"{p}.node_update.{i}.to_feats_out.0.weight"), 128)
                P.tc(c(f"UPD{i}_TCG"), c(f"UPD{i}_TCG_H"), _np(sd, f"{p}.node_update.{i}.scalar_to_vector_gates.weight"), 32)
            P.tc(c("WSRC_TC"), c("WSRC_TC_H"), w0[:S].T, 128)
        P.vec(c("LN_MSG_W"), _np(sd, f"{p}.message_layer_norm.feat_norm.weight"))
        P.vec(c("LN_MSG_B"), _np(sd, f"{p}.message_layer_norm.feat_norm.bias"))
        P.vec(c("LN_UPD_W"), _np(sd, f"{p}.update_layer_norm.feat_norm.weight"))
        P.vec(c("LN_UPD_B"), _np(sd, f"{p}.update_layer_norm.feat_norm.bias"))
    for u in range(NU):
        c = lambda n: WL.uid(L, u, n)
        for i in range(3):
            _pack_gvp(P, c(f"POS{i}_WHCP"), sd, f"node_position_updaters.{u}.gvps.{i}")
        p = f"edge_updaters.{u}"
        w1 = _np(sd, p + ".edge_update_fn.0.weight").T      # [2S+F+R, F], rows: s_src | s_dst | ef | d  (vector_field.py:866-875)
        P.gemm(c("EUPD_WN"), np.concatenate([w1[:S], w1[S:2 * S]], axis=1))          # [S, 2F]: per-node EA | EB
        P.vec(c("EUPD_BN"), np.concatenate([_np(sd, p + ".edge_update_fn.0.bias"), np.zeros(F, np.float32)]))
        P.gemm(c("EUPD_WE"), w1[2 * S:])                                              # rows ef | d
        _pack_linear(P, c("EUPD_W2"), c("EUPD_B2"), sd, p + ".edge_update_fn.2")
        P.vec(c("EUPD_LN_W"), _np(sd, p + ".edge_norm.weight")); P.vec(c("EUPD_LN_B"), _np(sd, p + ".edge_norm.bias"))
        if F == 128 and S % 128 == 0 and V <= 32:        # tensor-core images of the two EdgeUpdate linears (features on M)
            P.tc(c("EUPD_TC1"), c("EUPD_TC1_H"), w1[2 * S:].T, 128)                                  # [F, F + R], k order ef | d
            w2 = _np(sd, p + ".edge_update_fn.2.weight")
            P.tc(c("EUPD_TC2"), c("EUPD_TC2_H"), w2, 128)                                            # [F, F]
            qp = quad_perm(F)
            P.raw(c("EUPD_TC1_HP"), tc_units_h16(w1[2 * S:].T[qp, :], 128))                          # hidden features permuted
            P.raw(c("EUPD_TC2_HP"), tc_units_h16(w2[qp, :][:, qp], 128))                             # k = permuted hidden, rows = permuted outputs
        if S % 128 == 0 and V <= 32 and 2 * F == S:
            for i in range(3):
                q = f"node_position_updaters.{u}.gvps.{i}"
                P.tc(c(f"POS{i}_TCW"), c(f"POS{i}_TCW_H"), _np(sd, q + ".to_feats_out.0.weight"), 128)
                P.tc(c(f"POS{i}_TCG"), c(f"POS{i}_TCG_H"), _np(sd, q + ".scalar_to_vector_gates.weight"), 32)
            P.tc(c("EUPD_WN_TC"), c("EUPD_WN_TC_H"), np.concatenate([w1[:S], w1[S:2 * S]], axis=1).T, 128)   # [2F, S]: EA | EB rows
    return P.blob(), P.offsets
