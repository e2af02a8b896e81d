"""CTMCVectorFieldB200 -- host-side mirror of the reference's `CTMCVectorField` for the sampling path.

Same method names and argument meaning as flowmol/models/ctmc_vector_field.py (`integrate`, `forward`) and
flowmol/models/vector_field.py:212 (`forward(g, t, node_batch_idx, upper_edge_mask, apply_softmax, remove_com,
prev_dst_dict)`), so `model.vector_field = CTMCVectorFieldB200(...)` is a drop-in at the seam
flowmol/models/flowmol.py:557.  Everything numerical happens in libflowmol_b200.so (CUDA, sm_100a) through the C ABI
of include/flowmol_b200.h; PyTorch is only the container for device memory and the stream.  No autograd, no DGL
message passing, no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch
from torch.nn.functional import one_hot

from . import _lib
from . import weights as WT
from .config import ModelConfig
from .graph import MolGraphBatch, n_atoms_of


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("flowmol_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"flowmol_b200 runs on CUDA devices only, got {dev}")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def build_cat_temp_schedule(schedule, decay_max=0.8, decay_a=2):
    """ctmc_vector_field.py:71-81: 'decay' -> max * (1 - t)^a, a number -> constant, or a callable of the fp32 time tensor."""
    if schedule == 'decay':
        return lambda t: decay_max * torch.pow(1 - t, decay_a)
    if isinstance(schedule, (float, int)):
        return lambda t: schedule
    if callable(schedule):
        return schedule
    raise ValueError(f"Invalid cat_temperature_schedule: {schedule}")


def build_fw_schedule(schedule, beta_a=0.25, beta_b=0.25, beta_max=10.0):
    """ctmc_vector_field.py:83-95: 'beta' -> 1 + max * t^a * (1 - t)^b, a number -> constant, or a callable."""
    if schedule == 'beta':
        return lambda t: 1 + beta_max * torch.pow(t, beta_a) * torch.pow(1 - t, beta_b)
    if isinstance(schedule, (float, int)):
        return lambda t: schedule
    if callable(schedule):
        return schedule
    raise ValueError(f"Invalid forward_weight_schedule: {schedule}")


def _f32(v):
    """A schedule value as the fp32 number the reference's tensor arithmetic would use (python floats are cast to fp32 when they
    meet an fp32 tensor; 0-d tensors already are fp32)."""
    return float(torch.as_tensor(v, dtype=torch.float32))


class CTMCVectorFieldB200:
    canonical_feat_order = ['x', 'a', 'c', 'e']

    def __init__(self, cfg: ModelConfig, state_dict, device="cuda:0"):
        self.cfg = cfg
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        self.n_atom_types, self.n_charges, self.n_bond_types = cfg.n_atom_types, cfg.n_charges, cfg.n_bond_types
        self.mask_idxs = {'a': cfg.n_atom_types, 'c': cfg.n_charges, 'e': cfg.n_bond_types}   # ctmc_vector_field.py:64-68
        self.eta = cfg.stochasticity
        self.hc_thresh = cfg.high_confidence_threshold
        self.cat_temperature = cfg.cat_temperature
        # schedules of the reference's constructor (ctmc_vector_field.py:23-57,71-95): evaluated on the host once per step
        ex = cfg.extra
        self.dfm_type = ex.get('dfm_type', 'campbell')
        self.cat_temp_func = build_cat_temp_schedule(ex.get('cat_temperature_schedule', cfg.cat_temperature),
                                                     ex.get('cat_temp_decay_max', 0.8), ex.get('cat_temp_decay_a', 2))
        self.forward_weight_func = build_fw_schedule(ex.get('forward_weight_schedule', 'beta'), ex.get('fw_beta_a', 0.25),
                                                     ex.get('fw_beta_b', 0.25), ex.get('fw_beta_max', 10.0))
        blob, offsets = WT.pack(cfg, state_dict)
        self._blob, self._offsets = np.ascontiguousarray(blob), np.ascontiguousarray(offsets)
        if not (cfg.a_token_dim == cfg.c_token_dim == cfg.e_token_dim):
            raise NotImplementedError("token embedding dims must be equal")
        mn = cfg.message_norm
        c = _lib.FmConfig(
            n_atom_types=cfg.n_atom_types, n_charges=cfg.n_charges, n_bond_types=cfg.n_bond_types,
            n_hidden_scalars=cfg.n_hidden_scalars, n_vec_channels=cfg.n_vec_channels,
            n_hidden_edge_feats=cfg.n_hidden_edge_feats, n_cp_feats=cfg.n_cp_feats, rbf_dim=cfg.rbf_dim,
            time_embedding_dim=cfg.time_embedding_dim, token_dim=cfg.a_token_dim, n_convs=cfg.n_convs,
            n_updaters=cfg.n_updaters, convs_per_update=cfg.convs_per_update,
            separate_mol_updaters=int(cfg.separate_mol_updaters), self_conditioning=int(cfg.self_conditioning),
            use_dst_feats=int(cfg.use_dst_feats), s_dst=cfg.s_dst, v_dst=cfg.v_dst, rbf_dmax=float(cfg.rbf_dmax),
            message_norm=(0.0 if mn == 'sum' else -1.0 if mn == 'mean' else float(mn)))
        h = C.c_void_p()
        _lib.check(self.lib.fm_create(C.byref(c), self._blob.ctypes.data, self._blob.size, self._offsets.ctypes.data,
                                      self._offsets.size, self.device.index, C.byref(h)))
        self._h = h
        self._ws = None
        self._ws_key = None
        self.last_launches = 0
        # fp16x3 operands need |activation| < 65504; nobody knows the range of a given trained checkpoint.  When the device status
        # word reports an overflow, the call is repeated once with 3xTF32 operands (no range limit) and the handle stays there.
        self.auto_fallback = True

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.fm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- nn.Module-ish surface (readme.md:46-47 chains .cuda().eval()) -----------------------------------------------------
    def eval(self):
        return self

    def cuda(self, device=None):
        return self

    def to(self, device):
        if _require_cuda(device) != self.device:
            raise RuntimeError("a CTMCVectorFieldB200 is bound to the device it was created on")
        return self

    # -- workspace ---------------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _prepare(self, n_atoms):
        n = np.ascontiguousarray(np.asarray(n_atoms, dtype=np.int32).reshape(-1))
        key = n.tobytes()
        if self._ws_key != key:
            nbytes = C.c_size_t()
            _lib.check(self.lib.fm_workspace_bytes(self._h, n.ctypes.data, len(n), C.byref(nbytes)))
            if self._ws is None or self._ws.numel() < nbytes.value:
                self._ws = None
                self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.fm_batch_init(self._h, n.ctypes.data, len(n), self._ws.data_ptr(), self._ws.numel(),
                                                  self._stream()))
            self._ws_key = key
        return n

    def workspace_tensor(self, name):
        """Device view of a named workspace tensor (tests / debugging)."""
        ptr, cnt = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.fm_workspace_tensor(self._h, self._ws.data_ptr(), name.encode(), C.byref(ptr), C.byref(cnt)))
        off = ptr.value - self._ws.data_ptr()
        return self._ws[off:off + 4 * cnt.value].view(torch.float32)

    def set_option(self, name, value):
        """'conv_impl': 0 = fp32 CUDA-core message kernel, 1 = fused tcgen05 3xTF32 message kernel, 2 = wide tcgen05 pipeline
        (default for the flowmol3 dims); 'tc_prec': operand format of the wide pipeline, 1 = scaled fp16 hi/lo images
        ("fp16x3", default) or 0 = 3xTF32; 'node_impl', 'fuse_agg', 'eg_nh': see include/flowmol_b200.h."""
        _lib.check(self.lib.fm_set_option(self._h, name.encode(), int(value)))
        return self

    def get_option(self, name):
        v = C.c_int32()
        _lib.check(self.lib.fm_get_option(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    def check_status(self):
        """Read-and-clear of the device status word; raises if an activation left the fp16 operand range of the tensor-core
        linears (|x| >= 65504 with tc_prec 1) -- the results of that call are invalid."""
        if self._status_fired():
            raise RuntimeError("flowmol_b200: an activation left the fp16 operand range of the tensor-core linears; "
                               "call set_option('tc_prec', 0) (3xTF32 operands) and re-run")

    def _status_fired(self):
        with torch.cuda.device(self.device):
            return bool(self.get_option("status") & 1)

    def _guarded(self, run):
        """run() -> result; on an fp16-operand overflow: with auto_fallback, switch the handle to 3xTF32 operands (with a warning)
        and run once more; otherwise raise."""
        out = run()
        if not self._status_fired():
            return out
        if self.auto_fallback and self.get_option("tc_prec") == 1:
            import warnings
            warnings.warn("flowmol_b200: an activation left the fp16 operand range (|x| >= 65504); repeating the call with 3xTF32 "
                          "operands and keeping them for this model (set_option('tc_prec', 1) switches back)")
            self.set_option("tc_prec", 0)
            out = run()
            if not self._status_fired():
                return out
        raise RuntimeError("flowmol_b200: an activation left the fp16 operand range of the tensor-core linears; "
                           "call set_option('tc_prec', 0) (3xTF32 operands) and re-run")

    def kernel_profile(self, n_atoms, x_t, a_idx, c_idx, e_idx_upper, t=0.3, prev=None, n_forwards=1):
        """In-situ per-kernel timing of `n_forwards` network evaluations: fm_debug_kprof records a CUDA event after every launch
        on the launching stream; the differences are the launches' durations in the warm pipeline.  Returns
        {label: (launches per forward, ms per forward)} with labels read from csrc/api.cu by launch-site line."""
        import collections
        import os
        import re
        if prev is None:
            prev = self.forward_tokens(n_atoms, x_t, a_idx, c_idx, e_idx_upper, 0.0, None)
        prev = self.forward_tokens(n_atoms, x_t, a_idx, c_idx, e_idx_upper, t, prev)          # warm
        torch.cuda.synchronize(self.device)
        self.set_option("kprof", 1)
        for _ in range(n_forwards):
            prev = self.forward_tokens(n_atoms, x_t, a_idx, c_idx, e_idx_upper, t, prev)
        torch.cuda.synchronize(self.device)
        cap = 2048 * n_forwards
        lines, ms, n = (C.c_int32 * cap)(), (C.c_float * cap)(), C.c_int32()
        _lib.check(self.lib.fm_debug_kprof(self._h, lines, ms, cap, C.byref(n)))
        self.set_option("kprof", 0)
        src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "api.cu")).read().split("\n")

        def label(line):
            for back in range(0, 4):                       # the launch is on the LAUNCH_OK line or just above it
                tx = src[line - 1 - back]
                m = (re.search(r"launch_eg<D, fm::(EG_\w+), \d", tx) or re.search(r"fm::(k_\w+)<", tx)
                     or re.search(r"\b(scalar|gate|linear|sgate)\(", tx))
                if m:
                    return m.group(1) if m.group(1).startswith(("EG_", "k_")) else "node_" + m.group(1)
            return f"line{line}"
        agg, seen = collections.OrderedDict(), collections.Counter()
        for i in range(n.value):
            lab = label(lines[i])
            if lab.startswith("EG_MSG"):                   # one launch site, the three linears of a message pass in turn
                lab = ("EG_MSG0", "EG_MSG", "EG_MSGA")[seen[lines[i]] % 3]
                seen[lines[i]] += 1
            elif lab == "k_vec_b" and ms[i] < 0.1:
                lab = "k_vec_b(node rows)"
            a = agg.setdefault(lab, [0, 0.0])
            a[0] += 1
            a[1] += ms[i]
        return {k: (c / n_forwards, t_ / n_forwards) for k, (c, t_) in agg.items()}

    def time_conv_edge(self, layer=1, iters=5):
        """Mean duration (ms) of the hot kernel re-launched on the state left by the last forward (bench roofline)."""
        ms = C.c_float()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fm_time_conv_edge(self._h, self._ws.data_ptr(), int(layer), int(iters), C.byref(ms), self._stream()))
        return float(ms.value)

    def time_egemm_msg(self, layer=1, iters=5):
        """Mean duration (ms) of the 292 -> 256 message linear on the tensor cores (bench roofline, flowmol3 dims)."""
        ms = C.c_float()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fm_time_egemm_msg(self._h, self._ws.data_ptr(), int(layer), int(iters), C.byref(ms), self._stream()))
        return float(ms.value)

    def _pred_buffers(self, N, U):
        dev = self.device
        return {'x': torch.empty(N, 3, device=dev), 'a': torch.empty(N, self.n_atom_types, device=dev),
                'c': torch.empty(N, self.n_charges, device=dev), 'e': torch.empty(U, self.n_bond_types, device=dev)}

    @staticmethod
    def _pred_struct(d):
        return _lib.FmPred(x=d['x'].data_ptr(), a=d['a'].data_ptr(), c=d['c'].data_ptr(), e=d['e'].data_ptr())

    # -- token-level API (what the C ABI takes) -----------------------------------------------------------------------------
    def forward_tokens(self, n_atoms, x_t, a_idx, c_idx, e_idx_upper, t, prev=None, stop_after_conv=-1):
        """One network evaluation on token-index state.  Returns the dst dict {'x','a','c','e'} (e on upper edges)."""
        n = self._prepare(n_atoms)
        N, U = int(n.sum()), int((n.astype(np.int64) * (n - 1) // 2).sum())
        dev = self.device
        x_t = x_t.to(dev, torch.float32).contiguous()
        a = a_idx.to(dev, torch.uint8).contiguous()
        c = c_idx.to(dev, torch.uint8).contiguous()
        e = e_idx_upper.to(dev, torch.uint8).contiguous()
        assert x_t.shape == (N, 3) and a.shape == (N,) and c.shape == (N,) and e.shape == (U,)
        out = self._pred_buffers(N, U)
        pv = None
        if prev is not None:
            prev = {k: prev[k].to(dev, torch.float32).contiguous() for k in 'xace'}
            pv = self._pred_struct(prev)
        po = self._pred_struct(out)

        def run():
            with torch.cuda.device(dev):
                _lib.check(self.lib.fm_forward(self._h, self._ws.data_ptr(), x_t.data_ptr(), a.data_ptr(), c.data_ptr(),
                                               e.data_ptr(), float(t), C.byref(pv) if pv is not None else None, C.byref(po),
                                               int(stop_after_conv), self._stream()))
            self.last_launches = int(self.lib.fm_last_launch_count(self._h))
            return out
        return self._guarded(run)

    def _opts(self, n_timesteps, stochasticity, high_confidence_threshold, seed, mol_id_offset, tspan, cuda_graph,
              dfm_type=None, cat_temp_func=None, forward_weight_func=None, inv_temp_func=None):
        """FmSampleOpts + the host arrays it points to (keep the second return value alive for the duration of the call).
        Schedules are Python callables of the fp32 time tensor t_i, as in the reference (ctmc_vector_field.py:159-162,310-311);
        they are evaluated here once per step and handed over as per-step scalars."""
        if tspan is None:
            tspan = torch.linspace(0, 1, n_timesteps)                # ctmc_vector_field.py:170 (fp32, evaluated by torch)
        tt = tspan.detach().cpu().float()
        ts = np.ascontiguousarray(tt.numpy())
        dfm_type = dfm_type or self.dfm_type
        if dfm_type not in ('campbell', 'gat'):
            raise ValueError(f"Invalid dfm_type: {dfm_type}")                                 # ctmc_vector_field.py:61-62
        ctf = cat_temp_func if cat_temp_func is not None else self.cat_temp_func
        keep = [ts]
        ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        null = C.POINTER(C.c_float)()
        tau = np.array([_f32(ctf(tt[k])) for k in range(len(ts) - 1)], dtype=np.float32)
        fw_p, bw_p, it_p = null, null, null
        if dfm_type == 'gat':
            fwf = forward_weight_func if forward_weight_func is not None else self.forward_weight_func
            fws = [fwf(tt[k]) for k in range(len(ts) - 1)]
            fw = np.array([_f32(v) for v in fws], dtype=np.float32)
            bw = np.array([_f32(v - 1) for v in fws], dtype=np.float32)                        # `backward_weight = forward_weight - 1` (:491)
            keep += [fw, bw]
            fw_p, bw_p = ptr(fw), ptr(bw)
        if inv_temp_func is not None:
            it = np.array([_f32(inv_temp_func(tt[k])) for k in range(len(ts) - 1)], dtype=np.float32)
            keep.append(it)
            it_p = ptr(it)
        keep.append(tau)
        o = _lib.FmSampleOpts(
            n_timesteps=len(ts), stochasticity=float(self.eta if stochasticity is None else stochasticity),
            high_confidence_threshold=float(self.hc_thresh if high_confidence_threshold is None else high_confidence_threshold),
            cat_temperature=float(self.cat_temperature), seed=int(seed) & (2 ** 64 - 1), mol_id_offset=int(mol_id_offset),
            tspan_host=ptr(ts), use_cuda_graph=int(bool(cuda_graph)), dfm_type=1 if dfm_type == 'gat' else 0,
            tau_host=ptr(tau), fw_host=fw_p, bw_host=bw_p, inv_temp_host=it_p)
        return o, keep

    def integrate_tokens(self, n_atoms, x0, a0, c0, e0_upper, n_timesteps, seed, stochasticity=None,
                         high_confidence_threshold=None, mol_id_offset=0, tspan=None, cuda_graph=False, traj=False,
                         dfm_type=None, cat_temp_func=None, forward_weight_func=None, inv_temp_func=None):
        """Full trajectory on device-resident token state; returns final {'x','a','c','e'} (new tensors).
        traj=True also returns, under 'traj', the per-step frames the reference collects with visualize=True
        (ctmc_vector_field.py:187-202,235-255) as device tensors: 'x' f32 [T,N,3], 'a','c' u8 [T,N], 'e' u8 [T,U] (frame 0 = the
        prior) and the endpoint frames 'x_1_pred' f32 [T-1,N,3], 'a_1_pred','c_1_pred' u8 [T-1,N], 'e_1_pred' u8 [T-1,U]
        (tokens sampled by campbell_step) -- written by the step kernel itself, no host round trips."""
        n = self._prepare(n_atoms)
        dev = self.device
        x_in, a_in = x0.to(dev, torch.float32).contiguous(), a0.to(dev, torch.uint8).contiguous()
        c_in, e_in = c0.to(dev, torch.uint8).contiguous(), e0_upper.to(dev, torch.uint8).contiguous()
        o, keep = self._opts(n_timesteps, stochasticity, high_confidence_threshold, seed, mol_id_offset, tspan, cuda_graph,
                             dfm_type, cat_temp_func, forward_weight_func, inv_temp_func)
        frames, tr = None, None
        if traj:
            T, N, U = int(o.n_timesteps), x_in.shape[0], e_in.shape[0]
            u8 = dict(dtype=torch.uint8, device=dev)
            frames = {'x': torch.empty(T, N, 3, device=dev), 'a': torch.empty(T, N, **u8), 'c': torch.empty(T, N, **u8),
                      'e': torch.empty(T, U, **u8), 'x_1_pred': torch.empty(T - 1, N, 3, device=dev),
                      'a_1_pred': torch.empty(T - 1, N, **u8), 'c_1_pred': torch.empty(T - 1, N, **u8),
                      'e_1_pred': torch.empty(T - 1, U, **u8)}
            tr = _lib.FmTraj(x=frames['x'].data_ptr(), a=frames['a'].data_ptr(), c=frames['c'].data_ptr(), e=frames['e'].data_ptr(),
                             x1=frames['x_1_pred'].data_ptr(), a1=frames['a_1_pred'].data_ptr(),
                             c1=frames['c_1_pred'].data_ptr(), e1=frames['e_1_pred'].data_ptr())
        def run():
            x, a, c, e = x_in.clone(), a_in.clone(), c_in.clone(), e_in.clone()       # the trajectory updates its state in place
            with torch.cuda.device(dev):
                _lib.check(self.lib.fm_integrate_traj(self._h, self._ws.data_ptr(), x.data_ptr(), a.data_ptr(), c.data_ptr(),
                                                      e.data_ptr(), C.byref(o), C.byref(tr) if tr is not None else None,
                                                      self._stream()))
            self.last_launches = int(self.lib.fm_last_launch_count(self._h))
            return {'x': x, 'a': a, 'c': c, 'e': e}
        out = self._guarded(run)
        if traj:
            out['traj'] = frames
        return out

    def _traj_frames_per_molecule(self, n_atoms, frames):
        """Device frames -> the reference's `reshaped_traj_frames` (ctmc_vector_field.py:268-283): one dict per molecule,
        '<feat>' [T, n, K+1] / '<feat>_1_pred' [T-1, ...] one-hot floats on the CPU, 'x' / 'x_1_pred' [.., n, 3]; edge frames carry
        both directions in the reference's edge order (upper triangle, then its mirror: data_processing/utils.py:4-17)."""
        n = np.asarray(n_atoms, dtype=np.int64)
        noff = np.concatenate([[0], np.cumsum(n)])
        uoff = np.concatenate([[0], np.cumsum(n * (n - 1) // 2)])
        host = {k: v.cpu() for k, v in frames.items()}
        kdim = {'a': self.n_atom_types + 1, 'c': self.n_charges + 1, 'e': self.n_bond_types + 1}
        out = []
        for i in range(len(n)):
            d = {}
            for key in ('x', 'x_1_pred'):
                d[key] = host[key][:, noff[i]:noff[i + 1]].clone()
            for f in 'ace':
                lo, hi = (uoff[i], uoff[i + 1]) if f == 'e' else (noff[i], noff[i + 1])
                for key in (f, f + '_1_pred'):
                    oh = one_hot(host[key][:, lo:hi].long(), kdim[f]).float()
                    d[key] = torch.cat([oh, oh], dim=1) if f == 'e' else oh
            out.append(d)
        return out

    def decode_tokens(self, n_atoms, x, a, c, e_upper, fake_atom_token=-1):
        """FlowMol.sample's finalisation (flowmol.py:564-587 -> molecule_builder.py:217-265) on the device: final token state ->
        compact per-molecule arrays; ONE D2H of those.  Returns CPU tensors: 'x' [N,3], 'a' [N] tokens, 'charge' [N] int8,
        'atom_new' [N] int32 (index after dropping fake atoms, -1 = dropped), 'mol_kept' [B], 'bond_src' / 'bond_dst' [U] int32 and
        'bond_type' [U] uint8 (molecule b's bonds at [u_off[b], u_off[b] + mol_bonds[b])), 'mol_bonds' [B]."""
        n = self._prepare(n_atoms)
        dev = self.device
        N, U, B = int(n.sum()), int((n.astype(np.int64) * (n - 1) // 2).sum()), len(n)
        a = a.to(dev, torch.uint8).contiguous()
        c = c.to(dev, torch.uint8).contiguous()
        e = e_upper.to(dev, torch.uint8).contiguous()
        i32 = dict(dtype=torch.int32, device=dev)
        out = {'atom_new': torch.empty(N, **i32), 'charge': torch.empty(N, dtype=torch.int8, device=dev),
               'mol_kept': torch.empty(B, **i32), 'bond_src': torch.empty(max(U, 1), **i32), 'bond_dst': torch.empty(max(U, 1), **i32),
               'bond_type': torch.empty(max(U, 1), dtype=torch.uint8, device=dev), 'mol_bonds': torch.empty(B, **i32)}
        with torch.cuda.device(dev):
            _lib.check(self.lib.fm_decode(self._h, self._ws.data_ptr(), a.data_ptr(), c.data_ptr(), e.data_ptr(), int(fake_atom_token),
                                          out['atom_new'].data_ptr(), out['charge'].data_ptr(), out['mol_kept'].data_ptr(),
                                          out['bond_src'].data_ptr(), out['bond_dst'].data_ptr(), out['bond_type'].data_ptr(),
                                          out['mol_bonds'].data_ptr(), self._stream()))
        host = {k: v.cpu() for k, v in out.items()}
        host['x'] = x.detach().to(torch.float32).cpu()
        host['a'] = a.cpu()
        return host

    def sample_host(self, n_atoms, x0, a0, c0, e0_upper, n_timesteps, seed, stochasticity=None,
                    high_confidence_threshold=None, mol_id_offset=0, tspan=None, cuda_graph=False):
        """Host (pinned) buffers in, host buffers out: H2D + batch descriptor + trajectory + D2H inside one C call.
        x0 float32 [N,3], a0/c0 uint8 [N], e0_upper uint8 [U] are CPU tensors updated IN PLACE."""
        n = np.ascontiguousarray(np.asarray(n_atoms, dtype=np.int32).reshape(-1))
        for t_, dt_ in ((x0, torch.float32), (a0, torch.uint8), (c0, torch.uint8), (e0_upper, torch.uint8)):
            assert t_.device.type == "cpu" and t_.dtype == dt_ and t_.is_contiguous()
        nbytes = C.c_size_t()
        _lib.check(self.lib.fm_workspace_bytes(self._h, n.ctypes.data, len(n), C.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = None
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        self._ws_key = n.tobytes()
        o, ts = self._opts(n_timesteps, stochasticity, high_confidence_threshold, seed, mol_id_offset, tspan, cuda_graph)
        retry = self.auto_fallback and self.get_option("tc_prec") == 1
        saved = [t_.clone() for t_ in (x0, a0, c0, e0_upper)] if retry else None      # the call overwrites its host buffers
        with torch.cuda.device(self.device):
            rc = self.lib.fm_sample_host(self._h, n.ctypes.data, len(n), x0.data_ptr(), a0.data_ptr(), c0.data_ptr(),
                                         e0_upper.data_ptr(), C.byref(o), self._ws.data_ptr(), self._ws.numel(), self._stream())
            if rc != 0 and retry and b"fp16 operand range" in self.lib.fm_last_error():
                import warnings
                warnings.warn("flowmol_b200: an activation left the fp16 operand range (|x| >= 65504); repeating the call with 3xTF32 "
                              "operands and keeping them for this model")
                self.set_option("tc_prec", 0)
                for dst_, src_ in zip((x0, a0, c0, e0_upper), saved):
                    dst_.copy_(src_)
                rc = self.lib.fm_sample_host(self._h, n.ctypes.data, len(n), x0.data_ptr(), a0.data_ptr(), c0.data_ptr(),
                                             e0_upper.data_ptr(), C.byref(o), self._ws.data_ptr(), self._ws.numel(), self._stream())
            _lib.check(rc)
        self.last_launches = int(self.lib.fm_last_launch_count(self._h))
        return {'x': x0, 'a': a0, 'c': c0, 'e': e0_upper}

    # -- graph-level API (the reference's signatures) --------------------------------------------------------------------------
    def _tokens_from_graph(self, g, suffix):
        uem = g.upper_edge_mask() if isinstance(g, MolGraphBatch) else None
        return (g.ndata[f'a_{suffix}'].argmax(-1), g.ndata[f'c_{suffix}'].argmax(-1), g.edata[f'e_{suffix}'], uem)

    def forward(self, g, t, node_batch_idx=None, upper_edge_mask=None, apply_softmax=True, remove_com=True,
                prev_dst_dict=None):
        """EndpointVectorField.forward at sampling time (vector_field.py:212-293).  `t` is the [B] tensor of (equal) times."""
        if not (apply_softmax and remove_com):
            raise NotImplementedError("the sampling path always calls forward(apply_softmax=True, remove_com=True)")
        n_atoms = n_atoms_of(g)
        tt = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
        if not bool((tt == tt[0]).all()):
            raise NotImplementedError("all molecules share one time during sampling (ctmc_vector_field.py:320)")
        a_idx, c_idx, e_oh, uem = self._tokens_from_graph(g, 't')
        uem = upper_edge_mask if upper_edge_mask is not None else uem
        e_idx = e_oh[uem.to(e_oh.device)].argmax(-1)
        return self.forward_tokens(n_atoms, g.ndata['x_t'], a_idx, c_idx, e_idx, float(tt[0]), prev_dst_dict)

    __call__ = forward

    def integrate(self, g, node_batch_idx=None, upper_edge_mask=None, n_timesteps=None, visualize=False,
                  dfm_type='campbell', stochasticity=None, high_confidence_threshold=None, cat_temp_func=None,
                  forward_weight_func=None, tspan=None, seed=None, mol_id_offset=0, cuda_graph=False, **kwargs):
        """CTMCVectorField.integrate (ctmc_vector_field.py:145-285): reads x_0/a_0/c_0/e_0 from the graph, writes
        x_1/a_1/c_1/e_1 (and *_t); with visualize=True returns (g, per-molecule trajectory frames) like the reference.  `seed` selects the Philox noise stream (default: drawn from torch's global RNG so
        `torch.manual_seed` / seed_everything still controls reproducibility, cf. test.py:70-71)."""
        inv_temp_func = kwargs.pop('inv_temp_func', None)               # reaches step() through **kwargs in the reference (:222)
        if n_timesteps is None and tspan is None:
            raise ValueError("n_timesteps is required")
        n_atoms = n_atoms_of(g)
        uem = upper_edge_mask if upper_edge_mask is not None else g.upper_edge_mask()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        e0 = g.edata['e_0']
        uem_d = uem.to(e0.device)
        out = self.integrate_tokens(n_atoms, g.ndata['x_0'], g.ndata['a_0'].argmax(-1), g.ndata['c_0'].argmax(-1),
                                    e0[uem_d].argmax(-1), n_timesteps, seed, stochasticity, high_confidence_threshold,
                                    mol_id_offset, tspan, cuda_graph, traj=bool(visualize), dfm_type=dfm_type,
                                    cat_temp_func=cat_temp_func, forward_weight_func=forward_weight_func,
                                    inv_temp_func=inv_temp_func)
        a1 = one_hot(out['a'].long(), self.n_atom_types + 1).float()
        c1 = one_hot(out['c'].long(), self.n_charges + 1).float()
        eu = one_hot(out['e'].long(), self.n_bond_types + 1).float()
        e1 = torch.zeros(uem.shape[0], self.n_bond_types + 1, device=eu.device)
        um = uem.to(eu.device)
        e1[um] = eu
        e1[~um] = eu                                                  # both triangles carry the same state (:397-406)
        for k, val in (('x', out['x']), ('a', a1), ('c', c1)):
            g.ndata[f'{k}_t'] = val
            g.ndata[f'{k}_1'] = val
        g.edata['e_t'] = e1
        g.edata['e_1'] = e1
        if visualize:                                                 # ctmc_vector_field.py:265-283
            return g, self._traj_frames_per_molecule(n_atoms, out['traj'])
        return g


def from_named_config(name, n_atom_types, seed=0, device="cuda:0"):
    """Random-weight model with one of the embedded reference configs ('flowmol3', 'dev') -- tests and benchmarks."""
    cfg = ModelConfig.named(name, n_atom_types)
    return CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, seed), device=device)
