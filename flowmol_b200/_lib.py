"""ctypes binding of include/flowmol_b200.h -- the only way Python reaches the kernels.

There is NO fallback: if the shared library is missing or no CUDA device is present the product path raises.
"""
import ctypes as C
import os

from . import build as _build

c_i32, c_f32, c_u64, c_vp = C.c_int32, C.c_float, C.c_uint64, C.c_void_p


class FmConfig(C.Structure):
    _fields_ = [(n, c_i32) for n in (
        "n_atom_types", "n_charges", "n_bond_types", "n_hidden_scalars", "n_vec_channels", "n_hidden_edge_feats",
        "n_cp_feats", "rbf_dim", "time_embedding_dim", "token_dim", "n_convs", "n_updaters", "convs_per_update",
        "separate_mol_updaters", "self_conditioning", "use_dst_feats", "s_dst", "v_dst")] + [
        ("rbf_dmax", c_f32), ("message_norm", c_f32)]


class FmPred(C.Structure):
    _fields_ = [("x", c_vp), ("a", c_vp), ("c", c_vp), ("e", c_vp)]


class FmTraj(C.Structure):
    _fields_ = [("x", c_vp), ("a", c_vp), ("c", c_vp), ("e", c_vp), ("x1", c_vp), ("a1", c_vp), ("c1", c_vp), ("e1", c_vp)]


class FmSampleOpts(C.Structure):
    _fields_ = [("n_timesteps", c_i32), ("stochasticity", c_f32), ("high_confidence_threshold", c_f32),
                ("cat_temperature", c_f32), ("seed", c_u64), ("mol_id_offset", c_i32),
                ("tspan_host", C.POINTER(c_f32)), ("use_cuda_graph", c_i32), ("dfm_type", c_i32),
                ("tau_host", C.POINTER(c_f32)), ("fw_host", C.POINTER(c_f32)), ("bw_host", C.POINTER(c_f32)),
                ("inv_temp_host", C.POINTER(c_f32))]


# every symbol include/flowmol_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "fm_abi_version": (c_i32, []),
    "fm_last_error": (C.c_char_p, []),
    "fm_create": (c_i32, [C.POINTER(FmConfig), c_vp, C.c_size_t, c_vp, C.c_size_t, c_i32, C.POINTER(c_vp)]),
    "fm_destroy": (None, [c_vp]),
    "fm_workspace_bytes": (c_i32, [c_vp, c_vp, c_i32, C.POINTER(C.c_size_t)]),
    "fm_batch_init": (c_i32, [c_vp, c_vp, c_i32, c_vp, C.c_size_t, c_vp]),
    "fm_forward": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, C.POINTER(FmPred), C.POINTER(FmPred), c_i32, c_vp]),
    "fm_integrate": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(FmSampleOpts), c_vp]),
    "fm_integrate_traj": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(FmSampleOpts), C.POINTER(FmTraj), c_vp]),
    "fm_sample_host": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, C.POINTER(FmSampleOpts), c_vp, C.c_size_t, c_vp]),
    "fm_workspace_tensor": (c_i32, [c_vp, c_vp, C.c_char_p, C.POINTER(c_vp), C.POINTER(C.c_size_t)]),
    "fm_debug_time_grid": (None, [c_i32, c_vp]),
    "fm_last_launch_count": (C.c_int64, [c_vp]),
    "fm_set_option": (c_i32, [c_vp, C.c_char_p, c_i32]),
    "fm_get_option": (c_i32, [c_vp, C.c_char_p, C.POINTER(c_i32)]),
    "fm_debug_read_trace": (c_i32, [c_vp, c_vp]),
    "fm_debug_kprof": (c_i32, [c_vp, c_vp, c_vp, c_i32, C.POINTER(c_i32)]),
    "fm_debug_tc_gemm": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i32]),
    "fm_time_egemm_msg": (c_i32, [c_vp, c_vp, c_i32, c_i32, C.POINTER(c_f32), c_vp]),
    "fm_time_conv_edge": (c_i32, [c_vp, c_vp, c_i32, c_i32, C.POINTER(c_f32), c_vp]),
    "fm_decode": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "fm_debug_tmem_shapes": (c_i32, [c_vp, c_i32]),
    "fm_debug_ctmc_step": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, c_f32, c_f32, c_i32, c_i32]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load libflowmol_b200.so (built in-tree by flowmol_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m flowmol_b200.build` (nvcc, sm_100a). "
                           "flowmol_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here == header / library drift
        fn.restype = res
        fn.argtypes = args
    if lib.fm_abi_version() != 2:
        raise RuntimeError("libflowmol_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("flowmol_b200: " + load().fm_last_error().decode())
