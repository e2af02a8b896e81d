"""Shared helpers for the parity tests: golden loading, oracle construction, fixture -> inputs."""
import os

import numpy as np
import torch

from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from oracle import flowmol_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def model_from_golden(gd, dtype=torch.float32):
    cfg = ModelConfig.named(str(gd["config"]), int(gd["n_atom_types"]))
    sd = WT.init_state_dict(cfg, seed=int(gd["weight_seed"]))
    chk = WT.weights_checksum(sd)
    assert abs(chk - float(gd["weights_checksum"])) <= 1e-9 * max(1.0, abs(chk)), "weights differ from the golden run"
    return cfg, sd, O.OracleModel(cfg, sd, dtype=dtype)


def t(a, dtype=None):
    x = torch.from_numpy(np.asarray(a))
    return x if dtype is None else x.to(dtype)


def golden_schedules(gd):
    """(cat_temp_func, forward_weight_func) of a golden generated with non-default schedules (oracle/make_golden.py:GAT_SCHEDULES),
    rebuilt with the product's host-side builders (flowmol_b200/vector_field.py == ctmc_vector_field.py:71-95)."""
    from flowmol_b200.vector_field import build_cat_temp_schedule, build_fw_schedule
    sched = str(gd["cat_temperature_schedule"])
    ctf = build_cat_temp_schedule(sched if sched == "decay" else float(sched), float(gd["cat_temp_decay_max"]),
                                  float(gd["cat_temp_decay_a"]))
    fwf = build_fw_schedule(str(gd["forward_weight_schedule"]), float(gd["fw_beta_a"]), float(gd["fw_beta_b"]),
                            float(gd["fw_beta_max"]))
    return ctf, fwf
