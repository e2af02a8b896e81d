"""First-contact diagnostics on the GPU box: run every parity comparison, never assert, dump the error table.

    python tools/gpu_diag.py  > gpurun_out/diag.txt
"""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmol_b200 import graph as G          # noqa: E402
from flowmol_b200 import weights as WT       # noqa: E402
from flowmol_b200.config import ModelConfig  # noqa: E402
from oracle import flowmol_oracle as O       # noqa: E402
from tests.helpers import load_golden, t     # noqa: E402


def err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()), float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    from flowmol_b200.vector_field import CTMCVectorFieldB200
    print(torch.cuda.get_device_name(0))
    rows = []
    for name in ("fwd_dev_taps", "fwd_flowmol3_taps", "fwd_flowmol3_geom", "fwd_dev_qm9"):
        try:
            gd = load_golden(name)
            cfg = ModelConfig.named(str(gd["config"]), int(gd["n_atom_types"]))
            sd = WT.init_state_dict(cfg, int(gd["weight_seed"]))
            vf = CTMCVectorFieldB200(cfg, sd)
            om64 = O.OracleModel(cfg, sd, dtype=torch.float64)
            n_atoms = gd["n_atoms"]
            N = int(n_atoms.sum())
            bt = O.make_batch(n_atoms)
            perm = torch.from_numpy(G.ref_edge_to_internal(n_atoms))
            prev = {k: t(gd[f"c0.out.{k}"]) for k in "xace"}
            args = (n_atoms, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]), prev)
            if "c1.tap.conv0.s" in gd:
                for l in range(cfg.n_convs):
                    vf.forward_tokens(*args, stop_after_conv=l)
                    torch.cuda.synchronize()
                    s = vf.workspace_tensor("s").view(N, -1).cpu().numpy()
                    v = vf.workspace_tensor("v").view(N, 3, -1).permute(0, 2, 1).cpu().numpy()
                    print(name, f"conv{l} s", err(s, gd[f"c1.tap.conv{l}.s"]), "v", err(v, gd[f"c1.tap.conv{l}.v"]))
                    if l >= 1:
                        x = vf.workspace_tensor("x").view(N, 3).cpu().numpy()
                        ef = vf.workspace_tensor("ef").view(-1, cfg.n_hidden_edge_feats).cpu()[perm].numpy()
                        print(name, f"   pos{l}", err(x, gd[f"c1.tap.pos{l}"]), "eupd", err(ef, gd[f"c1.tap.eupd{l}"]))
            d0 = vf.forward_tokens(n_atoms, t(gd["c0.x_t"]), t(gd["c0.a"]), t(gd["c0.c"]), t(gd["c0.e"]), 0.0, None)
            d1 = vf.forward_tokens(*args)
            with torch.no_grad():
                w64 = om64.forward(bt, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]),
                                   {k: v.double() for k, v in prev.items()})
            for k in "xace":
                e0 = err(d0[k].cpu().numpy(), gd[f"c0.out.{k}"])
                e1 = err(d1[k].cpu().numpy(), gd[f"c1.out.{k}"])
                e64 = err(d1[k].cpu().numpy(), w64[k].numpy())
                r64 = err(gd[f"c1.out.{k}"], w64[k].numpy())
                am = bool(np.array_equal(d1[k].cpu().numpy().argmax(-1), gd[f"c1.out.{k}"].argmax(-1))) if k != "x" else None
                print(name, k, "first-step abs/rel", e0, "mid", e1, "cuda-vs-fp64", e64, "ref32-vs-fp64", r64, "argmax_eq", am)
                rows.append(dict(case=name, key=k, first=e0[0], mid=e1[0], cuda_vs_fp64=e64[0], ref_vs_fp64=r64[0]))
            del vf
        except Exception:
            traceback.print_exc()
    for name in ("itg_dev_T10", "itg_dev_T50", "itg_flowmol3_T10", "itg_flowmol3_T25"):
        try:
            gd = load_golden(name)
            cfg = ModelConfig.named(str(gd["config"]), int(gd["n_atom_types"]))
            sd = WT.init_state_dict(cfg, int(gd["weight_seed"]))
            vf = CTMCVectorFieldB200(cfg, sd)
            n_atoms = gd["n_atoms"]
            N, U = int(n_atoms.sum()), int((n_atoms * (n_atoms - 1) // 2).sum())
            A = cfg.n_atom_types
            t0 = time.time()
            out = vf.integrate_tokens(n_atoms, t(gd["x_0"]), torch.full((N,), A), torch.full((N,), 6), torch.full((U,), 4),
                                      int(gd["T"]), seed=int(gd["noise_seed"]))
            torch.cuda.synchronize()
            out = {k: v.cpu().numpy() for k, v in out.items()}
            print(name, "a mismatches", int((out["a"] != gd["a_1"]).sum()), "/", N, "c", int((out["c"] != gd["c_1"]).sum()),
                  "e", int((out["e"] != gd["e_1"]).sum()), "/", U, "x err", err(out["x"], gd["x_1"]), "launches",
                  vf.last_launches, "sec", round(time.time() - t0, 3))
            del vf
        except Exception:
            traceback.print_exc()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/diag.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
