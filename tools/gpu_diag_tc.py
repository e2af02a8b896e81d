"""Tensor-core message kernel vs goldens / fp32 kernel: error table + timing (no asserts)."""
import os, sys, time, traceback
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmol_b200 import graph as G, weights as WT
from flowmol_b200.config import ModelConfig
from oracle import flowmol_oracle as O
from tests.helpers import load_golden, t
from flowmol_b200.vector_field import CTMCVectorFieldB200


def err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


IMPLS = tuple(int(a) for a in sys.argv[1:]) or (0, 2)


def main():
    for name in ("fwd_flowmol3_taps", "fwd_flowmol3_geom"):
        gd = load_golden(name)
        cfg = ModelConfig.named("flowmol3", 11)
        sd = WT.init_state_dict(cfg, int(gd["weight_seed"]))
        vf = CTMCVectorFieldB200(cfg, sd)
        n_atoms = gd["n_atoms"]; N = int(n_atoms.sum())
        perm = torch.from_numpy(G.ref_edge_to_internal(n_atoms))
        prev = {k: t(gd[f"c0.out.{k}"]) for k in "xace"}
        args = (n_atoms, t(gd["c1.x_t"]), t(gd["c1.a"]), t(gd["c1.c"]), t(gd["c1.e"]), float(gd["c1.t"]), prev)
        for impl in IMPLS:
            try:
                vf.set_option("conv_impl", impl)
                vf.set_option("eg_nh", int(os.environ.get("EG_NH", "2")))
                if "c1.tap.conv0.s" in gd:
                    for l in range(cfg.n_convs):
                        vf.forward_tokens(*args, stop_after_conv=l); torch.cuda.synchronize()
                        s = vf.workspace_tensor("s").view(N, -1).cpu().numpy()
                        v = vf.workspace_tensor("v").view(N, 3, -1).permute(0, 2, 1).cpu().numpy()
                        print(name, "impl", impl, f"conv{l} s", err(s, gd[f"c1.tap.conv{l}.s"]), "v", err(v, gd[f"c1.tap.conv{l}.v"]))
                d0 = vf.forward_tokens(n_atoms, t(gd["c0.x_t"]), t(gd["c0.a"]), t(gd["c0.c"]), t(gd["c0.e"]), 0.0, None)
                d1 = vf.forward_tokens(*args); torch.cuda.synchronize()
                for k in "xace":
                    am = bool(np.array_equal(d1[k].cpu().numpy().argmax(-1), gd[f"c1.out.{k}"].argmax(-1))) if k != "x" else None
                    print(name, "impl", impl, k, "first", err(d0[k].cpu().numpy(), gd[f"c0.out.{k}"]), "mid", err(d1[k].cpu().numpy(), gd[f"c1.out.{k}"]), "argmax", am)
            except Exception:
                traceback.print_exc()
        del vf
    for name in ("itg_flowmol3_T10", "itg_flowmol3_T25"):
        gd = load_golden(name)
        cfg = ModelConfig.named("flowmol3", 11)
        vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, int(gd["weight_seed"])))
        n_atoms = gd["n_atoms"]; N, U = int(n_atoms.sum()), int((n_atoms * (n_atoms - 1) // 2).sum())
        for impl in IMPLS:
            try:
                vf.set_option("conv_impl", impl)
                vf.set_option("eg_nh", int(os.environ.get("EG_NH", "2")))
                out = vf.integrate_tokens(n_atoms, t(gd["x_0"]), torch.full((N,), 11), torch.full((N,), 6), torch.full((U,), 4), int(gd["T"]), seed=int(gd["noise_seed"]))
                torch.cuda.synchronize()
                out = {k: v.cpu().numpy() for k, v in out.items()}
                print(name, "impl", impl, "mismatch a/c/e", int((out["a"] != gd["a_1"]).sum()), int((out["c"] != gd["c_1"]).sum()), int((out["e"] != gd["e_1"]).sum()), "x err", err(out["x"], gd["x_1"]))
            except Exception:
                traceback.print_exc()
        del vf
    # timing at GEOM-512
    from bench import draw_sizes, make_prior
    cfg = ModelConfig.named("flowmol3", 11)
    vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0))
    n_atoms = draw_sizes("geom", 512)
    x0, a0, c0, e0 = make_prior(n_atoms, 11, 100)
    for impl in IMPLS:
        try:
            vf.set_option("conv_impl", impl)
            d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None); torch.cuda.synchronize()
            t0 = time.time(); d2 = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d); torch.cuda.synchronize(); dt = time.time() - t0
            print("GEOM-512 impl", impl, "forward ms", round(dt * 1e3, 2), "conv_edge ms", round(vf.time_conv_edge(1, 3), 3), "finite", bool(torch.isfinite(d2["x"]).all()))
            if impl == IMPLS[0]: ref = d2
            else: print("   tc vs fp32 forward:", {k: err(d2[k].cpu().numpy(), ref[k].cpu().numpy()) for k in "xace"})
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
