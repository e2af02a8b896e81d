"""Cluster-multicast weight streaming of k_egemm_p at GEOM-512: per-kernel in-situ times for eg_cluster = 1, 2, 4 and a bitwise
comparison of the network outputs.   python tools/gpu_cluster_probe.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from flowmol_b200.vector_field import CTMCVectorFieldB200
from bench import draw_sizes, make_prior
cfg = ModelConfig.named("flowmol3", 11)
vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0))
n_atoms = draw_sizes("geom", 512)
x0, a0, c0, e0 = make_prior(n_atoms, 11, 100)
ref = None
vf.set_option("eg_fuse_gate", 0)          # cluster launches exist for the un-fused chain only
for cl in (1, 2, 4, 1):
    vf.set_option("eg_cluster", cl)
    d0 = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None)
    d1 = {k: v.clone() for k, v in vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d0).items()}
    if ref is None:
        ref = d1
    same = all(torch.equal(ref[k], d1[k]) for k in "xace")
    prof = vf.kernel_profile(n_atoms, x0, a0, c0, e0, n_forwards=3)
    tot = sum(t for _, t in prof.values())
    eg = "  ".join(f"{k} {1e3 * t / c:6.1f}" for k, (c, t) in sorted(prof.items()) if k.startswith("EG_"))
    print(f"eg_cluster {cl} (max active clusters {vf.get_option('eg_clusters_seen')}): forward {tot:6.2f} ms  bitwise == cluster 1: {same}  us/launch: {eg}", flush=True)
