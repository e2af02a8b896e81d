"""In-situ per-kernel timing of warm forwards at GEOM-512 (flowmol3): CTMCVectorFieldB200.kernel_profile (fm_debug_kprof events).
    python tools/gpu_kprof.py [n_forwards]"""
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
for kv in filter(None, os.environ.get("FM_OPTS", "").split(",")):      # e.g. FM_OPTS=eg_pair=0,eu_fuse=0
    vf.set_option(kv.split("=")[0], int(kv.split("=")[1]))
n_atoms = draw_sizes("geom", 512)
x0, a0, c0, e0 = make_prior(n_atoms, 11, 100)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prof = vf.kernel_profile(n_atoms, x0, a0, c0, e0, n_forwards=nf)
tot = sum(t for _, t in prof.values())
print(f"{nf} forwards, {sum(c for c, _ in prof.values()):.0f} launches and {tot:.2f} ms per forward (event to event)")
for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:24s} n/fwd={c:5.1f} ms/fwd={t:7.3f} share={100 * t / tot:5.1f}% avg_us={1e3 * t / c:8.1f}")
