"""Knock-out timing of the gate-fused message linears (k_egemm_g) at GEOM-512: option tc_debug switches off one role's work at a
time (results are garbage, the barrier protocol is intact) and the in-situ per-launch times show which role paces the kernel.
    python tools/gpu_knockout.py
bits: 1 no weight copies, 2 no main MMAs, 4 no activation-image copies, 8 no epilogue math / stores, 16 no gate MMAs,
      32 no global stores of images / gates, 64 no segment sums; k_egemm_c only: 128 epilogue 1 idle (barrier protocol only),
      256 epilogue 2 idle, 512 no rbf conversion in the loader warps"""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from flowmol_b200.vector_field import CTMCVectorFieldB200
from bench import draw_sizes, make_prior
warnings.simplefilter("ignore")
cfg = ModelConfig.named("flowmol3", 11)
vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0))
vf.auto_fallback = False
vf._status_fired = lambda: False          # garbage activations are expected here
n_atoms = draw_sizes("geom", 512)
x0, a0, c0, e0 = make_prior(n_atoms, 11, 100)
combos = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 64, 1 | 4, 2 | 16, 1 | 2 | 4 | 16, 8 | 2 | 16, 1 | 4 | 8, 127]
keys = ("EG_MSG0", "EG_MSG", "EG_MSGA", "k_egemm_c", "EG_GATE")
print("dbg   " + " ".join(f"{k:>9s}" for k in keys) + "   (avg us per launch)")
for dbg in combos:
    vf.set_option("tc_debug", dbg)
    try:
        prof = vf.kernel_profile(n_atoms, x0, a0, c0, e0, n_forwards=2)
    except RuntimeError as e:
        vf.get_option("status")
        print(f"{dbg:4d}  status fired: {str(e)[:60]}")
        continue
    print(f"{dbg:4d}  " + " ".join(f"{1e3 * prof[k][1] / prof[k][0]:9.1f}" if k in prof else "        -" for k in keys), flush=True)
vf.set_option("tc_debug", 0)
