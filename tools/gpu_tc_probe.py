"""Where does k_conv_edge_tc spend its time?  Timing with parts disabled (results are garbage in those variants)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmol_b200 import weights as WT
from flowmol_b200.config import ModelConfig
from flowmol_b200.vector_field import CTMCVectorFieldB200
from bench import draw_sizes, make_prior
cfg = ModelConfig.named("flowmol3", 11)
vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0))
n_atoms = draw_sizes("geom", 512)
x0, a0, c0, e0 = make_prior(n_atoms, 11, 100)
vf.set_option("conv_impl", 0)
d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None); torch.cuda.synchronize()
print("fp32 kernel ms", vf.time_conv_edge(1, 3))
vf.set_option("conv_impl", 1)
for dbg, what in ((0, "full"), (1, "no weight copies"), (2, "no MMA issue"), (3, "no copies, no MMA"), (4, "no vector stages"), (8, "no x_store"),
                  (12, "no vec, no x_store"), (15, "skeleton: barriers + TMEM loads + gather only")):
    vf.set_option("tc_debug", dbg)
    print(f"tc dbg={dbg:2d} {what:45s} ms {vf.time_conv_edge(1, 3):8.3f}")
vf.set_option("tc_debug", 0)
