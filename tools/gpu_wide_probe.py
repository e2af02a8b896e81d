"""Per-kernel timing of the wide tensor-core message pipeline under ncu, with parts disabled (tc_debug variants)."""
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
vf.set_option("conv_impl", 2)
vf.set_option("eg_nh", int(os.environ.get("EG_NH", "2")))
d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None); torch.cuda.synchronize()
for dbg in [int(a) for a in (sys.argv[1:] or ["0"])]:
    vf.set_option("tc_debug", dbg)
    print(f"dbg {dbg}: message pass ms {vf.time_conv_edge(1, 1):.3f}", flush=True)
