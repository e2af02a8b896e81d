"""Quick per-forward timing at GEOM-512 (flowmol3) for the current build."""
import os, sys, time
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
for spec in (sys.argv[1:] or ["0"]):
    impl, nh, gnh, node, prec, pers, img = (spec.split(":") + ["1", "1", "1", "1", "1", "1"])[:7]
    impl, nh, gnh, node, prec, pers, img = int(impl), int(nh), int(gnh), int(node), int(prec), int(pers), int(img)
    if impl == 2:
        vf.set_option("tc_prec", prec)
        vf.set_option("eg_persist", pers)
        vf.set_option("eg_img", img)
    vf.set_option("conv_impl", impl)
    vf.set_option("eg_nh", nh)
    vf.set_option("eg_nh_gate", gnh)
    vf.set_option("node_impl", node if impl == 2 else 0)
    d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.time(); d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d); torch.cuda.synchronize(); ts.append(time.time() - t0)
    print(f"impl {impl} eg_nh {nh} gate_nh {gnh} node_impl {node} tc_prec {prec} persist {pers} img {img}: egemm_msg ms {vf.time_egemm_msg(1, 5) if impl == 2 else 0:.3f} forward ms {min(ts)*1e3:.2f}  conv_edge ms {vf.time_conv_edge(1, 3):.3f}")
