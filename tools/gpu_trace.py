"""Timeline of one k_egemm_tc CTA (clock64 stamps) per mode + per-forward timing."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmol_b200 import weights as WT, _lib
from flowmol_b200.config import ModelConfig
from flowmol_b200.vector_field import CTMCVectorFieldB200
from bench import draw_sizes, make_prior
cfg = ModelConfig.named("flowmol3", 11)
vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0))
n_atoms = draw_sizes("geom", 512)
x0, a0, c0, e0 = make_prior(n_atoms, 11, 100)
d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None); torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.time(); d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d); torch.cuda.synchronize(); ts.append(time.time() - t0)
print(f"forward ms {min(ts)*1e3:.2f}  message pass ms {vf.time_conv_edge(1, 3):.3f}", flush=True)
names = {0: "MSG0", 1: "MSG", 2: "GATE", 3: "EU1", 4: "EU2"}
for mode in (1, 0, 2, 4):
    for cta in (1000, 5000):
        vf.set_option("tc_trace_mode", mode); vf.set_option("tc_trace", cta)
        d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d); torch.cuda.synchronize()
        buf = np.zeros(64, np.int64)
        _lib.check(vf.lib.fm_debug_read_trace(vf._h, buf.ctypes.data))
        t0 = buf[0]
        rel = lambda i: int(buf[i] - t0) if buf[i] else None
        print(names[mode], "cta", cta, "first-fetch-issued", rel(1), "loader-arrive", [rel(2 + j) for j in range(10) if buf[2 + j]],
              "mma-issued", [rel(16 + j) for j in range(10) if buf[16 + j]], "epi-start", rel(30), "epi-end", rel(31), flush=True)
vf.set_option("tc_trace", -1)
