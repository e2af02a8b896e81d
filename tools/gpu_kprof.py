"""In-situ per-kernel timing of one warm forward at GEOM-512 (flowmol3): fm_debug_kprof events, aggregated by launch site.
    python tools/gpu_kprof.py [n_forwards]        (labels are read from csrc/api.cu by line number)"""
import collections, ctypes as C, os, re, sys
import numpy as np, torch
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
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 3
d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.0, None)
for _ in range(2):
    d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d)
torch.cuda.synchronize()
vf.set_option("kprof", 1)
for _ in range(nf):
    d = vf.forward_tokens(n_atoms, x0, a0, c0, e0, 0.3, d)
torch.cuda.synchronize()
cap = 4096
lines, ms, n = (C.c_int32 * cap)(), (C.c_float * cap)(), C.c_int32()
from flowmol_b200 import _lib
_lib.check(vf.lib.fm_debug_kprof(vf._h, lines, ms, cap, C.byref(n)))
vf.set_option("kprof", 0)
src = open(os.path.join(ROOT, "flowmol_b200", "csrc", "api.cu")).read().split("\n")
def label(line):
    for back in range(0, 4):                       # the launch is on the LAUNCH_OK line or just above it
        t = src[line - 1 - back]
        m = re.search(r"launch_eg<D, fm::(EG_\w+), \d(?:, ([^>]+))?>", t) or re.search(r"fm::(k_\w+)<", t) or re.search(r"\b(scalar|gate|linear)\(", t)
        if m:
            return m.group(1) + (" img" if m.lastindex and m.lastindex >= 2 and m.group(2) else "")
    return f"line {line}"
agg = collections.OrderedDict()
seen = collections.Counter()
for i in range(n.value):
    lab = label(lines[i])
    if lab.startswith("EG_MSG"):            # one launch site, three linears of a message pass in turn: MSG0, MSG, MSGA
        lab = ("EG_MSG0", "EG_MSG", "EG_MSGA")[seen[lines[i]] % 3]
        seen[lines[i]] += 1
    k = f"{lab:22s} @{lines[i]}"
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms[i]
tot = sum(v[1] for v in agg.values())
print(f"{nf} forwards, {n.value} launches, {tot / nf:.2f} ms per forward (event to event)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} n/fwd={c / nf:5.1f} ms/fwd={t / nf:7.3f} share={100 * t / tot:5.1f}% avg_us={1e3 * t / c:8.1f}")
