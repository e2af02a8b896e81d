#!/bin/bash
# round-end check on one GPU: smoke, full GPU test suite, default bench line: tools/gpu_final.sh <tag>
TAG=${1:-r01v}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches", "clocks")})
print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"])
print({k: round(v["frac_of_hbm_peak"], 3) for k, v in d["roofline"]["family"].items()})
PY
