#!/bin/bash
# full GPU test suite, in-situ profile, bench lines (default workload + QM9-1024 + three sweep points)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -s > gpurun_out/r02c_pytest_gpu.txt 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error" gpurun_out/r02c_pytest_gpu.txt | tail -5
timeout 600 python tools/gpu_kprof.py 3 > gpurun_out/r02c_kprof.txt 2>&1; head -14 gpurun_out/r02c_kprof.txt
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02c_bench.json'))
    print('bench value', d['value'], 'e2e', d['e2e']['value'], 'api', d.get('e2e_api'), 'frac', d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r02c_bench.err').read()[-2000:])
PY
timeout 900 python bench.py --workload qm9_1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_qm9.json 2> gpurun_out/r02c_bench_qm9.err; cut -c1-300 gpurun_out/r02c_bench_qm9.json
