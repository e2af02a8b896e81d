#!/bin/bash
# full GPU tests, in-situ profile, sweep bench lines (configs[4]: n = 10, 40, 80 fixed and mixed 40), ncu launch list + full capture, SASS
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -s > gpurun_out/r02e_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/r02e_pytest_gpu.txt | tail -3
timeout 600 python tools/gpu_kprof.py 3 > gpurun_out/r02e_kprof.txt 2>&1; head -16 gpurun_out/r02e_kprof.txt
for wl in sweep_n10 sweep_n40 sweep_mix40 sweep_n80; do
  timeout 1200 python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-api-e2e > gpurun_out/r02e_bench_$wl.json 2> gpurun_out/r02e_bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02e_bench_$wl.json')); print('$wl', round(d['value'],1), 'mol/s  e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step']), 'batch', d['batch'])
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/r02e_bench_$wl.err').read()[-1500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 1 --warmup 0 --timesteps 3 --no-cpu-baseline --no-api-e2e > gpurun_out/r02e_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_egemm_p|k_vecr" -s 30 -c 10 -f -o gpurun_out/r02e_full python bench.py --steps 1 --warmup 0 --timesteps 3 --no-cpu-baseline --no-api-e2e > gpurun_out/r02e_ncu_full.log 2>&1; tail -2 gpurun_out/r02e_ncu_full.log
