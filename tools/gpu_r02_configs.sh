#!/bin/bash
# BASELINE.json configs[1] (QM9-1024), configs[0] on the GPU (dev, 32 QM9-sized, T=50) and configs[4] size-sweep points on one B200
mkdir -p gpurun_out
for w in qm9_1024 dev_qm9_32 sweep_n10 sweep_n40 sweep_n80 sweep_mix40; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$w.json')); print('$w', round(d['value'],1), 'mol/s  ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'api', round(d['e2e_api']['value'],1))"
done
