#!/bin/bash
# One GPU call: parity tests, bench line, ncu launch list (3-timestep pass), one full capture of the top kernels.
# Usage: tools/gpu_round.sh <tag> [ncu-kernel-regex]
TAG=${1:-r01}
RE=${2:-k_egemm_p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 900 python bench.py --steps 1 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --timesteps 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"$RE" -s 12 -c 6 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 1 --warmup 0 --timesteps 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -8
