#!/bin/bash
# bench + ncu evidence in one GPU call.  Usage: tools/gpu_bench_profile.sh <tag> [full]
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
if [ "$2" == "full" ]; then
  timeout 2400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench_reference.json
fi
# launch list (cold-cache, serialised: compare SHARES) of a 3-timestep pass
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --timesteps 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
# full capture of the dominant kernel (the 292->256 message linear) and of the vector stage
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_egemm_tc -s 40 -c 3 -f -o gpurun_out/${TAG}_egemm \
    python bench.py --steps 1 --warmup 0 --timesteps 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
