#!/bin/bash
# per-kernel durations of forwards at GEOM-512 (cold-cache, serialised: compare SHARES): tools/gpu_launchlist.sh <tag> [spec]
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/gpu_quick_time.py ${2:-2:1:1:1:1:1} > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary.txt
