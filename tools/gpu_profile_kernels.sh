#!/bin/bash
# launch list of a few forwards + full ncu captures of the first message pass's kernels: tools/gpu_profile_kernels.sh <tag>
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/gpu_quick_time.py 2:1:1:1:1:1:1 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launch_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:"k_vec_a|k_vec_b|k_vec_c" -c 4 -f -o gpurun_out/${TAG}_vec \
    python tools/gpu_quick_time.py 2:1:1:1:1:1:1 > gpurun_out/${TAG}_vec_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_egemm_p" -c 5 -f -o gpurun_out/${TAG}_egemm \
    python tools/gpu_quick_time.py 2:1:1:1:1:1:1 > gpurun_out/${TAG}_egemm_ncu.log 2>&1
ls -la gpurun_out/${TAG}*
