#!/bin/bash
# full ncu capture of the first message pass's vector-stage kernels: tools/gpu_ncu_vec2.sh <tag>
TAG=${1:-r01o}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"k_vec_a|k_vec_b|k_vec_c" -c 4 -f -o gpurun_out/${TAG}_vec \
    python tools/gpu_quick_time.py 2:1:1:1:1:1:1 > gpurun_out/${TAG}_vec_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_vec_ncu.log
