#!/bin/bash
# full ncu capture of selected kernels of one forward: tools/gpu_ncu_vec.sh <regex> <tag> [skip] [count]
RE=${1:-"k_vec_a|k_vec_c"}; TAG=${2:-r01e_vec}; SKIP=${3:-4}; CNT=${4:-2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/$TAG python tools/gpu_quick_time.py 2:1:1:1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
