#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vec_b|k_node_update" -s 4 -c 2 -f -o gpurun_out/r01d_vecb python tools/gpu_quick_time.py 2:1:1 > gpurun_out/r01d_ncu.log 2>&1
tail -2 gpurun_out/r01d_ncu.log
