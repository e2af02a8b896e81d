#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_tc_probe.py 2>&1 | tail -15 | tee gpurun_out/tc_probe.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_edge_tc -s 2 -c 1 -f -o gpurun_out/tc_conv_edge python tools/gpu_tc_probe.py > gpurun_out/tc_ncu.log 2>&1
tail -3 gpurun_out/tc_ncu.log
