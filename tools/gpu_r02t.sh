#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_edge_init_r|k_edge_head_r" -s 2 -c 2 -f -o gpurun_out/r02t_edge_reg python tools/gpu_kprof.py 1 > gpurun_out/r02t_ncu.log 2>&1; tail -2 gpurun_out/r02t_ncu.log
