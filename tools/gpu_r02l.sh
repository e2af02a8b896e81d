#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_egemm_c" -s 2 -c 1 -f -o gpurun_out/r02l_egemm_c python tools/gpu_kprof.py 1 > gpurun_out/r02l_ncu.log 2>&1; tail -2 gpurun_out/r02l_ncu.log
