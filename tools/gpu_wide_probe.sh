#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_wide_probe.py 0 1 2 4 8 6 14 2>&1 | tail -9 | tee gpurun_out/wide_probe.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_egemm|k_vec" -c 60 --csv --log-file gpurun_out/wide_launches.csv python tools/gpu_wide_probe.py 0 2 4 8 > gpurun_out/wide_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_egemm_tc -s 13 -c 2 -f -o gpurun_out/wide_egemm python tools/gpu_wide_probe.py 0 > gpurun_out/wide_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vec_b -s 2 -c 1 -f -o gpurun_out/wide_vecb python tools/gpu_wide_probe.py 0 > gpurun_out/wide_ncu3.log 2>&1
tail -2 gpurun_out/wide_ncu2.log gpurun_out/wide_ncu3.log
