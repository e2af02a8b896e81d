#!/bin/bash
# final captures of the round: the quad-layout EdgeUpdate kernel and the twelve-epilogue-warp message linears
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"k_egemm_c|k_egemm_h" -c 3 -f -o /tmp/r02_full_final python tools/gpu_kprof.py 1 > gpurun_out/r02_full_final.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_full_final.ncu-rep --page raw --csv > gpurun_out/r02_full_final_raw.csv 2>/dev/null
timeout 200 python tools/gpu_kprof.py 3 > gpurun_out/r02_kprof_last.txt 2>&1; cat gpurun_out/r02_kprof_last.txt
