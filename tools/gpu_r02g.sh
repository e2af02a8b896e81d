#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu -k "agree or golden_at_probe or large_molecules" > gpurun_out/r02g_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02g_pytest.txt
timeout 600 python tools/gpu_kprof.py 3 > gpurun_out/r02g_kprof.txt 2>&1; head -14 gpurun_out/r02g_kprof.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_egemm_g" -s 2 -c 2 -f -o gpurun_out/r02g_egemm_g python tools/gpu_kprof.py 1 > gpurun_out/r02g_ncu.log 2>&1; tail -2 gpurun_out/r02g_ncu.log
