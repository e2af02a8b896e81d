#!/bin/bash
# gate-fused message linears (k_egemm_g): agreement test first (bounded), then profile, then the whole suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu -s -k "agree" > gpurun_out/r02f_pytest_agree.txt 2>&1; echo "agree rc=$?"; grep -E "parity\] (gate|edges|impl 2)|passed|failed|Error|error" gpurun_out/r02f_pytest_agree.txt | tail
timeout 600 python tools/gpu_kprof.py 3 > gpurun_out/r02f_kprof.txt 2>&1; head -16 gpurun_out/r02f_kprof.txt
timeout 1500 python -m pytest tests -q -x -m gpu -s > gpurun_out/r02f_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/r02f_pytest_gpu.txt | tail -3
