#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu -s -k "agree or golden_at_probe or large_molecules" > gpurun_out/r02h_pytest.txt 2>&1; echo "pytest rc=$?"; grep -E "parity\] gate|passed|failed|Error" gpurun_out/r02h_pytest.txt | tail -4
timeout 600 python tools/gpu_kprof.py 3 > gpurun_out/r02h_kprof.txt 2>&1; head -12 gpurun_out/r02h_kprof.txt
