#!/bin/bash
# first GPU contact: sanitizer on a tiny case, diagnostics table, then the gpu test-suite (no -x: collect everything)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer.txt 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer.txt
timeout 900 python tools/gpu_diag.py > gpurun_out/diag.txt 2>&1
echo "diag exit $?" >> gpurun_out/diag.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/sanitizer.txt; tail -40 gpurun_out/diag.txt; tail -30 gpurun_out/pytest_gpu.txt
