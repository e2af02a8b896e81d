#!/bin/bash
# quick validation of the fp16x3 operand path: building blocks, kernel agreement, goldens, then timing of both formats
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "building_block or agree or overflow or golden" 2>&1 | tail -15 | tee gpurun_out/r01e_probe_pytest.txt
timeout 300 python tools/gpu_quick_time.py 2:1:1:1:1 2:1:1:1:0 2>&1 | tail -4 | tee gpurun_out/r01e_probe_time.txt
