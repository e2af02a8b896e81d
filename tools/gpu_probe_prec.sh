#!/bin/bash
# quick validation of the tensor-core linears: building blocks, kernel agreement, goldens, then timing of the variants
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "building_block or agree or overflow or golden or edge_cases or large_mol" 2>&1 | tail -15 | tee gpurun_out/${TAG}_probe_pytest.txt
timeout 300 python tools/gpu_quick_time.py 2:1:1:1:1:1:1 2:1:1:1:1:1:0 2>&1 | tail -4 | tee gpurun_out/${TAG}_probe_time.txt
