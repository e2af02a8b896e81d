#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -s > gpurun_out/r02y_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r02y_pytest_gpu.txt | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02y_smoke.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/r02y_bench.json
