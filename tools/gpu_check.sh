#!/bin/bash
# One GPU call: parity tests (with the measured maxima, -s) and the in-situ per-kernel profile at GEOM-512.
# Usage: tools/gpu_check.sh <tag> [pytest -k expression]
TAG=${1:-r02}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -q -x -m gpu -s -k "$KEXPR" > gpurun_out/${TAG}_pytest_gpu.txt 2>&1
else
  timeout 1500 python -m pytest tests -q -x -m gpu -s > gpurun_out/${TAG}_pytest_gpu.txt 2>&1
fi
echo "pytest rc=$?"
grep -E "^\[parity\]|passed|failed|Error|error" gpurun_out/${TAG}_pytest_gpu.txt | tail -60
timeout 600 python tools/gpu_kprof.py 3 > gpurun_out/${TAG}_kprof.txt 2>&1
cat gpurun_out/${TAG}_kprof.txt | head -40
