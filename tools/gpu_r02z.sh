#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r02z_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02z_pytest_gpu.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02z_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_evaluation']['ms_event_to_event'])"
