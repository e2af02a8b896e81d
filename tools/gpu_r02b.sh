#!/bin/bash
# cluster-multicast probe + ncu capture of the register-resident vector stages
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu -s -k "tensor_core_message_kernels_agree" > gpurun_out/r02b_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.txt
timeout 600 python tools/gpu_cluster_probe.py > gpurun_out/r02b_cluster.txt 2>&1; cat gpurun_out/r02b_cluster.txt | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vecr -s 3 -c 3 -f -o gpurun_out/r02b_vecr python tools/gpu_kprof.py 1 > gpurun_out/r02b_ncu.log 2>&1; tail -3 gpurun_out/r02b_ncu.log
