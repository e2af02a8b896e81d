#!/bin/bash
# round-2 evidence: ncu launch list of the bench command + full captures of the hot kernels (one GPU, never under a timed run).
# gpurun brings back at most 64 MiB: the raw pages are exported to CSV on the box, only the tcgen05-linear report itself travels.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_ncu_launch_list.csv \
  python bench.py --steps 1 --warmup 1 --timesteps 3 --no-cpu-baseline --no-api-e2e > gpurun_out/r02_launchlist_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_egemm_g|k_egemm_c|k_egemm_p" -c 5 -f -o gpurun_out/r02_full_egemm \
  python tools/gpu_kprof.py 1 > gpurun_out/r02_full_egemm.log 2>&1; echo "full egemm rc=$?"
ncu -i gpurun_out/r02_full_egemm.ncu-rep --page raw --csv > gpurun_out/r02_full_egemm_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"k_vecr_|k_edge_init_r|k_edge_head_r|k_node_pre|k_node_mid|k_node_embed" -c 8 -f -o /tmp/r02_full_other \
  python tools/gpu_kprof.py 1 > gpurun_out/r02_full_other.log 2>&1; echo "full other rc=$?"
ncu -i /tmp/r02_full_other.ncu-rep --page raw --csv > gpurun_out/r02_full_other_raw.csv 2>/dev/null
du -sh gpurun_out
