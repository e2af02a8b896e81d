#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_quick_time.py 0 2:2 2:1 2>&1 | tail -4 | tee gpurun_out/wide_probe2.txt
EG_NH=1 timeout 600 python tools/gpu_diag_tc.py 2 2>&1 | grep -E "conv5|mid|mismatch|Traceback|Error" | tail -12 | tee -a gpurun_out/wide_probe2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_egemm|k_vec|k_conv_edge|k_edge_update|k_node_update" -c 40 --csv --log-file gpurun_out/wide_launches2.csv EG_NH=1 python tools/gpu_wide_probe.py 0 > gpurun_out/wide_ncu1.log 2>&1
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/wide_launches2.csv') if not l.startswith('==')]
for i,row in enumerate(csv.DictReader(lines)):
    k=row['Kernel Name']; m=re.search(r'\(int\)(\d)>', k)
    print(i, k.split('<')[0].replace('void ',''), m.group(1) if m and 'egemm' in k else '', row['Metric Value'])
PY
