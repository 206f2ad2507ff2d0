#!/bin/bash
# one full ncu capture of a kernel: tools/gpu_ncu1.sh PATH KERNEL_REGEX TAG  -> gpurun_out/ncu_{raw,source}_TAG.csv
p=$1; r=$2; tag=$3
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$r -s 3 -c 1 -f -o /tmp/ncu_$tag python bench.py --paths $p --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_source_$tag.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/ncu_raw_$tag.csv')))
h, u, v = rows[0], rows[1], rows[2]
for k in h:
    if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') or k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__instruction_throughput.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active'):
        i = h.index(k); print(k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), v[i], u[i])
PY
