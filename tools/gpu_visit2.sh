#!/bin/bash
# bign tests + verify bench + one full ncu capture of the verification kernel (DRAM bytes, stalls)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bign.py -k "G1_G2 or ragged or random" -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for p in bign_verify; do
  timeout 600 python bench.py --paths $p --no-cpu-baseline --steps 10 > gpurun_out/visit_$p.json 2> gpurun_out/visit_$p.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/visit_$p.json'))
    print('$p', d['value'], d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'))
except Exception as e:
    print('$p failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bign_verify_kernel -s 3 -c 1 -f -o /tmp/ncu_v python bench.py --paths bign_verify --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_bign_verify.log 2>&1
ncu -i /tmp/ncu_v.ncu-rep --page raw --csv > gpurun_out/ncu_raw_bign_verify_wg.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/ncu_raw_bign_verify_wg.csv')))
h, u, v = rows[0], rows[1], rows[2]
for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'sm__instruction_throughput.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'smsp__inst_executed.sum'):
    if k in h:
        i = h.index(k); print(k, v[i], u[i])
PY
