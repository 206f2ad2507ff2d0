#!/bin/bash
# bign tests + device-resident and end-to-end rates of the bign paths given as arguments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bign.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for p in ${@:-bign_sign2 bign_verify}; do
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
