#!/bin/bash
# one GPU visit: the whole -m gpu suite, then the bench paths given as arguments (default: the two bign paths)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
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
