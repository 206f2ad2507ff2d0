#!/bin/bash
# quick GPU visit: bign parity tests + bign bench (optionally for variant libs given as args)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bign.py -x -q > gpurun_out/pytest_bign.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_bign.log
timeout 300 python bench.py --paths bign_verify --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bign_main.json 2> gpurun_out/bign_main.err
python -c "
import json
d=json.load(open('gpurun_out/bign_main.json')); print('main', d['value'], d['ms_per_step'])"
for v in "$@"; do
  BEE2_B200_LIB=$PWD/gpurun_scratch/$v.so timeout 300 python bench.py --paths bign_verify --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bign_$v.json 2> gpurun_out/bign_$v.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/bign_$v.json')); print('$v', d['value'], d['ms_per_step'])
except Exception as e: print('$v failed', e)"
done
