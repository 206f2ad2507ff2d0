#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bign.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
for mode in staged nostaging nozerocopy; do
  case $mode in staged) env="";; nostaging) env="B2G_NO_STAGING=1";; nozerocopy) env="B2G_NO_ZEROCOPY=1";; esac
  env $env timeout 300 python bench.py --paths bign_verify --no-cpu-baseline --steps 10 > gpurun_out/ab_$mode.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_$mode.json')); print('$mode', 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2))"
done
