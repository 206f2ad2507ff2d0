#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bash.py -x -q 2>&1 | tail -2
for v in main "$@"; do
  if [ $v = main ]; then unset BEE2_B200_LIB; else export BEE2_B200_LIB=$PWD/gpurun_scratch/$v.so; fi
  timeout 300 python bench.py --paths bash512 --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bash_$v.json 2> gpurun_out/bash_$v.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/bash_$v.json')); print('$v', round(d['value'],1), round(d['ms_per_step'],3), d['checksum'] if 'checksum' in d else d.get('paths'))
except Exception as e: print('$v failed', e)"
done
