#!/bin/bash
# bign_verify device-resident rate for the main library and variant builds (gpurun_scratch/NAME.so given as args)
mkdir -p gpurun_out
run() {
  name=$1; lib=$2
  BEE2_B200_LIB=$lib timeout 300 python bench.py --paths ${PATHS:-bign_verify} --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/var_$name.json')); print('$name', round(d['value']/1e6,3), round(d['ms_per_step'],4))
except Exception as e: print('$name failed', e)"
}
run main $PWD/bee2_b200/libbee2_b200.so
for v in "$@"; do run $v $PWD/gpurun_scratch/$v.so; done
