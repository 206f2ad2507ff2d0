#!/bin/bash
# One GPU-box visit: parity tests, the default bench line, then bign launch-config variants.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
for v in "$@"; do
  BEE2_B200_LIB=$PWD/gpurun_scratch/$v.so timeout 300 python bench.py --paths bign_verify --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bign_$v.json 2> gpurun_out/bign_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bign_$v.json")); print("$v", d["value"], d["ms_per_step"])
except Exception as e: print("$v failed", e)
PY
done
