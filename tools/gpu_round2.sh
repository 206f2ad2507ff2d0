#!/bin/bash
# One 1-GPU visit: full GPU parity suite, default bench line + reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-10} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print(json.dumps(d["also"]))
print("e2e", d["e2e"])
r=json.load(open("gpurun_out/bench_ref_n1.json"))
print(json.dumps(r["also"]))
PY
tail -5 gpurun_out/bench_n1.err
