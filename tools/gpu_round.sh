#!/bin/bash
# One GPU-box visit: parity tests, the default bench line (+ reference arm), then optional ncu captures.
#   tools/gpu_round.sh [ncu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
  for k in bign_verify:bign_verify bign_sign2:bign_sign2; do
    p=${k%%:*}; r=${k#*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:${r}_kernel -s 3 -c 1 -f -o gpurun_out/ncu_$p python bench.py --paths $p --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$p.log 2>&1
  done
fi
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print("belt_ctr", d["value"], d["e2e"]["value"])
for k,v in d["paths"].items(): print(k, v["value"], (v.get("issue_roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"))
PY
