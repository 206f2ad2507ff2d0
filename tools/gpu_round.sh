#!/bin/bash
# final 1-GPU visit of a round: full GPU suite, bench + reference arm, ncu launch list + full captures, sanitizers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
oracle/_ref/reftests_b200 > gpurun_out/reftests.log 2>&1; echo "reftests rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print(json.dumps(d["also"]))
r=json.load(open("gpurun_out/bench_ref_n1.json"))
print(json.dumps(r["also"]))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
for k in bign_verify:bign_verify_kernel bign_sign2:bign_sign2_kernel belt_dwp:belt_dwp_mac_kernel belt_ecb:belt_ecb_kernel belt_ctr:belt_ctr_kernel bash512:bash_sponge_kernel; do
  p=${k%%:*}; r=${k#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$r -s 3 -c 1 -f -o gpurun_out/ncu_r02_$p python bench.py --paths $p --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$p.log 2>&1
  echo "ncu $p rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bign.py -x -q -k "ragged or G1_G2 or exceptional" --timeout 800 > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_bign.py tests/test_gpu_belt.py -x -q -k "G1_G2 or G6_G7 or dwp or DWP" --timeout 800 > gpurun_out/racecheck.log 2>&1; tail -3 gpurun_out/racecheck.log
