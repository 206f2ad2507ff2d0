#!/bin/bash
# 1-GPU visit: full GPU suite, bench, verify-chunk sweep, then ncu launch list + full captures of the dominant kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print(json.dumps(d["also"]))
PY
for c in 32768 65536 131072 262144; do
  B2G_BIGN_CHUNK=$c timeout 300 python bench.py --paths bign_verify --no-cpu-baseline --steps 5 > gpurun_out/chunk_$c.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/chunk_$c.json')); print('chunk $c e2e', round(d['e2e']['value']/1e6,2), 'M/s  value', round(d['value']/1e6,2))"
done
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
  for k in bign_verify:bign_verify_kernel bign_sign2:bign_sign2_kernel belt_dwp:belt_dwp_mac_kernel belt_ecb:belt_ecb_kernel belt_ctr:belt_ctr_kernel bash512:bash_sponge_kernel; do
    p=${k%%:*}; r=${k#*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$r -s 3 -c 1 -f -o gpurun_out/ncu_r02_$p python bench.py --paths $p --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$p.log 2>&1
    echo "ncu $p rc=$?"
  done
fi
