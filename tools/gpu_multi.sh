#!/bin/bash
# N-GPU visit: multi-GPU parity tests, then bench.py under torchrun at N = $1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_multi_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("value", d["value"], "e2e", d["e2e"])
print(json.dumps(d["config"].get("strong"), indent=1))
print(json.dumps(d["also"]))
PY
tail -5 gpurun_out/bench_n$N.err
