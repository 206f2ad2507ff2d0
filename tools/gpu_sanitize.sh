#!/bin/bash
# compute-sanitizer over the round-2 bign kernels (global window table, staged signing) + small-batch rates
mkdir -p gpurun_out
python tools/gpu_small_verify.py 8192 16384 32768 65536 131072 262144 > gpurun_out/small_verify.log 2>&1; cat gpurun_out/small_verify.log
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bign.py -x -q -k "ragged or G1_G2 or G6_G7 or exceptional" > gpurun_out/memcheck_r02b.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02b.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_bign.py -x -q -k "sign2_ragged or G6_G7 or G1_G2" > gpurun_out/racecheck_r02b.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck_r02b.log | tail -3
