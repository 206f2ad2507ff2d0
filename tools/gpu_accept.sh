#!/bin/bash
mkdir -p gpurun_out
oracle/_ref/reftests_b200 > gpurun_out/reftests.log 2>&1; echo "reftests rc=$?"; cat gpurun_out/reftests.log | tail -15
B2G_CPU_BELOW=1048576 oracle/_ref/reftests_b200 bash belt bign128 > gpurun_out/reftests_small.log 2>&1; echo "small rc=$?"; tail -6 gpurun_out/reftests_small.log
timeout 900 python -m pytest tests/test_gpu_reftests.py tests/test_gpu_bign.py -x -q 2>&1 | tail -5
tools/gpu_variants.sh
PATHS=belt_ecb tools/gpu_variants.sh | tail -1
