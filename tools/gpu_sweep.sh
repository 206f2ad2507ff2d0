#!/bin/bash
# verification rate against batch size (wave quantisation) and CTA size
python tools/gpu_small_verify.py 227328 262144 340992 454656 524288 1048576
for t in 224 192 160 128; do echo "threads=$t"; B2G_BIGN_THREADS=$t python tools/gpu_small_verify.py 262144 340992; done
