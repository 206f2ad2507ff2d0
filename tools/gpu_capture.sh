#!/bin/bash
# ncu launch list + full captures of the dominant kernels; reports are converted to CSV on the box and
# deleted (gpurun copies back at most 64 MiB)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
oracle/_ref/reftests_b200 > gpurun_out/reftests.log 2>&1; echo "reftests rc=$?"
B2G_CPU_BELOW=1048576 oracle/_ref/reftests_b200 bash belt bign128 > gpurun_out/reftests_small.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
for k in bign_verify:bign_verify_kernel bign_sign2:bign_sign2_kernel belt_dwp:belt_dwp_mac_kernel belt_ecb:belt_ecb_kernel belt_ctr:belt_ctr_kernel bash512:bash_sponge_kernel; do
  p=${k%%:*}; r=${k#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$r -s 3 -c 1 -f -o /tmp/ncu_$p python bench.py --paths $p --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$p.log 2>&1
  echo "ncu $p rc=$?"
  ncu -i /tmp/ncu_$p.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_$p.csv 2>/dev/null
  case $p in bign_*) ncu -i /tmp/ncu_$p.ncu-rep --page source --csv > gpurun_out/r02_ncu_source_$p.csv 2>/dev/null;; esac
done
ls -la gpurun_out | head -40; du -sh gpurun_out
