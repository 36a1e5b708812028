#!/bin/bash
# ncu --set full of 3 steady-state launches of the fused tcgen05 policy kernel (tools/policy_check.py 16384), the launch
# list of that workload and its un-profiled bench line -> gpurun_out/<stem>_*
STEM=${1:-r2_v7}
mkdir -p gpurun_out
timeout 200 python bench.py --workload rollout16384x8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${STEM}_bench_rollout16384x8.json 2> gpurun_out/${STEM}_bench_rollout.err
echo "bench rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 30 -c 3 -f -o gpurun_out/${STEM}_policy \
  python tools/policy_check.py 16384 > gpurun_out/${STEM}_policy_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${STEM}_policy.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${STEM}_launches_rollout.csv \
  python bench.py --workload rollout16384x8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${STEM}_launches_rollout.log 2>&1
echo "launch list rc=$?"
