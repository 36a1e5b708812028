#!/bin/bash
# ncu --set full capture of 3 steady-state launches of the step kernel (cfg 2) -> gpurun_out/$1.ncu-rep
STEM=${1:-r2_step}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 400 -c 3 -f -o gpurun_out/$STEM \
  python bench.py --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/${STEM}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${STEM}_ncu.log; ls -la gpurun_out/$STEM.ncu-rep
