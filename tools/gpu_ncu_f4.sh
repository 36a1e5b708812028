#!/bin/bash
# ncu --set full captures of 3 steady-state launches of the self-collision and terrain instantiations of the step kernel
# (bench sub-records f4_selfcol4096 / f4_terrain4096) -> gpurun_out/<stem>_{selfcol,terrain}.ncu-rep, plus their un-profiled bench lines
STEM=${1:-r2_v7}
mkdir -p gpurun_out
for W in selfcol4096 terrain4096; do
  timeout 200 python bench.py --workload $W --steps 500 --warmup 50 --no-cpu-baseline > gpurun_out/${STEM}_bench_$W.json 2> gpurun_out/${STEM}_bench_$W.err
  echo "bench $W rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 400 -c 3 -f -o gpurun_out/${STEM}_$W \
    python bench.py --workload $W --steps 500 --warmup 5 --no-cpu-baseline > gpurun_out/${STEM}_${W}_ncu.log 2>&1
  echo "ncu $W rc=$?"; ls -la gpurun_out/${STEM}_$W.ncu-rep
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${STEM}_launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${STEM}_launches.log 2>&1
echo "launch list rc=$?"
