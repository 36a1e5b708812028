#!/bin/bash
# bench (no profiler) -> launch list -> one ncu --set full capture of the step kernel; outputs under gpurun_out/<stem>*
S=${1:-r2c}
mkdir -p gpurun_out
bash tools/gpu_bench.sh > gpurun_out/${S}_bench.log 2>&1 || exit 1
cp gpurun_out/b20.json gpurun_out/${S}_bench.json; cp gpurun_out/b20_ref.json gpurun_out/${S}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${S}_launches.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/${S}_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 400 -c 3 -f -o gpurun_out/${S}_step \
  python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-extra-configs --no-graph > gpurun_out/${S}_ncu2.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/ | tail -12
