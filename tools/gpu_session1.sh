#!/bin/bash
# first GPU session of round 2: tests, layout sweep, contract bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.txt
tail -5 gpurun_out/s1_pytest.txt
timeout 600 python tools/layout_sweep.py 4096 8192 16384 32768 65536 > gpurun_out/s1_sweep.txt 2>&1
cat gpurun_out/s1_sweep.txt
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; tail -c 1500 gpurun_out/s1_bench.json
