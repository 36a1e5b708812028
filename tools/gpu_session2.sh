#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.txt
tail -5 gpurun_out/s2_pytest.txt
timeout 600 python tools/layout_sweep.py 4096 8192 16384 65536 > gpurun_out/s2_sweep.txt 2>&1
cat gpurun_out/s2_sweep.txt
