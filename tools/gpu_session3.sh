#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.txt
tail -4 gpurun_out/s3_pytest.txt
timeout 600 python tools/layout_sweep.py 4096 16384 65536 > gpurun_out/s3_sweep.txt 2>&1
cat gpurun_out/s3_sweep.txt
(timeout 300 python tools/warp_profile.py 4096; timeout 300 python tools/warp_profile.py 16384) > gpurun_out/r2_warp_phases.txt 2>&1
cat gpurun_out/r2_warp_phases.txt
