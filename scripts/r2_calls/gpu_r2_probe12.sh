#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -x -m gpu > gpurun_out/r2_tests12.log 2>&1
tail -n 5 gpurun_out/r2_tests12.log | cut -c1-200
python scripts/sweep_r2.py ip 1 2 3 > gpurun_out/r2_sweep12.log 2>&1
python scripts/sweep_r2.py ldg 1 2 3 >> gpurun_out/r2_sweep12.log 2>&1
python scripts/probe_small.py time > gpurun_out/r2_probe12_time.log 2>&1
python scripts/probe_small.py cta ip 1 8 > gpurun_out/r2_probe12_cta.log 2>&1
cat gpurun_out/r2_sweep12.log gpurun_out/r2_probe12_time.log gpurun_out/r2_probe12_cta.log
