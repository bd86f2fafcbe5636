#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py tests/test_full_size_gpu.py -q -x -m gpu > gpurun_out/r2_tests6.log 2>&1
tail -n 4 gpurun_out/r2_tests6.log
python scripts/sweep_r2.py ldg 1 2 3 > gpurun_out/r2_sweep6.log 2>&1
python scripts/sweep_r2.py ip 1 2 3 >> gpurun_out/r2_sweep6.log 2>&1
python scripts/probe_small.py time > gpurun_out/r2_probe6_time.log 2>&1
for o in 3; do python scripts/probe_small.py cta ldg $o; done > gpurun_out/r2_probe6_cta.log 2>&1
python scripts/probe_small.py cta ip 1 8 >> gpurun_out/r2_probe6_cta.log 2>&1
cat gpurun_out/r2_sweep6.log gpurun_out/r2_probe6_time.log gpurun_out/r2_probe6_cta.log
