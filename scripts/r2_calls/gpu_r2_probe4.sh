#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/r2_tests4.log 2>&1
tail -n 3 gpurun_out/r2_tests4.log
python scripts/sweep_r2.py ldg 1 2 3 > gpurun_out/r2_sweep4.log 2>&1
python scripts/sweep_r2.py ip 1 2 3 >> gpurun_out/r2_sweep4.log 2>&1
python scripts/probe_small.py time > gpurun_out/r2_probe4_time.log 2>&1
for o in 2 3; do python scripts/probe_small.py cta ldg $o; done > gpurun_out/r2_probe4_cta.log 2>&1
python scripts/quick_bench.py --ldg 1 --n 2000 --reps 20 > gpurun_out/r2_probe4_ldg2000.log 2>&1
cat gpurun_out/r2_sweep4.log gpurun_out/r2_probe4_time.log gpurun_out/r2_probe4_cta.log gpurun_out/r2_probe4_ldg2000.log
