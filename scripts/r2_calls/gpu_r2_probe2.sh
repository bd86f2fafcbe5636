#!/bin/bash
# round 2, probe 2 (one GPU): parity of the reworked march kernels (tile plan, tail queue, PDL, fused step), then timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py tests/test_full_size_gpu.py tests/test_fd_gpu.py -x -q -m gpu > gpurun_out/r2_tests2.log 2>&1
tail -n 5 gpurun_out/r2_tests2.log
python scripts/probe_small.py time > gpurun_out/r2_probe2_time.log 2>&1
STEFAN_PDL=0 python scripts/probe_small.py time > gpurun_out/r2_probe2_time_nopdl.log 2>&1
python scripts/probe_small.py step > gpurun_out/r2_probe2_step.log 2>&1
for o in 1 2 3; do python scripts/probe_small.py cta ldg $o; done > gpurun_out/r2_probe2_cta.log 2>&1
python scripts/probe_small.py cta ip 1 8 >> gpurun_out/r2_probe2_cta.log 2>&1
cat gpurun_out/r2_probe2_time.log gpurun_out/r2_probe2_step.log gpurun_out/r2_probe2_cta.log
