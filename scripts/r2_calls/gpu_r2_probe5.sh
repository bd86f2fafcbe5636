#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py tests/test_full_size_gpu.py tests/test_fd_gpu.py -q -m gpu > gpurun_out/r2_tests5.log 2>&1
tail -n 12 gpurun_out/r2_tests5.log
python scripts/probe_small.py time > gpurun_out/r2_probe5_time.log 2>&1
for o in 1 2 3; do python scripts/probe_small.py cta ldg $o; done > gpurun_out/r2_probe5_cta.log 2>&1
python scripts/probe_small.py cta ip 1 8 >> gpurun_out/r2_probe5_cta.log 2>&1
python scripts/probe_small.py cta ip 3 8 >> gpurun_out/r2_probe5_cta.log 2>&1
cat gpurun_out/r2_probe5_time.log gpurun_out/r2_probe5_cta.log
