#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/sweep_r2.py ldg 1 2 3 > gpurun_out/r2_sweep_ldg.log 2>&1
python scripts/sweep_r2.py ip 1 2 3 > gpurun_out/r2_sweep_ip.log 2>&1
STEFAN_DEV_SKIP_TAIL=1 python scripts/sweep_r2.py ldg 2 3 > gpurun_out/r2_sweep_ldg_notail.log 2>&1
cat gpurun_out/r2_sweep_ldg.log gpurun_out/r2_sweep_ip.log gpurun_out/r2_sweep_ldg_notail.log
