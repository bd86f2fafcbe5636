#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cutcell_gpu.py tests/test_blocks_gpu.py tests/test_ldg_gpu.py -q -x -m gpu > gpurun_out/r2_tests9.log 2>&1
tail -n 15 gpurun_out/r2_tests9.log
python examples/cylinder_ip_convergence.py > gpurun_out/r2_cylinder.log 2>&1
python examples/cylinder_ip_convergence.py robin > gpurun_out/r2_cylinder_robin.log 2>&1
cat gpurun_out/r2_cylinder.log gpurun_out/r2_cylinder_robin.log
python examples/uncut_ldg_convergence.py --device-solve > gpurun_out/r2_ldg_native_solve.log 2>&1
tail -n 14 gpurun_out/r2_ldg_native_solve.log
