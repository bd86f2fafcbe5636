# final tree: the march-kernel parity tests once more (march_plan.h was touched after the last full suite), FD tests
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py tests/test_fd_gpu.py -q -x 2>&1 | tail -n 3 > gpurun_out/r2_final_check.log
cat gpurun_out/r2_final_check.log
