# FD kernel: fast path for all-central vectors (bit-identical to the general path): tests, then timings
set -x
mkdir -p gpurun_out
{ timeout 300 python -m pytest tests/test_fd_gpu.py tests/test_full_size_gpu.py -q -x -k "fd or FD" 2>&1 | tail -n 3
  timeout 120 python scripts/quick_bench.py --fd 1 --reps 50; } > gpurun_out/r2_fd_fastpath.log 2>&1
cat gpurun_out/r2_fd_fastpath.log
