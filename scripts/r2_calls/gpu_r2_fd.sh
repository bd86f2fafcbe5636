# FD kernel with the one-iteration-ahead prefetch: tests, then timings against the variant without it and over the grid knob
set -x
mkdir -p gpurun_out
{ timeout 300 python -m pytest tests/test_fd_gpu.py tests/test_full_size_gpu.py -q -x -k "fd or FD" 2>&1 | tail -n 3
  for b in 32 16 8; do echo "== prefetch, blocks per SM $b"; STEFAN_FD_BLOCKS_PER_SM=$b timeout 120 python scripts/quick_bench.py --fd 1 --reps 50; done
  for b in 32 16; do echo "== no prefetch, blocks per SM $b"; STEFAN_B200_LIB=$PWD/variants/fd_nopf.so STEFAN_FD_BLOCKS_PER_SM=$b timeout 120 python scripts/quick_bench.py --fd 1 --reps 50; done
} > gpurun_out/r2_fd_prefetch.log 2>&1
cat gpurun_out/r2_fd_prefetch.log
