python -m pytest tests/test_fd_gpu.py -q -m gpu -k "131" 2>&1 | grep -E "^E|assert|Error" | head -12
for b in 1 2 4 8 32 100000; do echo -n "[$b] "; STEFAN_FD_BLOCKS_PER_SM=$b python scripts/quick_bench.py --fd 1 --reps 50 | head -1; done
