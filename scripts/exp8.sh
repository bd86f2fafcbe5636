timeout 900 python -m pytest tests/test_ldg_gpu.py -q -m gpu -x 2>&1 | grep -E "FAILED|passed|failed|assert|Error" | head -20
for n in 0 2000; do for a in 0 1; do timeout 200 python scripts/quick_bench.py --ldg 1 --n $n --algo $a --reps 50; done; done
