python -m pytest tests/test_ipdg_gpu.py -q -m gpu 2>&1 | grep -E "FAILED|passed|failed|assert|Error" | head -40
for o in 1 2 3; do timeout 120 python scripts/quick_bench.py --order $o --algo 0 --reps 50;  timeout 120 python scripts/quick_bench.py --order $o --algo 0 --reps 50 --two-phase 0; done
