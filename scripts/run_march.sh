set -x
timeout 600 python -m pytest tests/test_ipdg_gpu.py -x -q -m gpu 2>&1 | tail -15
for o in 1 2 3; do timeout 120 python scripts/quick_bench.py --order $o --algo 0; timeout 120 python scripts/quick_bench.py --order $o --algo 0; done
