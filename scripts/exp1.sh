python -m pytest tests/test_ipdg_gpu.py -q -m gpu 2>&1 | grep -E "FAILED|passed|failed|assert|Error" | head -40
python -m pytest tests/test_ipdg_gpu.py -q -m gpu -k "test_apply_matches_oracle or two_slabs or stream_equals" 2>&1 | grep -E "FAILED|passed|failed|assert|Error" | head -40
for o in 1 2 3; do timeout 120 python scripts/quick_bench.py --order $o --algo 4; done
