python -m pytest tests/test_ipdg_gpu.py -q -m gpu 2>&1 | grep -E "FAILED|passed|failed|assert|Error" | head -40
bash scripts/sweep.sh "u0 u1" "1 2 3" 1
bash scripts/sweep.sh "u0 u1" "1 2 3" 0
