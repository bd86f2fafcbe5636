python -m pytest tests/test_ipdg_gpu.py -q -m gpu -k "cg_solve" 2>&1 | grep -E "^E|passed|failed|Error" | head -12
python -m pytest tests/test_fd_gpu.py -q -m gpu 2>&1 | grep -E "^E|assert|Error|passed|failed" | head -12
