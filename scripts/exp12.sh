python -m pytest tests/test_dg1d_gpu.py tests/test_ldg_gpu.py -q -m gpu -x 2>&1 | grep -E "^E|passed|failed|Error" | head -12
