# cut-cell regime with the reference's own geometry restated (Q2-interpolated level set, height-function rules, legacy
# boundary form where the stored files used it): the GPU tests, with the studies' tables (the tests print them)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cutcell_gpu.py -q -x -s > gpurun_out/r2_cut_cell_gpu_final.log 2>&1
grep -E "nelmts|passed|failed|Error|assert" gpurun_out/r2_cut_cell_gpu_final.log | cut -c1-220
