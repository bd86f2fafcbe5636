# last GPU call of the round: the complete GPU suite on the final tree, with the cut-cell study tables (-s)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_last_gpu_suite.log 2>&1
tail -n 3 gpurun_out/r2_last_gpu_suite.log
grep -E "^(plane|curved) (ip|ldg) p=[12] nelmts +(17|33)" gpurun_out/r2_last_gpu_suite.log | cut -c1-200
