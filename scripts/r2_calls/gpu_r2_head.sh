# exception cells taken as a static share at the START of the march kernels (behind the first column copies) instead of
# a queue at their end: parity tests, then A/B timings (STEFAN_HEAD_SHARE=0 is the old behaviour)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py tests/test_full_size_gpu.py -q -x 2>&1 | tail -n 4 > gpurun_out/r2_head_tests.log
cat gpurun_out/r2_head_tests.log
for hs in 1 0; do
  echo "=== STEFAN_HEAD_SHARE=$hs"
  STEFAN_HEAD_SHARE=$hs timeout 300 python scripts/probe_small.py time
  STEFAN_HEAD_SHARE=$hs timeout 200 python scripts/quick_bench.py --ldg 1 --n 2000 --reps 50
done > gpurun_out/r2_head_share.log 2>&1
grep -v "^+" gpurun_out/r2_head_share.log
