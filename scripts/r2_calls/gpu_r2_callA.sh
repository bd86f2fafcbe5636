# Round 2, one GPU: full GPU suite with the order-3 progressive LDG rows, the coalesced b staging of the p = 2 fused
# step, the slab CG (two host threads) and the new post-processing entry points; then the timings they change and
# the 400 M-dof single-GPU point of the strong-scaling pair.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2a_gpu_tests.log
tail -n 3 gpurun_out/r2a_gpu_tests.log
{ timeout 300 python scripts/probe_small.py time; timeout 300 python scripts/probe_small.py step;
  timeout 300 python scripts/quick_bench.py --ldg 1 --n 2000 --reps 50; } > gpurun_out/r2a_quick.log 2>&1
tail -n 32 gpurun_out/r2a_quick.log
timeout 600 python bench.py --nx 10000 --ny 10000 --no-cpu-baseline --no-extras > gpurun_out/r2a_bench_400M_n1.json 2> gpurun_out/r2a_bench_400M_n1.err
cut -c1-900 gpurun_out/r2a_bench_400M_n1.json; tail -n 3 gpurun_out/r2a_bench_400M_n1.err
