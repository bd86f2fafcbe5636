# Round-2 FINAL evidence run (one GPU) with the code as committed at the end of the round: full GPU test suite, bench
# (both arms), launch list of the bench command, `ncu --set full` of the order-3 LDG kernel after its restructuring,
# timing table.  Outputs in gpurun_out/ (r2f_*).
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2f_gpu_tests.log
tail -n 2 gpurun_out/r2f_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_bench_short.json 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2f_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ncu_launches.log 2>&1
timeout 200 python scripts/quick_bench.py --ldg 1 --order 3 --reps 2 > gpurun_out/r2f_plain_ldg3.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ldg_apply_march -s 3 -c 1 -f -o gpurun_out/r2f_ldg_p3_small_march \
    python scripts/quick_bench.py --ldg 1 --order 3 --reps 2 > gpurun_out/r2f_ncu_ldg3.log 2>&1
{ for o in 1 2 3; do timeout 200 python scripts/quick_bench.py --order $o --reps 50; timeout 200 python scripts/quick_bench.py --order $o --reps 50 --two-phase 0; done
  timeout 200 python scripts/quick_bench.py --ldg 1 --reps 50; timeout 200 python scripts/quick_bench.py --ldg 1 --n 2000 --reps 50
  timeout 200 python scripts/quick_bench.py --fd 1 --reps 50; timeout 200 python scripts/quick_bench.py --blocks 1
  timeout 300 python scripts/probe_small.py time; timeout 300 python scripts/probe_small.py step; } > gpurun_out/r2f_quick.log 2>&1
cut -c1-600 gpurun_out/r2f_bench.json; tail -n 44 gpurun_out/r2f_quick.log
