# Round-2 evidence run (one GPU): full GPU test suite, bench (both arms), launch list of the bench command,
# `ncu --set full` captures of the dominant kernel and of the small-mesh kernels VERDICT r1 asked for.  Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests.log
tail -n 2 gpurun_out/r2_gpu_tests.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/quick_bench.py --order 1 --reps 2 > gpurun_out/r2_plain_p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ipdg_apply_march -s 3 -c 1 -f -o gpurun_out/r2_ipdg_p1_march \
    python scripts/quick_bench.py --order 1 --reps 2 > gpurun_out/r2_ncu_p1.log 2>&1
python scripts/quick_bench.py --order 1 --nx 625 --ny 5000 --reps 2 > gpurun_out/r2_plain_p1s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ipdg_apply_march -s 3 -c 1 -f -o gpurun_out/r2_ipdg_p1_slab8_march \
    python scripts/quick_bench.py --order 1 --nx 625 --ny 5000 --reps 2 > gpurun_out/r2_ncu_p1s.log 2>&1
for o in 1 2 3; do
python scripts/quick_bench.py --ldg 1 --order $o --reps 2 > gpurun_out/r2_plain_ldg$o.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ldg_apply_march -s 3 -c 1 -f -o gpurun_out/r2_ldg_p${o}_small_march \
    python scripts/quick_bench.py --ldg 1 --order $o --reps 2 > gpurun_out/r2_ncu_ldg$o.log 2>&1
done
{ for o in 1 2 3; do python scripts/quick_bench.py --order $o --reps 50; python scripts/quick_bench.py --order $o --reps 50 --two-phase 0; done
  python scripts/quick_bench.py --ldg 1 --reps 50; python scripts/quick_bench.py --ldg 1 --n 2000 --reps 50
  python scripts/quick_bench.py --fd 1 --reps 50; python scripts/quick_bench.py --blocks 1
  python scripts/probe_small.py time; python scripts/probe_small.py step; } > gpurun_out/r2_quick.log 2>&1
cat gpurun_out/r2_bench.json | cut -c1-600; tail -n 40 gpurun_out/r2_quick.log
