# Round-1 evidence run (one GPU): full GPU test suite, bench, launch list of the bench command,
# one `ncu --set full` capture per production kernel.  Outputs in gpurun_out/.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r1_gpu_tests.log
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launches.log 2>&1
for o in 1 2 3; do
python scripts/quick_bench.py --order $o --reps 2 > gpurun_out/r1_plain_p$o.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ipdg_apply_march -s 3 -c 1 -f -o gpurun_out/r1_ipdg_p${o}_march \
    python scripts/quick_bench.py --order $o --reps 2 > gpurun_out/r1_ncu_p$o.log 2>&1
done
for o in 1 2 3; do
python scripts/quick_bench.py --ldg 1 --order $o --n 2000 --reps 2 > gpurun_out/r1_plain_ldg$o.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ldg_apply_march -s 3 -c 1 -f -o gpurun_out/r1_ldg_p${o}_march \
    python scripts/quick_bench.py --ldg 1 --order $o --n 2000 --reps 2 > gpurun_out/r1_ncu_ldg$o.log 2>&1
done
python scripts/quick_bench.py --fd 1 --reps 2 > gpurun_out/r1_plain_fd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fd_apply -s 3 -c 2 -f -o gpurun_out/r1_fd \
    python scripts/quick_bench.py --fd 1 --reps 2 > gpurun_out/r1_ncu_fd.log 2>&1
{ for o in 1 2 3; do python scripts/quick_bench.py --order $o --reps 50; python scripts/quick_bench.py --order $o --reps 50 --two-phase 0; done
  python scripts/quick_bench.py --ldg 1 --reps 50; python scripts/quick_bench.py --ldg 1 --n 2000 --reps 50
  python scripts/quick_bench.py --fd 1 --reps 50; python scripts/quick_bench.py --blocks 1; } > gpurun_out/r1_quick.log 2>&1
tail -3 gpurun_out/r1_gpu_tests.log; cat gpurun_out/r1_bench.json gpurun_out/r1_quick.log
