set -x
python scripts/quick_bench.py > gpurun_out/qb_all.log 2>&1
python scripts/quick_bench.py --ldg 1 >> gpurun_out/qb_all.log 2>&1
python scripts/quick_bench.py --ldg 1 --n 2000 >> gpurun_out/qb_all.log 2>&1
for o in 2 3; do
python scripts/quick_bench.py --order $o --reps 2 > gpurun_out/plain_ip$o.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ipdg_apply_stream -s 3 -c 1 -f -o gpurun_out/prof_ip_p${o}_stream python scripts/quick_bench.py --order $o --reps 2 > gpurun_out/ncu_ip$o.log 2>&1
done
for o in 1 3; do
python scripts/quick_bench.py --ldg 1 --order $o --reps 2 > gpurun_out/plain_ldg$o.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ldg_apply -s 3 -c 1 -f -o gpurun_out/prof_ldg_p${o} python scripts/quick_bench.py --ldg 1 --order $o --reps 2 > gpurun_out/ncu_ldg$o.log 2>&1
done
cat gpurun_out/qb_all.log
