for o in 2 3; do
python scripts/quick_bench.py --order $o --reps 2 > gpurun_out/plain_m$o.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ipdg_apply_march -s 3 -c 1 -f -o gpurun_out/prof_ip_p${o}_march python scripts/quick_bench.py --order $o --reps 2 > gpurun_out/ncu_m$o.log 2>&1
done
for o in 1 2 3; do python scripts/quick_bench.py --order $o --reps 50; done
