for v in old new2; do
export STEFAN_B200_LIB=$PWD/variants/$v.so
python scripts/quick_bench.py --order 1 --algo 4 --reps 2 > gpurun_out/plain_$v.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ipdg_apply_march -s 3 -c 1 -f -o gpurun_out/prof_p1_$v python scripts/quick_bench.py --order 1 --algo 4 --reps 2 > gpurun_out/ncu_$v.log 2>&1
done
