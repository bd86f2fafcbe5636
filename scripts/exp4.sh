for v in n3_nst3 n3_nst4 n3_nst8; do
  echo -n "[$v] "; STEFAN_B200_LIB=$PWD/variants/$v.so timeout 120 python scripts/quick_bench.py --order 1 --algo 4 --reps 50 2>&1 | tail -1
done
for v in old new3; do
STEFAN_B200_LIB=$PWD/variants/$v.so ncu --cache-control none --clock-control none -k regex:ipdg_apply_march -s 12 -c 2 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max python scripts/quick_bench.py --order 1 --algo 4 --reps 20 2>&1 | grep -E "time_duration|dram__|lts__|cycles" 
done
