for rep in 1 2; do for v in old new2 new3; do
  echo -n "[$v] "; STEFAN_B200_LIB=$PWD/variants/$v.so timeout 120 python scripts/quick_bench.py --order 1 --algo 4 --reps 100 2>&1 | tail -1
done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
