for v in $1; do for o in $2; do for n in 0 2000; do
  echo -n "[$v] "; STEFAN_B200_LIB=$PWD/variants/$v.so timeout 120 python scripts/quick_bench.py --ldg 1 --order $o --n $n --algo 2 --reps 50 2>&1 | tail -1
done; done; done
