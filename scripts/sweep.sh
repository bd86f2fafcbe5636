# sweep.sh "A B C" "1 2 3" [two_phase]: time the production kernel of every variant library in variants/
for v in $1; do for o in $2; do
  echo -n "[$v] "; STEFAN_B200_LIB=$PWD/variants/$v.so timeout 120 python scripts/quick_bench.py --order $o --algo 0 --reps 50 --two-phase ${3:-1} 2>&1 | tail -1
done; done
