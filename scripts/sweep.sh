# sweep.sh "A B C" "1 2 3": time algo 4 of every variant library in variants/
for v in $1; do for o in $2; do
  echo -n "[$v] "; STEFAN_B200_LIB=$PWD/variants/$v.so timeout 120 python scripts/quick_bench.py --order $o --algo 4 2>&1 | tail -1
done; done
