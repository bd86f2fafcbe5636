for n in 0 2000; do timeout 200 python scripts/quick_bench.py --ldg 1 --n $n --algo 0 --reps 50 --order 1; done
bash scripts/sweep.sh "x1 x2 x3 x4" "1" 1
bash scripts/sweep.sh "x1 x2 x3 x4" "1" 0
