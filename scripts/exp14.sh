for mc in 0 1 2 4; do for o in 1 2 3; do echo -n "[cost $mc] "; STEFAN_MIXED_COST=$mc python scripts/quick_bench.py --order $o --reps 50; done; done
