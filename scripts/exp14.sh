for o in 1 2 3; do for mc in 0 1 2 4 6; do echo -n "[cost_q $mc] "; STEFAN_MIXED_COST_Q=$mc python scripts/quick_bench.py --order $o --reps 50; done; done
