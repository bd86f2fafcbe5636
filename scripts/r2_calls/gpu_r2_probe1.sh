#!/bin/bash
# round 2, probe 1 (one GPU): small-mesh timings, per-CTA phase times, ncu of the LDG kernels at the 10 M-dof meshes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python scripts/probe_small.py time > gpurun_out/r2_probe_time.log 2>&1
STEFAN_DEV_SKIP_TAIL=1 python scripts/probe_small.py time > gpurun_out/r2_probe_time_notail.log 2>&1
for o in 1 2 3; do python scripts/probe_small.py cta ldg $o; done > gpurun_out/r2_probe_cta.log 2>&1
for o in 1 3; do python scripts/probe_small.py cta ip $o 8; done >> gpurun_out/r2_probe_cta.log 2>&1
python scripts/probe_small.py cta ip 1 1 >> gpurun_out/r2_probe_cta.log 2>&1
for o in 1 2 3; do
  python scripts/quick_bench.py --ldg 1 --order $o --reps 5 > gpurun_out/r2_plain_ldg$o.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:ldg_apply_march -s 4 -c 1 \
      -o gpurun_out/r2_ldg_p${o}_small -f python scripts/quick_bench.py --ldg 1 --order $o --reps 5 > gpurun_out/r2_ncu_ldg$o.log 2>&1
done
tail -n 30 gpurun_out/r2_probe_time.log gpurun_out/r2_probe_cta.log
