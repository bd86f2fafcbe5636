# order-2 fused step: b staged coalesced through shared memory (s1) or read per cell (s0), two CTAs of 128 registers
# (m2) or one CTA of up to 255 registers (m1) per SM.  Each variant: the fused-step tests, then the timing.
set -x
mkdir -p gpurun_out
for v in s1m2 s0m2 s1m1 s0m1; do
  export STEFAN_B200_LIB=$PWD/variants/p2step_$v.so
  echo "=== variant $v"
  timeout 300 python -m pytest tests/test_ipdg_gpu.py -q -x -k "step" 2>&1 | tail -n 2
  timeout 200 python scripts/probe_small.py step 2
done > gpurun_out/r2_p2step_variants.log 2>&1
cat gpurun_out/r2_p2step_variants.log
