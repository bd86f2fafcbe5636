#!/bin/bash
# round 2: compute-sanitizer over the march kernels (one tool per gpurun call, as the profiling guide asks)
# usage: bash scripts/gpu_r2_sanitize.sh racecheck|memcheck|synccheck|initcheck
TOOL=${1:-racecheck}
set -x
mkdir -p gpurun_out
SEL='degenerate or two_slabs_equal_one or hundred_fused_steps or column_ranges_compose'
# plain run first (must exit 0 before a tool is put on it)
timeout 600 python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py -q -x -m gpu -k "$SEL" > gpurun_out/r2_sanitize_plain.log 2>&1 || { tail -n 20 gpurun_out/r2_sanitize_plain.log; exit 1; }
tail -n 2 gpurun_out/r2_sanitize_plain.log
EXTRA=""
if [ "$TOOL" = "memcheck" ]; then EXTRA="--leak-check no"; fi
timeout 1500 compute-sanitizer --tool $TOOL $EXTRA --print-limit 40 --log-file gpurun_out/r2_sanitize_$TOOL.log \
    python -m pytest tests/test_ipdg_gpu.py tests/test_ldg_gpu.py -q -x -m gpu -k "$SEL" > gpurun_out/r2_sanitize_${TOOL}_pytest.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitize_${TOOL}_pytest.log
tail -n 5 gpurun_out/r2_sanitize_${TOOL}_pytest.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/r2_sanitize_$TOOL.log | sort | uniq -c | sort -rn | head -20
