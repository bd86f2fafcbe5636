#!/bin/bash
N=${1:-8}
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_w${N}_balanced.json 2> gpurun_out/r2_bench_w${N}_balanced.err
tail -n 2 gpurun_out/r2_bench_w${N}_balanced.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_w${N}_balanced.json') if l.startswith('{')][0])
print('balanced: ms_per_step', d['ms_per_step']*1e3, 'us value', d['value']/1e9, 'G bounds', d['slab_bounds'], 'per-rank ms (equal slabs)', d['per_rank_kernel_ms_equal_slabs'])
print('step_loop', d.get('step_loop'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 100 --warmup 10 --no-balance --no-extras 2>> gpurun_out/r2_bench_w${N}_balanced.err | cut -c1-200
