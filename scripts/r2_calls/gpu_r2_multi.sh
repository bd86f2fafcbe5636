#!/bin/bash
# round 2, multi-GPU: correctness of the slab partition (kept under profiles/), then bench.py at N ranks
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/dist_peer_check.py > gpurun_out/r2_peer_check_w$N.log 2>&1
tail -n 5 gpurun_out/r2_peer_check_w$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_w$N.json 2> gpurun_out/r2_bench_w$N.err
tail -n 3 gpurun_out/r2_bench_w$N.err; cat gpurun_out/r2_bench_w$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 50 --warmup 5 --no-graph --no-extras > gpurun_out/r2_bench_w${N}_nograph.json 2>> gpurun_out/r2_bench_w$N.err
cat gpurun_out/r2_bench_w${N}_nograph.json
