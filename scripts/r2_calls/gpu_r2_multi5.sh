#!/bin/bash
# round 2, 8 GPUs, last multi-GPU call: slab partition checks incl. the CG across slabs (NCCL callbacks), then the strong
# scaling of a 400 M-dof mesh (10000 x 10000 cells, p = 1; its one-GPU time is in gpurun_out/r2a_bench_400M_n1.json)
N=${1:-8}
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/dist_peer_check.py > gpurun_out/r2b_peer_check_w$N.log 2>&1
tail -n 6 gpurun_out/r2b_peer_check_w$N.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 50 --warmup 5 --nx 10000 --ny 10000 --no-cpu-baseline --no-extras \
    > gpurun_out/r2b_bench_400M_w$N.json 2> gpurun_out/r2b_bench_400M_w$N.err
tail -n 3 gpurun_out/r2b_bench_400M_w$N.err; cut -c1-1500 gpurun_out/r2b_bench_400M_w$N.json
