#!/bin/bash
# where does the multi-GPU step time go: flags never waited for (1) / neither waited for nor posted (3)
# (development switches; x is static so the results stay right.  NEVER run "posted off, waits on": every wait would time out)
N=${1:-8}
set -x
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 100 --warmup 10 --no-extras $EXTRA 2>> gpurun_out/r2_matrix3_w$N.err | \
      python -c "import sys,json; l=[x for x in sys.stdin if x.startswith('{')]; d=json.loads(l[0]); print('$name', 'ms_per_step', round(d['ms_per_step']*1e3,2), 'us  value', round(d['value']/1e9,1), 'G  kernel_ms(rank0, plain)', round(d['roofline']['kernel_ms']*1e3,2))"
}
{
EXTRA="" run "full protocol" X=1
EXTRA="" run "never wait" STEFAN_DEV_PEER=1
EXTRA="" run "neither wait nor post" STEFAN_DEV_PEER=3
} > gpurun_out/r2_matrix3_w$N.log 2>&1
grep ms_per_step gpurun_out/r2_matrix3_w$N.log | grep -v "^+"
tail -n 3 gpurun_out/r2_matrix3_w$N.err

timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/dist_peer_check.py > gpurun_out/r2_peer_check_w$N.log 2>&1
tail -n 3 gpurun_out/r2_peer_check_w$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_w$N.json 2> gpurun_out/r2_bench_w$N.err
tail -n 2 gpurun_out/r2_bench_w$N.err; cut -c1-400 gpurun_out/r2_bench_w$N.json
