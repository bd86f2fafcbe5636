#!/bin/bash
# round 2, multi-GPU timing matrix of the strong-scaling step (no extras): sync mode x mailboxes x edge-tile cost
N=${1:-8}
set -x
mkdir -p gpurun_out
run() {  # name, extra env, extra args
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 100 --warmup 10 --no-extras $EXTRA 2>> gpurun_out/r2_matrix_w$N.err | \
      python -c "import sys,json; l=[x for x in sys.stdin if x.startswith('{')]; d=json.loads(l[0]); print('$name', 'ms_per_step', round(d['ms_per_step']*1e3,2), 'us  value', round(d['value']/1e9,1), 'G  parity', d['parity_max_rel'], 'graph', d['config']['cuda_graph'])"
}
{
EXTRA="" run "chained+push edge8" X=1
EXTRA="" run "chained+push edge0" STEFAN_EDGE_COST_16=0
EXTRA="" run "chained+push edge16" STEFAN_EDGE_COST_16=64
EXTRA="--sync ready" run "ready+push edge8" X=1
EXTRA="--sync ready --no-push" run "ready+poll edge8" X=1
EXTRA="--no-push" run "chained+poll edge8" X=1
EXTRA="--no-graph" run "chained+push nograph" X=1
EXTRA="" run "chained+push nopdl" STEFAN_PDL=0
} > gpurun_out/r2_matrix_w$N.log 2>&1
cat gpurun_out/r2_matrix_w$N.log
tail -n 3 gpurun_out/r2_matrix_w$N.err
