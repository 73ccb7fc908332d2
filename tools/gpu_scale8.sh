#!/bin/bash
# 8-GPU visit: the peer-memory exchange at N = 8 (correctness + time per all-reduce), then the weak-scaling bench at 8 and 1 ranks.
# Usage: gpurun --gpus 8 -- bash tools/gpu_scale8.sh <tag> [extra bench configs as ENV=V,ENV=V ...]
TAG=${1:-s8}; shift; O=gpurun_out; mkdir -p $O; LOG=$O/${TAG}_summary.log; : > $LOG
run() {  # N, port, env...
  local N=$1 P=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --no-extra --no-roofline --no-cpu-baseline --repeats 3 \
    > $O/${TAG}_n${N}_$P.json 2> $O/${TAG}_n${N}_$P.err
  (python -c "
import json
d=json.loads([l for l in open('$O/${TAG}_n${N}_$P.json').read().splitlines() if l.startswith('{')][-1]);print('N=$N $*', round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],4))" || grep -v "OMP_NUM\|\*\*\*\|Warning\|return func" $O/${TAG}_n${N}_$P.err | tail -6) 2>&1 | tee -a $LOG
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tools/dp_p2p_check.py --time > $O/${TAG}_p2p.log 2>&1
grep -E "dp_p2p_check|TIMING|Error|error|assert" $O/${TAG}_p2p.log | cut -c1-1500 | head -6 | tee -a $LOG
run 8 29702 GG_DP_P2P=1
P=29703
for cfg in "$@"; do run 8 $P ${cfg//,/ }; P=$((P+1)); done
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-extra --no-roofline --repeats 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]);print('N=1', round(d['ms_per_step'],4), round(d['value']))" | tee -a $LOG
