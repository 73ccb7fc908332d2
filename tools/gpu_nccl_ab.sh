#!/bin/bash
# NCCL knob A/B of the data-parallel bench.  Usage: gpurun --gpus N -- bash tools/gpu_nccl_ab.sh <tag> <N>
TAG=${1:-nccl}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
LOG=$O/${TAG}_summary.log; : > $LOG
timeout 200 python tools/dp_check.py single > $O/${TAG}_single.log 2>&1
timeout 300 $TR --master-port 29530 tools/dp_check.py dp > $O/${TAG}_check.log 2>&1; grep -E "worst|dcgan|Error|assert" $O/${TAG}_check.log | cut -c1-300 | head -6 | tee -a $LOG
i=0
for cfg in "X=1" "NCCL_MAX_NCHANNELS=2" "NCCL_MAX_NCHANNELS=4" "NCCL_MAX_NCHANNELS=8" "NCCL_PROTO=Simple" "NCCL_PROTO=LL128" "NCCL_ALGO=Tree" "GG_DP_EARLY_MB=0" "GG_DP_EARLY_MB=8" "NCCL_MAX_NCHANNELS=4 GG_DP_EARLY_MB=8"; do
  i=$((i+1))
  env $cfg timeout 300 $TR --master-port 2954$i bench.py --gpus $N --no-extra --no-roofline --repeats 3 > $O/${TAG}_n${N}_$i.json 2> $O/${TAG}_n${N}_$i.err
  (python -c "
import json
d=json.loads([l for l in open('$O/${TAG}_n${N}_$i.json').read().splitlines() if l.startswith('{')][-1]);print('N=$N $cfg', round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['ms_per_step'],4))" || tail -3 $O/${TAG}_n${N}_$i.err) 2>&1 | tee -a $LOG
done
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-extra --no-roofline --repeats 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]);print('N=1', round(d['ms_per_step'],4), round(d['value']))" | tee -a $LOG
