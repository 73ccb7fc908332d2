#!/bin/bash
# Multi-GPU visit: DP consistency against a single GPU, then bench A/B of the exchange options.
# Usage: gpurun --gpus 2 -- bash tools/gpu_dp.sh <tag> <N> [check|bench|full ...]      (everything is also written to gpurun_out/<tag>_*.log)
TAG=${1:-dp}; N=${2:-2}; shift 2; WHAT=${*:-check bench full}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
LOG=$O/${TAG}_summary.log; : > $LOG
for what in $WHAT; do case $what in
p2p)   # the exchange kernel alone: correctness + time per all-reduce, at two grid sizes
  for nb in 64 128; do
    GG_DP_BLOCKS=$nb timeout 300 $TR --master-port 2952$((nb/64)) tools/dp_p2p_check.py --time > $O/${TAG}_p2p_$nb.log 2>&1
    echo "== GG_DP_BLOCKS=$nb" | tee -a $LOG; grep -E "dp_p2p_check|TIMING|Error|error|assert" $O/${TAG}_p2p_$nb.log | cut -c1-1500 | head -8 | tee -a $LOG
  done;;
check)
  timeout 300 $TR --master-port 29529 tools/dp_p2p_check.py --time > $O/${TAG}_p2p.log 2>&1; grep -E "dp_p2p_check|TIMING|Error|error|assert" $O/${TAG}_p2p.log | cut -c1-600 | head -8 | tee -a $LOG
  timeout 200 python tools/dp_check.py single > $O/${TAG}_single.log 2>&1; tail -1 $O/${TAG}_single.log | cut -c1-200 | tee -a $LOG
  i=0
  for cfg in "GG_DP_P2P=1" "GG_DP_P2P=1 GG_DP_OVERLAP_UPDATE=0" "GG_DP_P2P=0 GG_DP_GRAD_DTYPE=bf16"; do
    i=$((i+1)); echo "== check $cfg" | tee -a $LOG
    env $cfg timeout 300 $TR --master-port 2953$i tools/dp_check.py dp > $O/${TAG}_check$i.log 2>&1
    grep -E "worst|dcgan|Error|error|assert" $O/${TAG}_check$i.log | cut -c1-400 | head -12 | tee -a $LOG
  done;;
bench)
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-extra --no-roofline > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err
  python -c "import json;d=json.loads([l for l in open('$O/${TAG}_n1.json').read().splitlines() if l.startswith('{')][-1]);print('N=1', d['ms_per_step'], d['value'])" | tee -a $LOG
  i=0
  for cfg in ${DP_BENCH_CFGS:-"GG_DP_P2P=1" "GG_DP_P2P=1,GG_DP_OVERLAP_UPDATE=0" "GG_DP_P2P=0,GG_DP_EARLY_MB=0" "GG_DP_P2P=0"}; do
    cfg=${cfg//,/ }
    i=$((i+1))
    env $cfg timeout 300 $TR --master-port 2954$i bench.py --gpus $N --no-extra --no-roofline > $O/${TAG}_n${N}_$i.json 2> $O/${TAG}_n${N}_$i.err
    (python -c "import json;d=json.loads([l for l in open('$O/${TAG}_n${N}_$i.json').read().splitlines() if l.startswith('{')][-1]);print('N=$N $cfg', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])" || grep -v "OMP_NUM\|\*\*\*\|Warning\|return func" $O/${TAG}_n${N}_$i.err | tail -8) 2>&1 | tee -a $LOG
  done;;
full)   # default configuration incl. config 4 (video GAN, global 256 clips) at N ranks
  timeout 400 $TR --master-port 29533 bench.py --gpus $N --no-roofline > $O/${TAG}_n${N}_full.json 2> $O/${TAG}_n${N}_full.err
  (python -c "
import json;d=json.loads([l for l in open('$O/${TAG}_n${N}_full.json').read().splitlines() if l.startswith('{')][-1]);print('N=$N full', d['ms_per_step'], d['value']);
for e in d['extra']: print('   extra', {k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items() if k not in ('workload','repeats_ms_per_step','e2e')})" || tail -5 $O/${TAG}_n${N}_full.err) 2>&1 | tee -a $LOG;;
esac; done
