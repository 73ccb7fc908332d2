#!/bin/bash
# Short GPU visit: parity tests (stop at the first failure), then the config-2 bench line alone under each A/B switch setting
# given as an argument ("-" = defaults; several switches in one setting are joined by commas), interleaved twice.
# Usage (under gpurun): bash tools/gpu_switches.sh <tag> - GG_FUSE_LOSS_HEAD=0 GG_ZERO_ON_SIDE=0,GG_FUSE_LOSS_HEAD=0
TAG=${1:-sw}; shift
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_${TAG}.log | cut -c1-300
for rep in 1 2; do
  for sw in "$@"; do
    envs=$(echo "$sw" | tr ',' ' '); [ "$sw" = "-" ] && envs="GG_NOOP=1"
    timeout 200 env $envs python bench.py --no-cpu-baseline --no-extra --no-roofline > $O/bench_${TAG}_${sw}_$rep.json 2> $O/bench_${TAG}_${sw}_$rep.err
    echo "bench [$sw] rep $rep rc=$? $(python -c "import json,sys; d=json.load(open('$O/bench_${TAG}_${sw}_$rep.json')); print(d['ms_per_step'], d.get('gpu_launches_per_step'), d['losses'])" 2>&1 | tail -1)"
  done
done
