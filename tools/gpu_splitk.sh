#!/bin/bash
# split-K visit: parity, then the sweep.  Usage: gpurun -- bash tools/gpu_splitk.sh <tag>
TAG=${1:-sk}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --no-header 2>&1 | tail -15 | cut -c1-300
GG_PROF=1 timeout 600 python tools/splitk_sweep.py > $O/${TAG}_sweep.log 2>&1; tail -3 $O/${TAG}_sweep.log | cut -c1-300
for i in 1 2; do for sk in 1 auto; do GG_TC_SPLITK=$sk timeout 300 python bench.py --no-cpu-baseline --no-extra --no-roofline --steps 20 --repeats 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('BENCH splitk=$sk', d['ms_per_step'], d['e2e']['ms_per_step'], d['losses'])"; done; done 2>&1 | tee $O/${TAG}_bench.log
