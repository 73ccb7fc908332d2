#!/bin/bash
# Short GPU visit: parity tests, per-layer kernel times, clock64 profile, bench lines under A/B switches given as args.
TAG=${1:-q}; shift
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_${TAG}.log
timeout 200 python tools/layer_kernels.py --reps 3 --launches 10 > $O/layers_$TAG.log 2>&1; echo "layers rc=$?"
GG_PROF=1 timeout 200 python tools/tc_sweep.py --inproc > $O/tc_prof_$TAG.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err; echo "bench rc=$?"; cut -c1-220 $O/bench_${TAG}.json
for sw in "$@"; do
  timeout 300 env $sw python bench.py --no-cpu-baseline > $O/bench_${TAG}_$sw.json 2> $O/bench_${TAG}_$sw.err; echo "bench $sw rc=$?"; cut -c1-220 $O/bench_${TAG}_$sw.json
done
