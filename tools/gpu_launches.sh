#!/bin/bash
# Launch list only (ncu, minimal metrics) of the bench step + tests + A/B bench lines given as env assignments.
TAG=${1:-l}; shift
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_${TAG}.log
python bench.py --no-cpu-baseline > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err; echo "bench rc=$?"; cut -c1-200 $O/bench_${TAG}.json
for sw in "$@"; do
  env $sw python bench.py --no-cpu-baseline > $O/bench_${TAG}_$sw.json 2> $O/bench_${TAG}_$sw.err; echo "bench $sw rc=$?"; cut -c1-200 $O/bench_${TAG}_$sw.json
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1
python tools/layer_kernels.py --reps 3 --launches 10 > $O/layers_$TAG.log 2>&1; echo "layers rc=$?"
