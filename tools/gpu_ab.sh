#!/bin/bash
# A/B visit: tests, bench with the default library and with an alternative build (GG_LIB), warm in-graph kernel timeline.
TAG=${1:-ab}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_${TAG}.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err; echo "bench rc=$?"; cut -c1-200 $O/bench_${TAG}.json
GG_LIB=$PWD/gif-gan_b200/lib/libgifgan_ub1.so timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}_ub1.json 2> $O/bench_${TAG}_ub1.err; echo "bench ub1 rc=$?"; cut -c1-200 $O/bench_${TAG}_ub1.json
timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}_again.json 2> $O/bench_${TAG}_again.err; echo "bench again rc=$?"; cut -c1-200 $O/bench_${TAG}_again.json
timeout 300 python tools/step_profile.py > $O/step_profile_${TAG}.log 2>&1; echo "step_profile rc=$?"; head -40 $O/step_profile_${TAG}.log
