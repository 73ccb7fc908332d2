#!/bin/bash
# Follow-up visit: run-to-run spread diagnosis, whole suite and bench / smoke on the defaults, ncu launch list of one step.
TAG=${1:-final2}
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
timeout 60 python tools/pack_ab_diag.py > $O/pack_ab_${TAG}.log 2>&1; echo "diag rc=$? t=$(( $(date +%s) - t0 ))"; cat $O/pack_ab_${TAG}.log | tail -n 6
timeout 120 python -m pytest tests -m gpu -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))"; tail -n 12 $O/pytest_${TAG}.log | cut -c1-300
timeout 55 python bench.py --no-cpu-baseline > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))"; cut -c1-200 $O/bench_${TAG}.json
timeout 40 python __graft_entry__.py smoke > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - t0 ))"; tail -n 1 $O/smoke_${TAG}.log
timeout 30 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_$TAG.log 2>&1 && \
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - t0 ))"
