#!/bin/bash
# Last GPU visit of a round on a small budget: whole parity suite with the batched re-pack switched on, the bench line
# with it off and on, latent-search throughput, smoke.  Every step has its own timeout; results land in gpurun_out/.
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
GG_PACK_BATCH=1 timeout 160 python -m pytest tests -m gpu -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))"; tail -n 25 $O/pytest_${TAG}.log | cut -c1-300
GG_PACK_BATCH=0 timeout 55 python bench.py --no-cpu-baseline --steps 60 > $O/bench_${TAG}_off.json 2> $O/bench_${TAG}_off.err; echo "bench off rc=$? t=$(( $(date +%s) - t0 ))"; cut -c1-200 $O/bench_${TAG}_off.json
GG_PACK_BATCH=1 timeout 55 python bench.py --no-cpu-baseline --steps 60 > $O/bench_${TAG}_on.json 2> $O/bench_${TAG}_on.err; echo "bench on rc=$? t=$(( $(date +%s) - t0 ))"; cut -c1-200 $O/bench_${TAG}_on.json
timeout 50 python tools/latent_bench.py --steps 100 > $O/latent_${TAG}.jsonl 2> $O/latent_${TAG}.err; echo "latent rc=$? t=$(( $(date +%s) - t0 ))"; cut -c1-400 $O/latent_${TAG}.jsonl; tail -n 5 $O/latent_${TAG}.err
GG_PACK_BATCH=1 timeout 40 python __graft_entry__.py smoke > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - t0 ))"; tail -n 3 $O/smoke_${TAG}.log
