#!/bin/bash
# Short GPU visit: parity tests, microbenchmarks, per-layer kernel times, bench lines with and without PDL.
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
export GG_PDL=0
python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}_nopdl.log 2>&1; echo "pytest(nopdl) rc=$?"; tail -12 $O/pytest_${TAG}_nopdl.log
[ -x tools/bin/mma_bench ] && timeout 120 tools/bin/mma_bench > $O/mma_bench_$TAG.log 2>&1; echo "mma_bench rc=$?"
python tools/layer_kernels.py --reps 3 --launches 10 > $O/layers_$TAG.log 2>&1; echo "layers rc=$?"
GG_PROF=1 python tools/tc_sweep.py --inproc > $O/tc_prof_$TAG.log 2>&1
python bench.py --no-cpu-baseline > $O/bench_${TAG}_nopdl.json 2> $O/bench_${TAG}_nopdl.err; echo "bench(nopdl) rc=$?"; cut -c1-220 $O/bench_${TAG}_nopdl.json
export GG_PDL=1
python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}_pdl.log 2>&1; echo "pytest(pdl) rc=$?"; tail -12 $O/pytest_${TAG}_pdl.log
python bench.py --no-cpu-baseline > $O/bench_${TAG}_pdl.json 2> $O/bench_${TAG}_pdl.err; echo "bench(pdl) rc=$?"; cut -c1-220 $O/bench_${TAG}_pdl.json; tail -5 $O/bench_${TAG}_pdl.err
