#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): parity tests on one GPU, then the data-parallel bench at 1 and N ranks.
TAG=${1:-m}; N=${2:-2}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_${TAG}.log
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err; echo "bench n1 rc=$?"; cut -c1-200 $O/bench_${TAG}_n1.json

timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 \
   > $O/bench_${TAG}_n$N.json 2> $O/bench_${TAG}_n$N.err; echo "bench n$N rc=$?"; cut -c1-200 $O/bench_${TAG}_n$N.json; tail -3 $O/bench_${TAG}_n$N.err
python tools/layer_kernels.py --reps 3 --launches 10 > $O/layers_$TAG.log 2>&1; echo "layers rc=$?"
