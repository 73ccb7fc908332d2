#!/bin/bash
# Scaling visit (gpurun --gpus N): the data-parallel bench at the rank counts given -> gpurun_out/scale_<tag>.jsonl
TAG=${1:-sc}; shift
O=gpurun_out
mkdir -p $O
: > $O/scale_$TAG.jsonl
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline >> $O/scale_$TAG.jsonl 2> $O/scale_${TAG}_n1.err; echo "n1 rc=$?"
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 20 --warmup 3 \
       >> $O/scale_$TAG.jsonl 2> $O/scale_${TAG}_n$N.err; echo "n$N rc=$?"; tail -2 $O/scale_${TAG}_n$N.err
  fi
done
python - <<PY
import json
rows = [json.loads(l) for l in open("$O/scale_$TAG.jsonl") if l.startswith("{")]
base = rows[0]["value"] / rows[0]["n_gpus"]
for r in rows:
    print(r["n_gpus"], "GPUs", round(r["ms_per_step"], 3), "ms/step", round(r["value"]), "frames/s", "efficiency", round(r["value"] / (base * r["n_gpus"]), 3))
PY
