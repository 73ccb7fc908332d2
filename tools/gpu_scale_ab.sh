#!/bin/bash
# N-rank data-parallel bench under exchange variants (env assignments, ';'-separated groups as arguments)
TAG=${1:-sa}; N=${2:-8}; shift; shift
O=gpurun_out
mkdir -p $O
: > $O/scaleab_$TAG.jsonl
i=0
for sw in "default" "$@"; do
  i=$((i + 1))
  envs=$(echo "$sw" | tr ';' ' ')
  [ "$sw" = "default" ] && envs="GG_NOP=1"
  timeout 300 env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + i)) bench.py --gpus $N --steps 20 --warmup 3 \
     > $O/scaleab_${TAG}_$i.json 2> $O/scaleab_${TAG}_$i.err; echo "$sw rc=$?"
  python - <<PY
import json
try:
    r = json.loads(open("$O/scaleab_${TAG}_$i.json").read().strip().splitlines()[-1])
    print("   ", "$sw", r["n_gpus"], "GPUs", round(r["ms_per_step"], 3), "ms/step", round(r["value"]), "frames/s")
except Exception as e:
    print("   ", "$sw", "no result", e)
PY
  grep -m3 -i "nvls\|Algo\|Using network" $O/scaleab_${TAG}_$i.err | cut -c1-160
done
