#!/bin/bash
# BASELINE config 5 (batch sweep of the image DCGAN step): one bench line per batch size -> gpurun_out/sweep_<tag>.jsonl
TAG=${1:-s}
O=gpurun_out
mkdir -p $O
: > $O/sweep_$TAG.jsonl
for B in 16 32 64 128 256 512; do
  timeout 300 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu-baseline >> $O/sweep_$TAG.jsonl 2>> $O/sweep_$TAG.err
  echo "batch $B rc=$?"
done
python - <<PY
import json
for l in open("$O/sweep_$TAG.jsonl"):
    r = json.loads(l)
    print(r["config"]["global_batch"], round(r["ms_per_step"], 3), "ms", round(r["value"]), "frames/s", "step TFLOP/s", round(r["roofline"]["step"]["achieved_tflops"], 1),
          "frac_sustained", round(r["roofline"]["step"]["frac_of_sustained"], 3))
PY
