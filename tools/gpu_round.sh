#!/bin/bash
# One GPU-box visit: parity tests, bench line, launch list, full ncu capture of the conv kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$TAG.txt
python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log | cut -c1-300
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $O/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "ref rc=$?"
GG_PROF=1 python tools/tc_sweep.py --inproc > $O/tc_prof_$TAG.log 2>&1
python tools/layer_kernels.py --reps 3 --launches 10 > $O/layers_$TAG.log 2>&1; echo "layers rc=$?"
# launch list of the bench command's timed workload (the other configs / the per-kernel tables are left out of the ncu pass: every
# kernel is serialised and replayed under ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-roofline > $O/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-roofline > $O/ncu_launches_$TAG.log 2>&1
python tools/layer_kernels.py > $O/plain_layers_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'tc_pixgemm|tc_wgrad|c3m_|c3_' -c 64 -o $O/prof_layers_$TAG \
    python tools/layer_kernels.py > $O/ncu_full_$TAG.log 2>&1
ncu -i $O/prof_layers_$TAG.ncu-rep --page raw --csv > $O/prof_layers_${TAG}_raw.csv 2>/dev/null
SZ=$(stat -c %s $O/prof_layers_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 45000000 ]; then rm -f $O/prof_layers_$TAG.ncu-rep; echo "ncu-rep too large ($SZ), kept raw csv only"; fi
ls -la $O | tail -20
