#!/bin/bash
# Final visit of a session with a small GPU budget left: parity tests, smoke, the full bench line + the CPU arm, the ncu launch
# list of the bench command's timed workload, a `--set full` capture of the kernels that are NEW since the last full layer capture
# (the tcgen05 layer capture of tools/gpu_round.sh stays valid while tc_tapgemm.cu / tc_common.cuh / common.cuh are unchanged:
# bench.py checks that), and the warm in-graph kernel timeline.
# Usage (under gpurun): bash tools/gpu_final3.sh <tag>
TAG=${1:-fin}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$TAG.txt
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log | cut -c1-300
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-300 $O/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "ref rc=$?"
timeout 200 python tools/layer_kernels.py --reps 3 --launches 10 > $O/layers_$TAG.log 2>&1; echo "layers rc=$?"
timeout 200 python tools/step_profile.py > $O/timeline_$TAG.log 2>&1; echo "timeline rc=$?"
GG_PDL=0 timeout 200 python tools/step_profile.py --timeline > $O/timeline_nopdl_$TAG.log 2>&1; echo "timeline nopdl rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-roofline > $O/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-roofline > $O/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'loss_head|act_bwd_bias|c3m_wgrad|adam_dev' -c 14 -o $O/prof_new_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-roofline > $O/ncu_new_$TAG.log 2>&1
echo "ncu new rc=$?"
ncu -i $O/prof_new_$TAG.ncu-rep --page raw --csv > $O/prof_new_${TAG}_raw.csv 2>/dev/null
ls -la $O | tail -16
