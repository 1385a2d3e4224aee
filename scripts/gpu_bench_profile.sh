#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py --steps ${STEPS:-1} --warmup ${WARMUP:-1} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python scripts/profile_step.py --reps 3 --graph 1 > gpurun_out/step_graph.log 2>&1; tail -1 gpurun_out/step_graph.log
timeout 600 python scripts/profile_step.py --reps 3 > gpurun_out/step.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "profile rc=$?"; cat gpurun_out/step.log | tail -2
