#!/bin/bash
# Validation of the round's last kernel state: whole GPU suite exactly as the driver runs it, smoke, the bench line, the
# graph-replayed step and the ncu launch list of one eager step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu -s -p no:cacheprovider > gpurun_out/r2h_gpu_tests.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r2h_gpu_tests.log)"
grep "\[parity\]" gpurun_out/r2h_gpu_tests.log > gpurun_out/r2h_parity.txt; wc -l gpurun_out/r2h_parity.txt
timeout 900 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r2h_smoke.log)"
timeout 1500 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2h_bench.json
timeout 600 python scripts/profile_step.py --reps 3 --graph 1 > gpurun_out/r2h_step_graph.log 2>&1; tail -1 gpurun_out/r2h_step_graph.log
timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/r2h_step.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2h_launches_step.csv python scripts/profile_step.py > gpurun_out/r2h_ncu_step.log 2>&1
echo "launch list rc=$?"
