#!/bin/bash
# Re-validate a restored tree: kernel + model parity, smoke, step timing (graph + eager), launch list, per-shape GEMM dump
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/kernels.log)"
timeout 1800 python -m pytest tests/test_model_gpu.py -q -s --timeout 900 -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "model rc=$? $(tail -1 gpurun_out/model.log)"
grep "\[parity\]" gpurun_out/model.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout 600 python scripts/profile_step.py --reps 3 --graph 1 > gpurun_out/step_graph.log 2>&1; tail -1 gpurun_out/step_graph.log
PD_DUMP=gpurun_out/gemm_shapes_d.csv timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/step.log 2>&1; tail -2 gpurun_out/step.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_d.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "launch list rc=$?"
