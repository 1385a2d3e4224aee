#!/bin/bash
# quick end-to-end check: model parity tests, one denoising step timing (graph), per-shape conv_tc dump
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_model_gpu.py -q -s --timeout 900 -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "model rc=$? $(tail -1 gpurun_out/model.log)"
grep "\[parity\]" gpurun_out/model.log
timeout 600 python scripts/profile_step.py --reps 3 --graph 1 > gpurun_out/step_graph.log 2>&1; tail -1 gpurun_out/step_graph.log
PD_DUMP=gpurun_out/gemm_shapes9.csv timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/step.log 2>&1; tail -1 gpurun_out/step.log
