#!/bin/bash
# new sampler-branch / diffusers-route tests + A/B of the software-pipelined exponential section (attention_tc variants)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_branches_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert" | tail -40
scripts/gpu_attn_variants.sh 2>&1 | tail -8
cp gpurun_out/attn_variants.txt gpurun_out/r2g_attn_swp_variants.txt
