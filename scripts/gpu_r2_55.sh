#!/bin/bash
# hand-scheduled attention as default: kernel tests, model parity (the bf16 gates sit close to 1e-2), config-4 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -3
timeout 1800 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py tests/test_sampler_branches_gpu.py -m gpu -x -q -s 2>&1 | grep -E "parity|passed|failed|Error" > gpurun_out/r2g_parity_hand.txt; tail -3 gpurun_out/r2g_parity_hand.txt
grep -E "apply_model|config|smoke|50" gpurun_out/r2g_parity_hand.txt | head -20
for v in 0 1; do echo -n "config 4 PD_B200_ATTN3=$v: "; PD_B200_ATTN3=$v timeout 300 python scripts/profile_step.py --graph 1 --reps 8 --batch 16 --size 768 2>&1 | tail -1; done | tee gpurun_out/r2g_ab_config4_attn_hand.txt
