#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_30_tests.log
scripts/gpu_ab_step.sh r2_30_ab_attn
