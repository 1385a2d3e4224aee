#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv or linear or geglu" 2>&1 | tail -3
scripts/gpu_ab_step.sh r2_42_ab_compact_staging
