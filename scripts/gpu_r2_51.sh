#!/bin/bash
# attention kernel tests with the ping-pong / hand-pipelined engine 3 as the default, then same-box A/B of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -4
scripts/gpu_ab_step.sh r2g_ab_attn_pingpong
