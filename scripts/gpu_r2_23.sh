#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -5 | tee gpurun_out/r2_23_tests.log
timeout 300 python scripts/attn_d160.py 2>&1 | tee gpurun_out/r2_23_attn_d160.txt
