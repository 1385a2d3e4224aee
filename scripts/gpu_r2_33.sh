#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "persistent" 2>&1 | tail -8 | tee gpurun_out/r2_33_tests.log
timeout 600 python scripts/attn_d160.py 2>&1 | tail -4 | tee gpurun_out/r2_33_attn.txt
