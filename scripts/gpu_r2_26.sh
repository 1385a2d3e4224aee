#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "short_keys" 2>&1 | tail -8 | tee gpurun_out/r2_26_tests.log
timeout 300 python scripts/attn_d160.py 2>&1 | tee gpurun_out/r2_26_attn.txt
