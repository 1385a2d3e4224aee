#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -3
timeout 300 python scripts/attn_bench.py 2>&1 | grep -E " 3 \|| 6 \|" | grep " 40 " | tee gpurun_out/attn3_active.txt
