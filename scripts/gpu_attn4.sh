#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "four_groups or three_groups or outgrow" 2>&1 | tail -4
timeout 300 python scripts/attn_bench.py 2>&1 | grep " 40 " | tee gpurun_out/attn4_bench.txt
