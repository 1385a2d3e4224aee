#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -k "conv_tcgen05" -q -x --timeout 200 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/gemm_bench.py --modes 2>&1 | tee gpurun_out/gemm_modes.txt
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1; cat gpurun_out/gemm_bench.txt
