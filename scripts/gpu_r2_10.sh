#!/bin/bash
# resident-B schedule: kernel tests, then the short-K shape sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "resident or tcgen05_bf16" 2>&1 | tail -15 > gpurun_out/r2_10_tests.log; cat gpurun_out/r2_10_tests.log
timeout 600 python scripts/gemm_bench.py --bres > gpurun_out/r2_10_bres_bench.txt 2>&1; cat gpurun_out/r2_10_bres_bench.txt
