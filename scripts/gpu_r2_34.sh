#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "vertical_halo" 2>&1 | tail -12 | tee gpurun_out/r2_34_tests.log
timeout 600 python scripts/gemm_bench.py --vh 2>&1 | tee gpurun_out/r2_34_vh_bench.txt
