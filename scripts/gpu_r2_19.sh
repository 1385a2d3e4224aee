#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "stream_k or tcgen05" 2>&1 | tail -3
echo "== base"; PD_B200_LIB=build/variants/libpd_base.so timeout 300 python scripts/gemm_bench.py --sk 2>&1 | tee gpurun_out/r2_19_sk_base.txt
echo "== new"; timeout 300 python scripts/gemm_bench.py --sk 2>&1 | tee gpurun_out/r2_19_sk_new.txt
