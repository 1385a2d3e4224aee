#!/bin/bash
# stream-K bring-up: conv parity tests (all tile variants), then the single-layer table with the stream-K columns
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv or geglu" 2>&1 | tail -15
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench_sk.txt 2>&1; cat gpurun_out/gemm_bench_sk.txt
