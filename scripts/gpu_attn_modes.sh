#!/bin/bash
# attention_tc MMA issue modes: default build = FA_MMA_MODE 2 (interleaved), variant mma1 = group after group
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -3
echo "== FA_MMA_MODE=2 (default build)"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -E " 3 \|| 6 \|" | tee gpurun_out/attn_modes.txt
echo "== FA_MMA_MODE=1"; PD_B200_LIB=build/variants/libpd_mma1.so timeout 300 python scripts/attn_bench.py 2>&1 | grep -E " 3 \|" | tee -a gpurun_out/attn_modes.txt
