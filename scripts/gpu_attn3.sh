#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "three_groups or outgrow" 2>&1 | tail -3
echo "== stagger 1200 (default)"; timeout 300 python scripts/attn_bench.py 2>&1 | grep -E " 3 \|| 6 \|" | grep " 40 " | tee gpurun_out/attn3_stagger.txt
for v in stag0 stag2400; do echo "== $v"; PD_B200_LIB=build/variants/libpd_$v.so timeout 300 python scripts/attn_bench.py 2>&1 | grep -E " 6 \|" | tee -a gpurun_out/attn3_stagger.txt; done
python scripts/attn4_timeline.py 6 2>&1 | head -12 | tee gpurun_out/attn3_timeline_b.txt
