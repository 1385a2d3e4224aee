#!/bin/bash
# A/B of attention_tc build variants (build/variants/libpd_*.so) on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/attn_variants.txt
for lib in build/variants/libpd_*.so; do
  PD_B200_LIB=$PWD/$lib timeout 120 python scripts/attn_quick.py >> gpurun_out/attn_variants.txt 2>&1 || echo "$lib FAILED" >> gpurun_out/attn_variants.txt
done
cat gpurun_out/attn_variants.txt
for t in ${TIMELINE_TAGS:-o0p0}; do
  echo "== timeline $t"; PD_B200_LIB=$PWD/build/variants/libpd_$t.so timeout 120 python scripts/attn_timeline.py 2>&1 | head -12 | tee gpurun_out/attn_timeline_$t.txt
done
