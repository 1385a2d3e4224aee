#!/bin/bash
# same-box A/B of one graph-replayed denoise step: build/variants/libpd_base.so (A) vs the in-tree library (B), alternating
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/${1:-ab_step}.txt
: > $OUT
for i in 1 2 3; do
  for lib in build/variants/libpd_base.so prompt-diffusion_b200/libpd_b200.so; do
    echo -n "$lib: " >> $OUT
    PD_B200_LIB=$lib timeout 300 python scripts/profile_step.py --graph 1 --reps 20 2>&1 | tail -1 >> $OUT
  done
done
cat $OUT
