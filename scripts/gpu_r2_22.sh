#!/bin/bash
# three-group attention with a polling back-off in its MMA warp: kernel alone and inside the step (same box)
mkdir -p gpurun_out
OUT=gpurun_out/r2_22_tc3_sleep.txt; : > $OUT
for n in 0 50 150; do
  echo "== F3_POLL_SLEEP=$n" >> $OUT
  PD_B200_LIB=build/variants/libpd_tc3s$n.so timeout 400 python scripts/attn_ab.py 2>&1 | tail -4 >> $OUT
done
cat $OUT
