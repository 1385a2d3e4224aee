#!/bin/bash
# config 4 (768^2, batch 16, 9216-token self-attention): three-group kernel (auto from 6144 queries) vs the ping-pong engine 3
mkdir -p gpurun_out
OUT=gpurun_out/r2g_ab_config4_attn.txt; : > $OUT
python scripts/attn_bench.py 2>&1 | grep -E " (3|6) \|" | tee -a $OUT
for i in 1 2; do
  for v in 1 0; do
    echo -n "PD_B200_ATTN3=$v: " >> $OUT
    PD_B200_ATTN3=$v timeout 300 python scripts/profile_step.py --graph 1 --reps 8 --batch 16 --size 768 2>&1 | tail -1 >> $OUT
  done
done
cat $OUT
