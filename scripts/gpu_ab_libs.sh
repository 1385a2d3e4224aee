#!/bin/bash
# same-box A/B of one graph-replayed denoise step between two libraries: scripts/gpu_ab_libs.sh <libA> <libB> <tag> [reps]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/$3.txt
: > $OUT
for i in 1 2 3; do
  for lib in $1 $2; do
    echo -n "$lib: " >> $OUT
    PD_B200_LIB=$PWD/$lib timeout 300 python scripts/profile_step.py --graph 1 --reps ${4:-20} 2>&1 | tail -1 >> $OUT
  done
done
cat $OUT
