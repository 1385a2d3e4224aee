#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_31_tests.log
for i in 1 2; do
timeout 300 python scripts/profile_step.py --graph 1 --reps 20 2>&1 | tail -1
PD_TIME_TABLE=0 timeout 300 python scripts/profile_step.py --graph 1 --reps 20 2>&1 | tail -1
done | tee gpurun_out/r2_31_time_table_ab.txt
