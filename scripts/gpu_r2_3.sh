#!/bin/bash
# round 2, GPU call 3: four-group attention kernel + GEGLU fit -> tests, attention bench, in-process A/B of the statistics switches, tune dump
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention or geglu or statistics or producer" 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2_3_newtests.log
tail -12 gpurun_out/r2_3_newtests.log
timeout 300 python scripts/attn_bench.py > gpurun_out/r2_3_attn_bench.txt 2>&1; cat gpurun_out/r2_3_attn_bench.txt
timeout 900 python scripts/stats_ab.py > gpurun_out/r2_3_stats_ab.txt 2>&1; cat gpurun_out/r2_3_stats_ab.txt
PD_B200_AUTOTUNE=1 timeout 900 python scripts/make_tune_table.py --dump gpurun_out/r2_3_tune.inc > gpurun_out/r2_3_tune.log 2>&1; tail -3 gpurun_out/r2_3_tune.log
