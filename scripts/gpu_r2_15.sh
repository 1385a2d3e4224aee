#!/bin/bash
mkdir -p gpurun_out
PD_B200_AUTOTUNE=1 PD_B200_RETUNE=1 timeout 1200 python scripts/make_tune_table.py --dump gpurun_out/r2_15_tune_dump.inc > gpurun_out/r2_15_tune.log 2>&1; tail -3 gpurun_out/r2_15_tune.log
wc -l gpurun_out/r2_15_tune_dump.inc
