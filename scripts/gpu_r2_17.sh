#!/bin/bash
mkdir -p gpurun_out
AB_QUICK=1 timeout 600 python scripts/stats_ab.py > gpurun_out/r2_17_stats_ab.txt 2>&1; cat gpurun_out/r2_17_stats_ab.txt
