#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py > gpurun_out/insitu_plain.log 2>&1 && timeout 600 ncu --profile-from-start off -k regex:attention_tc --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/insitu_attn_tc3.csv python scripts/profile_step.py > /dev/null 2>&1
PD_B200_ATTN3=0 timeout 600 ncu --profile-from-start off -k regex:attention_tc --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/insitu_attn_tc.csv python scripts/profile_step.py > /dev/null 2>&1
grep -h "attention_tc" gpurun_out/insitu_attn_tc3.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | head -16
echo ==; grep -h "attention_tc" gpurun_out/insitu_attn_tc.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | head -16
