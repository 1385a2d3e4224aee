#!/bin/bash
mkdir -p gpurun_out
{
echo "== 3 CTAs/SM (as built)"; python scripts/attn_d160.py 2>&1 | tail -2
echo "== launch bounds 4 CTAs/SM"; PD_B200_LIB=build/variants/libpd_xsb4.so python scripts/attn_d160.py 2>&1 | tail -2
echo "== launch bounds 5 CTAs/SM (spills)"; PD_B200_LIB=build/variants/libpd_xsb5.so python scripts/attn_d160.py 2>&1 | tail -2
} | tee gpurun_out/r2_25_xattn_occ.txt
