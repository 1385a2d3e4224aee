#!/bin/bash
mkdir -p gpurun_out
{
echo "== full"; python scripts/attn_d160.py 2>&1 | tail -5
echo "== no stores"; PD_B200_LIB=build/variants/libpd_xtNOSTORE.so python scripts/attn_d160.py 2>&1 | tail -5
echo "== no Q loads after the first slots"; PD_B200_LIB=build/variants/libpd_xtNOQLOAD.so python scripts/attn_d160.py 2>&1 | tail -5
} | tee gpurun_out/r2_28_xtc_parts.txt
