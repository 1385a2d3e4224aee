#!/bin/bash
mkdir -p gpurun_out
{
echo "== full"; python scripts/attn_d160.py 2>&1 | tail -3
echo "== no Q loads"; PD_B200_LIB=build/variants/libpd_xsNOQ.so python scripts/attn_d160.py 2>&1 | tail -3
echo "== no stores"; PD_B200_LIB=build/variants/libpd_xsNOSTORE.so python scripts/attn_d160.py 2>&1 | tail -3
} | tee gpurun_out/r2_24_xattn_parts.txt
