#!/bin/bash
# ncu --set full of the narrow short-K GEMM (65536 x 320 x 320, residual): tabled variant and resident-B
mkdir -p gpurun_out
python scripts/gemm_bench.py --one 16 64 64 320 320 1 1 --iters 3 --blocked > gpurun_out/r2_11_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2_11_narrow \
    python scripts/gemm_bench.py --one 16 64 64 320 320 1 1 --iters 3 --blocked > gpurun_out/r2_11_ncu_a.log 2>&1
echo "narrow rc=$?"
PD_BENCH_BRES=1,160 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2_11_narrow_bres \
    python scripts/gemm_bench.py --one 16 64 64 320 320 1 1 --iters 3 --blocked > gpurun_out/r2_11_ncu_b.log 2>&1
echo "narrow bres rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2_11_narrow_nores \
    python scripts/gemm_bench.py --one 16 64 64 320 320 1 0 --iters 3 --blocked > gpurun_out/r2_11_ncu_c.log 2>&1
echo "narrow nores rc=$?"
ls -la gpurun_out/*.ncu-rep
