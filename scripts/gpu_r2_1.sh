#!/bin/bash
# round 2, GPU call 1: tests + smoke + bench under both tile orders + tune dumps + per-shape GEMM dump
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2_1_tests.log
grep "\[parity\]\|\[property\]\|passed\|failed\|Error" gpurun_out/r2_1_tests.log | head -60
python -m pytest tests/test_model_gpu.py -m gpu -x -q -s 2>&1 | grep "\[parity\]\|\[property\]" > gpurun_out/r2_1_parity.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_1_smoke.log 2>&1; tail -3 gpurun_out/r2_1_smoke.log
python bench.py > gpurun_out/r2_1_bench_nfast1.json 2> gpurun_out/r2_1_bench_nfast1.err; tail -c 3000 gpurun_out/r2_1_bench_nfast1.json
PD_B200_NFAST=0 python bench.py --no-cpu-baseline --no-gpu-eager --no-config4 > gpurun_out/r2_1_bench_nfast0.json 2> gpurun_out/r2_1_bench_nfast0.err; tail -c 1500 gpurun_out/r2_1_bench_nfast0.json
PD_DUMP=gpurun_out/r2_1_gemm_shapes_nfast1.csv python scripts/profile_step.py --graph 1 > gpurun_out/r2_1_step_nfast1.log 2>&1; tail -2 gpurun_out/r2_1_step_nfast1.log
PD_B200_NFAST=0 PD_DUMP=gpurun_out/r2_1_gemm_shapes_nfast0.csv python scripts/profile_step.py --graph 1 > gpurun_out/r2_1_step_nfast0.log 2>&1; tail -2 gpurun_out/r2_1_step_nfast0.log
PD_B200_AUTOTUNE=1 python scripts/make_tune_table.py --dump gpurun_out/r2_1_tune_nfast1.inc > gpurun_out/r2_1_tune1.log 2>&1; tail -2 gpurun_out/r2_1_tune1.log
PD_B200_NFAST=0 PD_B200_AUTOTUNE=1 python scripts/make_tune_table.py --dump gpurun_out/r2_1_tune_nfast0.inc > gpurun_out/r2_1_tune0.log 2>&1; tail -2 gpurun_out/r2_1_tune0.log
