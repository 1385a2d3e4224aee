#!/bin/bash
# round 2, GPU call 2: new kernels (epilogue statistics, phase upsample, split convs) -> tests, parity log, bench, tune dump, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "statistics or epilogue or phase or split or producer" 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2_2_newtests.log
tail -25 gpurun_out/r2_2_newtests.log
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r2_2_tests_full.log
grep "\[parity\]\|\[property\]" gpurun_out/r2_2_tests_full.log > gpurun_out/r2_2_parity.log; cat gpurun_out/r2_2_parity.log
grep "passed\|failed\|^FAILED\|^ERROR\|Error:" gpurun_out/r2_2_tests_full.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_2_smoke.log 2>&1; tail -2 gpurun_out/r2_2_smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-config4 > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err; tail -c 2000 gpurun_out/r2_2_bench.json; tail -5 gpurun_out/r2_2_bench.err
PD_B200_AUTOTUNE=1 timeout 600 python scripts/make_tune_table.py --dump gpurun_out/r2_2_tune.inc > gpurun_out/r2_2_tune.log 2>&1; tail -3 gpurun_out/r2_2_tune.log
PD_DUMP=gpurun_out/r2_2_gemm_shapes.csv timeout 300 python scripts/profile_step.py --graph 1 > gpurun_out/r2_2_step.log 2>&1; tail -2 gpurun_out/r2_2_step.log
timeout 300 python scripts/profile_step.py > gpurun_out/r2_2_plain.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_2_launches.csv python scripts/profile_step.py > gpurun_out/r2_2_ncu.log 2>&1; tail -2 gpurun_out/r2_2_ncu.log
