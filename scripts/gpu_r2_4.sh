#!/bin/bash
# round 2, GPU call 4: validation of the settled configuration (tune table installed, statistics hand-off off by default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r2_4_tests_full.log
grep "\[parity\]\|\[property\]" gpurun_out/r2_4_tests_full.log > gpurun_out/r2_4_parity.log; cat gpurun_out/r2_4_parity.log | grep -v "vae\|clip\|tokens"
grep "passed\|failed\|^FAILED\|^ERROR\|Error:" gpurun_out/r2_4_tests_full.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_4_smoke.log 2>&1; tail -2 gpurun_out/r2_4_smoke.log
timeout 900 python bench.py > gpurun_out/r2_4_bench.json 2> gpurun_out/r2_4_bench.err; tail -c 3500 gpurun_out/r2_4_bench.json; tail -5 gpurun_out/r2_4_bench.err
PD_DUMP=gpurun_out/r2_4_gemm_shapes.csv timeout 300 python scripts/profile_step.py --graph 1 > gpurun_out/r2_4_step.log 2>&1; tail -2 gpurun_out/r2_4_step.log
timeout 300 python scripts/profile_step.py > gpurun_out/r2_4_plain.log 2>&1 && timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_4_launches.csv python scripts/profile_step.py > gpurun_out/r2_4_ncu.log 2>&1; tail -2 gpurun_out/r2_4_ncu.log
