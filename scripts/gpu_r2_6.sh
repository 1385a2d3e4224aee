#!/bin/bash
# round 2, GPU call 6: attention without the per-tile max pass -> tests, attention bench, step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | grep -v Warning | tail -8 > gpurun_out/r2_6_attn_tests.log; tail -6 gpurun_out/r2_6_attn_tests.log
timeout 300 python scripts/attn_bench.py > gpurun_out/r2_6_attn_bench.txt 2>&1; cat gpurun_out/r2_6_attn_bench.txt
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s 2>&1 | grep "\[parity\] apply_model\|\[parity\] 50\|passed\|failed" > gpurun_out/r2_6_parity.log; cat gpurun_out/r2_6_parity.log
timeout 300 python scripts/profile_step.py --graph 1 > gpurun_out/r2_6_step.log 2>&1; tail -1 gpurun_out/r2_6_step.log
