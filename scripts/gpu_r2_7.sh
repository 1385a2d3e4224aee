#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s 2>&1 | grep "\[parity\] apply_model\|\[parity\] 50\|\[parity\] sampler\|passed\|failed" > gpurun_out/r2_7_parity.log; cat gpurun_out/r2_7_parity.log
timeout 300 python scripts/profile_step.py --graph 1 > gpurun_out/r2_7_step.log 2>&1; tail -1 gpurun_out/r2_7_step.log
timeout 300 python scripts/profile_step.py --graph 1 --batch 16 --size 768 > gpurun_out/r2_7_step_cfg4.log 2>&1; tail -1 gpurun_out/r2_7_step_cfg4.log
