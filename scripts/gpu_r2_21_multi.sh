#!/bin/bash
# 2-GPU call: NCCL equivalence test of parallel.sample_sharded (fp32: bit-identical to the single-GPU run)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2_21_multigpu_test.log; cat gpurun_out/r2_21_multigpu_test.log
