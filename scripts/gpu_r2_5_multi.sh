#!/bin/bash
# round 2, 2-GPU call: NCCL equivalence test of parallel.sample_sharded + the N=2 bench line (gather_ok)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2_5_multigpu_test.log; cat gpurun_out/r2_5_multigpu_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_5_bench_n2.json 2> gpurun_out/r2_5_bench_n2.err; tail -c 1800 gpurun_out/r2_5_bench_n2.json; tail -3 gpurun_out/r2_5_bench_n2.err
