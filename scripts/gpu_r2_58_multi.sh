#!/bin/bash
# 2-GPU call: N = 2 bench line of the final state (gather_ok) + the NCCL equivalence test
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; tail -c 600 gpurun_out/r2i_bench_n2.json; tail -3 gpurun_out/r2i_bench_n2.err
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -6
