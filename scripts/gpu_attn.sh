#!/bin/bash
# tcgen05 attention: parity tests, micro-benchmark, per-tile timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -k "attention" -q -x --timeout 300 -p no:cacheprovider > gpurun_out/attn.log 2>&1
echo "attn rc=$? $(tail -1 gpurun_out/attn.log)"; grep -E "Error|error|assert|timeout" gpurun_out/attn.log | head -10
timeout 300 python scripts/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 120 python scripts/attn_timeline.py > gpurun_out/attn_timeline.log 2>&1; cat gpurun_out/attn_timeline.log
