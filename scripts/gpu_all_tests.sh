#!/bin/bash
# Full GPU suite + smoke, logs in gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/kernels.log)"
timeout 2400 python -m pytest tests/test_model_gpu.py -q -s --timeout 900 -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "model rc=$? $(tail -1 gpurun_out/model.log)"
grep "\[parity\]" gpurun_out/model.log
timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_clip_gpu.py tests/test_pipeline_gpu.py -q -s --timeout 600 -p no:cacheprovider > gpurun_out/widen.log 2>&1
echo "vae+clip rc=$? $(tail -1 gpurun_out/widen.log)"; grep "\[parity\]" gpurun_out/widen.log
timeout 900 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
