#!/bin/bash
# Run the GPU kernel tests in isolated processes (a trapping kernel kills its CUDA context, so
# each tcgen05 case gets its own interpreter). Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout $to python -m pytest "$@" -q -x --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"
}
run simt 600 tests/test_kernels_gpu.py -k "simt or group_norm or layer_norm or geglu or timestep or bridges or cfg_ddim"
run attn 600 tests/test_kernels_gpu.py -k "attention"
for i in 0 1 2 3 4 5 6 7 8 9 10 11; do
  run tc$i 300 "tests/test_kernels_gpu.py::test_conv_tcgen05_bf16[case$i]"
done
run tcrej 120 tests/test_kernels_gpu.py -k "rejects"
