#!/bin/bash
# tcgen05 conv engine: parity tests (isolated processes: a trapping kernel kills its CUDA context), micro-benchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
fail=0
for i in 0 1 2 3 4 5 6 7 8 9 10 11; do
  for cg in 2 1; do
    timeout 200 python -m pytest "tests/test_kernels_gpu.py::test_conv_tcgen05_bf16_forced_tile[case$i-$cg]" -q -x --timeout 150 -p no:cacheprovider > gpurun_out/tcf_${i}_$cg.log 2>&1
    rc=$?; if [ $rc -ne 0 ]; then fail=1; echo "case$i cg$cg rc=$rc"; grep -E "Error|error|assert|timeout|tag=" gpurun_out/tcf_${i}_$cg.log | head -5; fi
  done
done
echo "forced-tile tests fail=$fail"
timeout 300 python -m pytest tests/test_kernels_gpu.py -k "test_conv_tcgen05_bf16 and not forced" -q --timeout 200 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1; cat gpurun_out/gemm_bench.txt
