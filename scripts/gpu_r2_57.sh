#!/bin/bash
# Round-end measurement set of the final state (prefix r2i_): GPU suite as the driver runs it, smoke, bench line, graph step,
# ncu launch list, one ncu --set full capture of the self-attention kernel, attention / GEMM micro-benchmarks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu -s -p no:cacheprovider > gpurun_out/r2i_gpu_tests.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r2i_gpu_tests.log)"
grep "\[parity\]" gpurun_out/r2i_gpu_tests.log > gpurun_out/r2i_parity.txt; tail -1 gpurun_out/r2i_gpu_tests.log >> gpurun_out/r2i_parity.txt
timeout 900 python __graft_entry__.py smoke > gpurun_out/r2i_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r2i_smoke.log)"
timeout 1500 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2i_bench.json
timeout 600 python scripts/profile_step.py --reps 20 --graph 1 > gpurun_out/r2i_step_graph.log 2>&1; tail -1 gpurun_out/r2i_step_graph.log
timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/r2i_step.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2i_launches_step.csv python scripts/profile_step.py > gpurun_out/r2i_ncu_step.log 2>&1
echo "launch list rc=$?"
python scripts/attn_one.py > gpurun_out/r2i_attn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/r2i_prof_attn \
    python scripts/attn_one.py > gpurun_out/r2i_ncu_attn.log 2>&1
echo "attention full rc=$?"
python scripts/attn_bench.py > gpurun_out/r2i_attn_bench.txt 2>&1; tail -12 gpurun_out/r2i_attn_bench.txt
python scripts/attn_timeline.py > gpurun_out/r2i_attn_timeline.txt 2>&1
