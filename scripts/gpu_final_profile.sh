#!/bin/bash
# Round-end measurement set: bench line, ncu launch list of one denoising step (cold and warm caches), DRAM traffic of the
# conv engine over the step, and one `ncu --set full` capture of each hot kernel.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_final.json
timeout 600 python scripts/profile_step.py --reps 3 --graph 1 > gpurun_out/step_graph.log 2>&1; tail -1 gpurun_out/step_graph.log
timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/step.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_final.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "launch list rc=$?"
timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/step.log 2>&1 && \
ncu --profile-from-start off -k regex:conv_tc --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/conv_tc_dram.csv python scripts/profile_step.py > gpurun_out/ncu_step2.log 2>&1
echo "conv dram rc=$?"
python scripts/gemm_bench.py --one 16 64 64 320 320 3 0 --iters 3 > gpurun_out/one_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/prof_conv_tc \
    python scripts/gemm_bench.py --one 16 64 64 320 320 3 0 --iters 3 > gpurun_out/ncu_gemm.log 2>&1
echo "conv_tc full rc=$?"; cat gpurun_out/one_gemm.log
python scripts/attn_one.py > gpurun_out/attn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn \
    python scripts/attn_one.py > gpurun_out/ncu_attn.log 2>&1
echo "attention full rc=$?"
python scripts/gn_one.py > gpurun_out/gn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gn_fused -s 2 -c 1 -f -o gpurun_out/prof_gn \
    python scripts/gn_one.py > gpurun_out/ncu_gn.log 2>&1
echo "gn full rc=$?"
python scripts/attn_one.py > /dev/null 2>&1; python scripts/xattn_one.py > gpurun_out/xattn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_short -s 2 -c 1 -f -o gpurun_out/prof_xattn \
    python scripts/xattn_one.py > gpurun_out/ncu_xattn.log 2>&1
echo "short attention full rc=$?"
CROSS=1 python scripts/attn_quick.py > gpurun_out/attn_quick.log 2>&1
python scripts/attn_bench.py > gpurun_out/attn_bench.log 2>&1; python scripts/norm_bench.py > gpurun_out/norm_bench.log 2>&1
python scripts/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1
