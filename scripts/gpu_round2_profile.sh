#!/bin/bash
# Round-2 measurement set: bench line, ncu launch list of one denoising step, DRAM traffic of the conv engine over the
# step, and one `ncu --set full` capture of each hot kernel.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# gpurun copies back at most 64 MiB: every .ncu-rep (~20 MB) is exported to text (details / raw csv / source csv) and removed
export_rep() {
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2f_bench.json
PD_DUMP=gpurun_out/r2f_gemm_shapes.csv timeout 600 python scripts/profile_step.py --reps 3 --graph 1 > gpurun_out/r2f_step_graph.log 2>&1; tail -2 gpurun_out/r2f_step_graph.log
timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/r2f_step.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2f_launches.csv python scripts/profile_step.py > gpurun_out/r2f_ncu_step.log 2>&1
echo "launch list rc=$?"
timeout 600 python scripts/profile_step.py --reps 1 > gpurun_out/r2f_step.log 2>&1 && \
ncu --profile-from-start off -k regex:conv_tc --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2f_conv_tc_dram.csv python scripts/profile_step.py > gpurun_out/r2f_ncu_step2.log 2>&1
echo "conv dram rc=$?"
python scripts/gemm_bench.py --one 16 64 64 320 320 3 0 --iters 3 > gpurun_out/r2f_one_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2f_prof_conv_tc \
    python scripts/gemm_bench.py --one 16 64 64 320 320 3 0 --iters 3 > gpurun_out/r2f_ncu_gemm.log 2>&1
echo "conv_tc 3x3 full rc=$?"; cat gpurun_out/r2f_one_gemm.log; export_rep r2f_prof_conv_tc
python scripts/gemm_bench.py --one 16 64 64 320 320 1 1 --iters 3 > gpurun_out/r2f_one_gemm_narrow.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/r2f_prof_conv_tc_narrow \
    python scripts/gemm_bench.py --one 16 64 64 320 320 1 1 --iters 3 > gpurun_out/r2f_ncu_gemm_narrow.log 2>&1
echo "conv_tc 1x1 full rc=$?"; cat gpurun_out/r2f_one_gemm_narrow.log; export_rep r2f_prof_conv_tc_narrow
python scripts/attn_one.py > gpurun_out/r2f_attn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/r2f_prof_attn \
    python scripts/attn_one.py > gpurun_out/r2f_ncu_attn.log 2>&1
echo "attention full rc=$?"; export_rep r2f_prof_attn
python scripts/attn_one.py 40 4096 16 6 > gpurun_out/r2f_attn3_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 2 -c 1 -f -o gpurun_out/r2f_prof_attn3 \
    python scripts/attn_one.py 40 4096 16 6 > gpurun_out/r2f_ncu_attn3.log 2>&1
echo "attention (three groups) full rc=$?"; export_rep r2f_prof_attn3
python scripts/gn_one.py > gpurun_out/r2f_gn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gn_fused -s 2 -c 1 -f -o gpurun_out/r2f_prof_gn \
    python scripts/gn_one.py > gpurun_out/r2f_ncu_gn.log 2>&1
echo "gn full rc=$?"; export_rep r2f_prof_gn
python scripts/gemm_bench.py > gpurun_out/r2f_gemm_bench.txt 2>&1; tail -20 gpurun_out/r2f_gemm_bench.txt
