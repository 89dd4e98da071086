#!/bin/bash
# One profiling session (B200_PROFILING.md): plain run first, then the launch list, then --set full captures.
set -x
B="python bench.py --steps 2 --warmup 1 --no-gibbs --no-fbgmm --no-ingest --no-e2e --no-cpu"
$B > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv $B > gpurun_out/ncu_l.log 2>&1
# sweep 2 of the headline: filter, refine, DP, collect (launch order: collect(init) | filter refine dp collect | ...)
ncu --set full --clock-control none --import-source on -k regex:"kmeans_filter_kernel|refine_rows8_kernel|dp_staged_kernel|km_collect_kernel" -s 5 -c 4 -o gpurun_out/r2_prof_kmeans -f $B > gpurun_out/ncu_k.log 2>&1
# FBGMM log_marg_i: filter (same kernel, LSE threshold) + float64 refine, 4.19M rows x K = 5000
F="python tools/fvf_microbench.py --tokens-per-k 20 --reps 1"
$F > gpurun_out/ncu_plain_fv.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"kmeans_filter_kernel|fv_refine_kernel" -s 2 -c 2 -o gpurun_out/r2_prof_fv -f $F > gpurun_out/ncu_f.log 2>&1
# the fused score kernel (optional path), one launch
G="python tools/fvf_microbench.py --tokens-per-k 20 --reps 1 --fused --rows 2097152"
$G > gpurun_out/ncu_plain_fused.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"score_fused_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_fused -f $G > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out/*.ncu-rep
