#!/bin/bash
# One profiling session (B200_PROFILING.md): plain run first, then the launch list, then --set full captures.
# Kernels are selected by MANGLED name (the filter GEMM has several template instances per run).
set -x
B="python bench.py --steps 2 --warmup 1 --no-gibbs --no-fbgmm --no-ingest --no-e2e --no-cpu --no-diffuse"
$B > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv $B > gpurun_out/ncu_l.log 2>&1
# sweep 2 of the headline (default precision: e4m3 first level): filter, refine, DP, collect
# (matched launches: collect(init) | filter refine dp collect | filter refine dp collect ...)
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"kmeans_filter_kernelILi5ELi1ELi0ELb1|refine_rows8_kernel|dp_staged_kernel|km_collect_kernel" -s 5 -c 4 \
    -o gpurun_out/r2_prof_kmeans -f $B > gpurun_out/ncu_k.log 2>&1
# the fp16 first-level filter (--precision fp16), sweep 2
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"kmeans_filter_kernelILi9ELi1ELi0ELb0" -s 1 -c 1 \
    -o gpurun_out/r2_prof_kmeans16 -f $B --precision fp16 > gpurun_out/ncu_k16.log 2>&1
# FBGMM log_marg_i: filter (same kernel, LSE threshold) + float64 refine, 4.19M rows x K = 5000
F="python tools/fvf_microbench.py --tokens-per-k 20 --reps 1"
$F > gpurun_out/ncu_plain_fv.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"kmeans_filter_kernelILi9ELi1ELi0ELb0|fv_refine_kernel" -s 2 -c 2 -o gpurun_out/r2_prof_fv -f $F > gpurun_out/ncu_f.log 2>&1
# the same with the e4m3 first level (isotropic variances)
F8="python tools/fvf_microbench.py --tokens-per-k 20 --reps 1 --fp8"
$F8 > gpurun_out/ncu_plain_fv8.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"kmeans_filter_kernelILi5ELi1ELi0ELb1|fv_refine_kernel" -s 2 -c 2 -o gpurun_out/r2_prof_fv8 -f $F8 > gpurun_out/ncu_f8.log 2>&1
# the diffuse model: second-level pass (gather, bitmap filter, bitmap refine) of a steady-state sweep
D="python bench.py --only-diffuse --no-cpu"
$D > gpurun_out/ncu_plain_dif.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"gather_undecided_kernel|kmeans_filter_kernelILi9ELi1ELi1ELb0|refine_bitmap_kernel" -s 43 -c 3 \
    -o gpurun_out/r2_prof_dif -f $D > gpurun_out/ncu_d.log 2>&1
# the fused score kernel (optional path), one launch
G="python tools/fvf_microbench.py --tokens-per-k 20 --reps 1 --fused --rows 2097152"
$G > gpurun_out/ncu_plain_fused.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"score_fused_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_fused -f $G > gpurun_out/ncu_g.log 2>&1
for r in kmeans kmeans16 fv fv8 dif fused; do
  ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_raw_$r.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
