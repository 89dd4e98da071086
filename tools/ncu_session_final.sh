#!/bin/bash
# Profiling session for the FINAL round-2 kernels (packed top-3 keys in the e4m3 filter, refine kernels specialised
# for D = 130): plain run first, then the launch list, then --set full captures of the headline sweep's kernels and of
# the FBGMM log_marg_i pair with the e4m3 first level.  Same recipe as tools/ncu_session.sh (B200_PROFILING.md).
set -x
B="python bench.py --steps 2 --warmup 1 --no-gibbs --no-fbgmm --no-ingest --no-e2e --no-cpu --no-diffuse"
$B > gpurun_out/ncu_plain_final.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final_bench_steps2.csv $B > gpurun_out/ncu_l_final.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"kmeans_filter_kernelILi5ELi1ELi0ELb1|refine_rows8_kernel|dp_staged_kernel|km_collect_kernel" -s 5 -c 4 \
    -o gpurun_out/r2_prof_kmeans_final -f $B > gpurun_out/ncu_k_final.log 2>&1
F8="python tools/fvf_microbench.py --tokens-per-k 20 --reps 1 --fp8"
$F8 > gpurun_out/ncu_plain_fv8_final.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"kmeans_filter_kernelILi5ELi1ELi0ELb1|fv_refine_kernel" -s 2 -c 2 -o gpurun_out/r2_prof_fv8_final -f $F8 > gpurun_out/ncu_f8_final.log 2>&1
for r in kmeans_final fv8_final; do
  ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_raw_$r.csv 2>/dev/null
done
ls -la gpurun_out/*final*.ncu-rep
