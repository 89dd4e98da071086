#!/bin/bash
# development: fused score kernel vs the two-kernel path at full size, same box, alternating
for i in 1 2; do
  timeout 300 python bench.py --no-gibbs --no-fbgmm --no-diffuse --no-cpu --steps 4 --warmup 3 > gpurun_out/fx_fused_$i.log 2>&1
  echo "fused $i"; python tools/bench_extract.py gpurun_out/fx_fused_$i.log value ms_per_step phases_ms roofline.kernel_ms roofline.frac e2e.ms_per_step
  timeout 300 python bench.py --two-kernel --no-gibbs --no-fbgmm --no-diffuse --no-cpu --steps 4 --warmup 3 > gpurun_out/fx_2k_$i.log 2>&1
  echo "two-kernel $i"; python tools/bench_extract.py gpurun_out/fx_2k_$i.log value ms_per_step phases_ms roofline.kernel_ms roofline.frac e2e.ms_per_step
done
