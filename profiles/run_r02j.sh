#!/bin/bash
# 1 GPU: ncu --set full of the 16-bit stencil, unpartitioned vs self-partitioned (same sites), to see where the
# partitioned kernel loses its 15 %.
tag=${1:-r02j}
mkdir -p gpurun_out
for mode in none zt; do
  if [ $mode = zt ]; then export B200KS_FORCE_PARTITION=zt; else unset B200KS_FORCE_PARTITION; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:dslash_half_kernel" -s 6 -c 2 -f \
      -o /tmp/prof_half_${mode}_${tag} python profiles/prof_half_part.py > gpurun_out/prof_half_${mode}_${tag}.log 2>&1
  ncu -i /tmp/prof_half_${mode}_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_half_${mode}_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_half_${mode}_${tag}.ncu-rep --page source --csv > gpurun_out/prof_half_${mode}_${tag}_source.csv 2>/dev/null
done
ls -la gpurun_out/prof_half_*_${tag}*; tail -2 gpurun_out/prof_half_zt_${tag}.log
