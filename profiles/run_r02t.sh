#!/bin/bash
# 1 GPU: --set full capture of the stencil variants with the final code (roofline.traffic of bench.py comes from it)
tag=${1:-r02t}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:dslash_half_kernel|dslash_kernel<float|dslash_kernel<double, 0" -c 8 -f \
    -o /tmp/prof_dslash_${tag} python profiles/prof_target2.py > gpurun_out/prof_target_${tag}.log 2>&1
ncu -i /tmp/prof_dslash_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_dslash_${tag}_raw.csv 2>/dev/null
tail -3 gpurun_out/prof_target_${tag}.log; ls -la gpurun_out/prof_dslash_${tag}_raw.csv
