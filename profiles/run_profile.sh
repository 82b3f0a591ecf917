#!/bin/bash
# Runs on the GPU box (under gpurun): launch lists of the bench command and of the block solver,
# and one --set full capture of the stencil variants (report kept in /tmp, only its raw CSV page
# comes back: gpurun_out/ is limited to 64 MiB).  Usage: profiles/run_profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 500 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
    --log-file gpurun_out/launches_block_${tag}.csv python profiles/prof_block.py \
    > gpurun_out/prof_block_${tag}.log 2>&1
ncu --set full --clock-control none -k "regex:dslash_half_kernel|dslash_mrhs_kernel|dslash_kernel<float|dslash_kernel<double, 0" -c 14 -f \
    -o /tmp/prof_dslash_${tag} python profiles/prof_target2.py > gpurun_out/prof_target_${tag}.log 2>&1
ncu -i /tmp/prof_dslash_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_dslash_${tag}_raw.csv 2>/dev/null
tail -4 gpurun_out/prof_target_${tag}.log
