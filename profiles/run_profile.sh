#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one --set full capture
# of the stencil kernels.  Usage: profiles/run_profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 400 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dslash_ -c 10 -f \
    -o gpurun_out/prof_dslash_${tag} python profiles/prof_target.py > gpurun_out/prof_target_${tag}.log 2>&1
tail -5 gpurun_out/prof_target_${tag}.log
