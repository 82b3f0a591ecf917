#!/bin/bash
# 1 GPU, the last 90 s of round 2: duration and DRAM traffic of every backward staple launch of the fermion force in
# three forms, one 3-term call each (168 launches per call): fused body / plain order, fused body / parities interleaved,
# two-role kernel / interleaved.  (ncu's -k matches the function name without template arguments.)
tag=${1:-r02v}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 72 ncu --metrics $M --clock-control none -k regex:"force_site_kernel2|force_pair_kernel" -c 504 --csv \
    --log-file gpurun_out/ncu_fbwd_${tag}.csv python profiles/force_ab.py --only 0,0,0:1,0,0:1,2,0 > gpurun_out/ncu_fbwd_${tag}.log 2>&1
echo "ncu rc=$?"; tail -n 4 gpurun_out/ncu_fbwd_${tag}.log; wc -l gpurun_out/ncu_fbwd_${tag}.csv
