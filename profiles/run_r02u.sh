#!/bin/bash
# 1 GPU, last call of round 2 (~6 GPU-minutes left): A/B of the launch variants of the link construction and the fermion
# force (profiles/force_ab.py), the parity tests of both under the default and the candidate switches, then -- time
# permitting -- DRAM traffic and duration of the backward staple kernel in the baseline and the candidate form.
tag=${1:-r02u}
mkdir -p gpurun_out
timeout 150 python profiles/force_ab.py > gpurun_out/force_ab_${tag}.json 2> gpurun_out/force_ab_${tag}.err
echo "force_ab rc=$?"
FILES="tests/test_gpu_links.py tests/test_gpu_zy_force.py tests/test_dropin_apps.py tests/test_config0_l6666.py"
timeout 200 python -m pytest $FILES -q -m gpu -x > gpurun_out/pytest_${tag}_default.log 2>&1
echo "pytest default rc=$?"; tail -n 3 gpurun_out/pytest_${tag}_default.log
B200KS_SITE_ORDER=1 B200KS_FORCE_SPLIT=2 B200KS_FORCE_OVERLAP=1 timeout 200 python -m pytest $FILES -q -m gpu -x > gpurun_out/pytest_${tag}_candidate.log 2>&1
echo "pytest candidate rc=$?"; tail -n 3 gpurun_out/pytest_${tag}_candidate.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
# (the target runs the link chain twice = 264 staple_kernel launches, then one 3-term force: the window holds the last
# staple passes of the link construction and the first forward / backward staple passes of the force)
for v in 0,0,0 1,2,0; do
  timeout 100 ncu --metrics $M --clock-control none -k regex:"StapleBwd|StapleFwd|staple_kernel" -s 250 -c 60 --csv \
      --log-file gpurun_out/ncu_force_${tag}_${v//,/}.csv python profiles/force_ab.py --only $v > gpurun_out/ncu_force_${tag}_${v//,/}.log 2>&1
  echo "ncu $v rc=$?"
done
cat gpurun_out/force_ab_${tag}.json | head -c 3000
