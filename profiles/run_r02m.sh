#!/bin/bash
# 1 GPU: fused backward staple kernel of the fermion force compiled for 2 / 3 / 4 CTAs per SM (254 / 168 / 128 registers)
tag=${1:-r02m}
mkdir -p gpurun_out
for mb in 2 3 4; do
  B200KS_LIB=$PWD/profiles/variants/libb200ks_fbwd$mb.so timeout 300 python bench.py --workload force --steps 4 --no-cpu-baseline \
      > gpurun_out/bench_force_${tag}_minb$mb.json 2> gpurun_out/bench_force_${tag}_minb$mb.err
  B200KS_LIB=$PWD/profiles/variants/libb200ks_fbwd$mb.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k "regex:StapleBwdSite" -c 12 \
      --log-file gpurun_out/launches_force_${tag}_minb$mb.csv python profiles/prof_force.py > gpurun_out/prof_force_${tag}_minb$mb.log 2>&1
done
for mb in 2 3 4; do python - <<P
import json, csv
d = json.load(open("gpurun_out/bench_force_${tag}_minb$mb.json"))
rows = [r for r in csv.reader(l for l in open("gpurun_out/launches_force_${tag}_minb$mb.csv") if l.startswith('"'))]
t = [float(r[-1]) / 1e3 for r in rows[1:]]
print("minb $mb seconds_per_call", d["seconds_per_call"], d["all_calls_s"], "StapleBwdSite us", sum(t) / max(len(t), 1), len(t))
P
done
