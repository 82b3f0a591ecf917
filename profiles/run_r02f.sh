#!/bin/bash
# GPU-box script, round 2 fourth pass (1 GPU): whole GPU suite (meson tie-ups and programmatic dependent launch are new),
# A/B of B200KS_PDL, meson workload, partitioned stencil on one GPU as its own neighbour, launch lists.
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
for pdl in 0 1 0 1; do
  B200KS_PDL=$pdl timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline \
      >> gpurun_out/bench_${tag}_pdl${pdl}.json 2>> gpurun_out/bench_${tag}_pdl${pdl}.err
done
timeout 300 python bench.py --workload meson --nmom 20 > gpurun_out/bench_meson_${tag}.json 2> gpurun_out/bench_meson_${tag}.err
timeout 300 python bench.py --workload meson --nmom 1 --no-cpu-baseline >> gpurun_out/bench_meson_${tag}.json 2>> gpurun_out/bench_meson_${tag}.err
timeout 300 python bench.py --workload meson --nmom 100 --no-cpu-baseline >> gpurun_out/bench_meson_${tag}.json 2>> gpurun_out/bench_meson_${tag}.err
B200KS_PDL=0 timeout 300 python profiles/halo_probe.py > gpurun_out/halo_probe_${tag}.json 2> gpurun_out/halo_probe_${tag}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k "regex:meson" \
    --log-file gpurun_out/launches_meson_${tag}.csv python bench.py --workload meson --nmom 20 --no-cpu-baseline --steps 5 > gpurun_out/prof_meson_${tag}.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n1.err
grep -v "^\s*$" gpurun_out/pytest_${tag}.log | tail -n 24
for pdl in 0 1; do python - <<P
import json
for ln in open("gpurun_out/bench_${tag}_pdl${pdl}.json"):
    d = json.loads(ln); print("pdl ${pdl}", d["ms_per_step"], d["roofline"]["achieved"], [m["cg_time_to_solution_s"] for m in d["other_precision_modes"]])
P
done
cut -c1-330 gpurun_out/bench_meson_${tag}.json; tail -c 300 gpurun_out/bench_${tag}_n1.err
