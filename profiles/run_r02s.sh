#!/bin/bash
# GPU-box script, closing 1-GPU pass of round 2: the whole GPU suite as the driver runs it, smoke, reference arm, default bench line.
tag=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_${tag}.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_n1_reference.json 2> gpurun_out/bench_${tag}_n1_reference.err
timeout 900 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n1.err
grep -v "^\s*$" gpurun_out/pytest_${tag}.log | tail -n 14
cat gpurun_out/smoke_${tag}.log; tail -c 200 gpurun_out/bench_${tag}_n1.err; head -c 400 gpurun_out/bench_${tag}_n1.json
