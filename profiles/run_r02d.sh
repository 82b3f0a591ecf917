#!/bin/bash
# GPU-box script, round 2 third pass (1 GPU): the whole GPU suite as the driver runs it (+ durations), smoke, the
# default bench line and the reference arm.
tag=${1:-r02d}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${tag}.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --durations=25 > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_${tag}.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_n1_reference.json 2> gpurun_out/bench_${tag}_n1_reference.err
timeout 900 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n1.err
grep -v "^\s*$" gpurun_out/pytest_${tag}.log | tail -n 45
cat gpurun_out/smoke_${tag}.log; tail -c 400 gpurun_out/bench_${tag}_n1.err; head -c 600 gpurun_out/bench_${tag}_n1.json
