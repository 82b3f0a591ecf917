#!/bin/bash
# GPU-box script (1 GPU): relative residual in the mixed solvers, host upload probe, reliable-update trace.
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seam.py tests/test_gpu_parity.py tests/test_gpu_block.py -q -m gpu -k "not multi_gpu and not several" --durations=5 > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
timeout 300 python profiles/upload_probe.py > gpurun_out/upload_probe_${tag}.json 2> gpurun_out/upload_probe_${tag}.err
B200KS_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_trace_${tag}.json 2> gpurun_out/bench_trace_${tag}.err
grep -v "^\s*$" gpurun_out/pytest_${tag}.log | tail -n 12
cat gpurun_out/upload_probe_${tag}.json; tail -3 gpurun_out/upload_probe_${tag}.err
grep "b200ks mixed cg" gpurun_out/bench_trace_${tag}.err | tail -n 40
