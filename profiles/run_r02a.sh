#!/bin/bash
# GPU-box script, round 2 first pass: host/PCIe probe, parity + seam tests, default bench line.
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${tag}.txt 2>&1
free -g >> gpurun_out/gpus_${tag}.txt 2>&1; nproc >> gpurun_out/gpus_${tag}.txt
profiles/bin/hostbw_probe > gpurun_out/hostbw_${tag}.json 2> gpurun_out/hostbw_${tag}.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_seam.py -x -q -m gpu > gpurun_out/pytest_a_${tag}.log 2>&1
echo "pytest_a rc=$?" >> gpurun_out/pytest_a_${tag}.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n1.err
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_seam.py > gpurun_out/pytest_b_${tag}.log 2>&1
echo "pytest_b rc=$?" >> gpurun_out/pytest_b_${tag}.log
tail -3 gpurun_out/pytest_a_${tag}.log gpurun_out/pytest_b_${tag}.log; tail -c 600 gpurun_out/bench_${tag}_n1.err
