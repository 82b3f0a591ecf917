#!/bin/bash
# 1 GPU: K-wide stencil and block CG on a partitioned context (the GPU as its own neighbour), plus the unpartitioned block tests
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_block.py tests/test_gpu_sequences.py -q -m gpu --durations=6 -x > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
grep -v "^\s*$" gpurun_out/pytest_${tag}.log | tail -n 40
