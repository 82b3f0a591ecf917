#!/bin/bash
# 2-GPU box: partitioned block solve on real peers -- parity (mgpu_check, single-process context), then timing
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    tests/mgpu_check.py --dims 8 8 12 24 > gpurun_out/mgpu_check_${tag}_n2.log 2>&1
echo "rc=$?" >> gpurun_out/mgpu_check_${tag}_n2.log
timeout 200 python -m pytest tests/test_gpu_seam.py -q -m gpu -k "multi_gpu_context_matches_oracle and (2-dims0 or 2-dims4)" > gpurun_out/pytest_multi_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_${tag}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    profiles/block_multi_probe.py > gpurun_out/block_multi_${tag}_n2.json 2> gpurun_out/block_multi_${tag}_n2.err
grep -E "FAIL|MGPU|block" gpurun_out/mgpu_check_${tag}_n2.log; tail -n 2 gpurun_out/pytest_multi_${tag}.log
grep "^{" gpurun_out/block_multi_${tag}_n2.json; tail -n 3 gpurun_out/block_multi_${tag}_n2.err
