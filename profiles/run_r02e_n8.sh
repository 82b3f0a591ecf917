#!/bin/bash
# GPU-box script for the 8-GPU box (charged 8x: keep it short): the single-process context with 4 and 8 real devices,
# the reference's application on 4 devices behind the seam, one-rank-per-GPU parity on 8, the 8-GPU bench line
# (64^3x96 strong scaling with the one-GPU anchor, 96^3x192 point).
tag=${1:-r02e}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${tag}.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
timeout 200 python -m pytest tests/test_gpu_seam.py -q -m gpu -k "multi_gpu_context_matches_oracle and (8-dims3 or 4-dims2)" --durations=5 > gpurun_out/pytest_multi_${tag}.log 2>&1
echo "pytest rc=$? on $NG GPUs" >> gpurun_out/pytest_multi_${tag}.log
timeout 200 python -m pytest tests/test_dropin_apps.py -q -m gpu -k "several_gpus and 4-" --durations=5 > gpurun_out/pytest_apps_${tag}.log 2>&1
echo "pytest rc=$? on $NG GPUs" >> gpurun_out/pytest_apps_${tag}.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
    tests/mgpu_check.py --dims 8 8 12 24 > gpurun_out/mgpu_check_${tag}_n8.log 2>&1
echo "rc=$?" >> gpurun_out/mgpu_check_${tag}_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n8.json 2> gpurun_out/bench_${tag}_n8.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n8.err
tail -n 4 gpurun_out/pytest_multi_${tag}.log gpurun_out/pytest_apps_${tag}.log
tail -n 3 gpurun_out/mgpu_check_${tag}_n8.log; tail -c 300 gpurun_out/bench_${tag}_n8.err; head -c 400 gpurun_out/bench_${tag}_n8.json
