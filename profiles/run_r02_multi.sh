#!/bin/bash
# GPU-box script for an N-GPU box: multi-GPU parity (one rank per GPU and one process for all GPUs), the
# reference's applications with B200KS_NGPU devices behind the seam, and the bench lines.
#   profiles/run_r02_multi.sh <tag> <N for bench> [more N ...]
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${tag}.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus_${tag}.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_seam.py tests/test_multigpu.py -q -m gpu -k "multi" > gpurun_out/pytest_multi_${tag}.log 2>&1
echo "pytest rc=$? on $NG GPUs" >> gpurun_out/pytest_multi_${tag}.log
timeout 900 python -m pytest tests/test_dropin_apps.py tests/test_config0_l6666.py -q -m gpu -k "several_gpus" > gpurun_out/pytest_apps_${tag}.log 2>&1
echo "pytest rc=$? on $NG GPUs" >> gpurun_out/pytest_apps_${tag}.log
for n in "$@"; do
  port=$((29500 + n))
  case $n in 2) dims="8 8 12 24";; 4) dims="8 8 8 24";; 8) dims="8 8 12 24";; esac
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      tests/mgpu_check.py --dims $dims > gpurun_out/mgpu_check_${tag}_n${n}.log 2>&1
  echo "rc=$?" >> gpurun_out/mgpu_check_${tag}_n${n}.log
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((port + 50)) \
      bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n${n}.json 2> gpurun_out/bench_${tag}_n${n}.err
  echo "bench rc=$?" >> gpurun_out/bench_${tag}_n${n}.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((port + 70)) \
      bench.py --gpus $n --steps 1 --warmup 0 --impl reference > gpurun_out/bench_${tag}_n${n}_reference.json 2> gpurun_out/bench_${tag}_n${n}_reference.err
done
tail -n 3 gpurun_out/pytest_multi_${tag}.log gpurun_out/pytest_apps_${tag}.log
for n in "$@"; do tail -n 2 gpurun_out/mgpu_check_${tag}_n${n}.log; tail -c 300 gpurun_out/bench_${tag}_n${n}.err; done
