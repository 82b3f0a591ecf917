#!/bin/bash
# 8-GPU box (charged 8x): one-rank-per-GPU parity incl. the mixed solvers, then the 8-GPU bench line.
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
    tests/mgpu_check.py --dims 8 8 12 24 > gpurun_out/mgpu_check_${tag}_n8.log 2>&1
echo "rc=$?" >> gpurun_out/mgpu_check_${tag}_n8.log
timeout 200 python -m pytest tests/test_gpu_seam.py -q -m gpu -k "multi_gpu_context_matches_oracle and 8-dims3" > gpurun_out/pytest_multi_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi_${tag}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n8.json 2> gpurun_out/bench_${tag}_n8.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n8.err
grep -E "FAIL|MGPU|mixed" gpurun_out/mgpu_check_${tag}_n8.log; tail -n 2 gpurun_out/pytest_multi_${tag}.log
tail -c 200 gpurun_out/bench_${tag}_n8.err; python - <<P
import json
d = json.loads([l for l in open("gpurun_out/bench_${tag}_n8.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["efficiency_same_lattice"], d["roofline"]["16bit"], d["e2e"]["ms_per_step"], d["weak_point"]["seconds"] if d.get("weak_point") else None)
P
