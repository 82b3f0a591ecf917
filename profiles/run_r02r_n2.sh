#!/bin/bash
# 2-GPU box: the driver's bench command at N = 2 with the final code (block_solve summary now included at N > 1)
tag=${1:-r02r}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n2.err
tail -c 300 gpurun_out/bench_${tag}_n2.err; python - <<P
import json
d = json.loads([l for l in open("gpurun_out/bench_${tag}_n2.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["efficiency_same_lattice"], d["e2e"]["ms_per_step"], d["true_residual"])
print(json.dumps(d["block_solve"]))
P
