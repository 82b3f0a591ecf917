#!/bin/bash
# 4-GPU box: the driver's bench command at N = 4 (64^3x96 strong scaling + the 12-shift 48^3x96 summary on 4 GPUs)
tag=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29554 \
    bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n4.json 2> gpurun_out/bench_${tag}_n4.err
echo "bench rc=$?" >> gpurun_out/bench_${tag}_n4.err
tail -c 300 gpurun_out/bench_${tag}_n4.err; python - <<P
import json
d = json.loads([l for l in open("gpurun_out/bench_${tag}_n4.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["efficiency_same_lattice"], d["roofline"]["16bit"], d["e2e"]["ms_per_step"], d["true_residual"])
print(json.dumps(d["multishift"])[:900])
P
