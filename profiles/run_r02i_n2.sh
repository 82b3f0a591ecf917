#!/bin/bash
# 2-GPU box: fused halo push A/B over real NVLink at the 8-GPU local volume (t split only), parity first.
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    tests/mgpu_check.py --dims 8 8 12 24 > gpurun_out/mgpu_check_${tag}_n2.log 2>&1
echo "rc=$?" >> gpurun_out/mgpu_check_${tag}_n2.log
for fp in 0 1 0 1; do
  B200KS_FUSED_PUSH=$fp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29510 + fp)) \
      bench.py --gpus 2 --lattice 64 64 32 48 --steps 5 --warmup 3 --no-extras --no-cpu-baseline >> gpurun_out/bench_${tag}_fused${fp}.json 2>> gpurun_out/bench_${tag}_fused${fp}.err
done
tail -n 4 gpurun_out/mgpu_check_${tag}_n2.log
for fp in 0 1; do python - <<P
import json
for ln in open("gpurun_out/bench_${tag}_fused${fp}.json"):
    if not ln.startswith("{"): continue
    d = json.loads(ln); print("fused ${fp}", d["ms_per_step"], d["cg_iters_per_solve"], d["roofline"]["16bit"], d["e2e"]["ms_per_step"], d["true_residual"])
P
done
