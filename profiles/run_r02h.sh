#!/bin/bash
# GPU-box script (1 GPU): fused halo push in the 16-bit iteration, checked and measured with the GPU as its own neighbour.
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "forced_self_partition" --durations=5 > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
for fp in 0 1; do
  B200KS_FUSED_PUSH=$fp timeout 300 python profiles/halo_probe.py > gpurun_out/halo_probe_${tag}_fused${fp}.json 2> gpurun_out/halo_probe_${tag}_fused${fp}.err
done
grep -v "^\s*$" gpurun_out/pytest_${tag}.log | tail -n 30
for fp in 0 1; do python - <<P
import json
d = json.load(open("gpurun_out/halo_probe_${tag}_fused${fp}.json"))
for k, v in d.items(): print("fused ${fp}", k, "f16 dslash ms", v["dslash_ms_f16"], "cg us/iter mixed2", v["cg_mixed2"]["us_per_iter"], "mixed1", v["cg_mixed1"]["us_per_iter"], "iters", v["cg_mixed2"]["iters"])
P
done
