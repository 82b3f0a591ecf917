#!/bin/bash
# 1 GPU: how boundary CTAs poll the arrival flags (B200KS_ACQ variants), self-partitioned stencil and CG at the 8-GPU local volume
tag=${1:-r02k}
mkdir -p gpurun_out
for a in 0 1 2 3; do
  B200KS_LIB=$PWD/profiles/variants/libb200ks_acq$a.so timeout 300 python profiles/halo_probe.py > gpurun_out/halo_probe_${tag}_acq$a.json 2> gpurun_out/halo_probe_${tag}_acq$a.err
done
for a in 0 1 2 3; do python - <<P
import json
d = json.load(open("gpurun_out/halo_probe_${tag}_acq$a.json"))
for k, v in d.items(): print("acq $a", k, "dslash ms f64 %.4f f32 %.4f f16 %.4f" % (v["dslash_ms_f64"], v["dslash_ms_f32"], v["dslash_ms_f16"]), "cg us/iter mixed2 %.1f mixed1 %.1f mixed0 %.1f" % (v["cg_mixed2"]["us_per_iter"], v["cg_mixed1"]["us_per_iter"], v["cg_mixed0"]["us_per_iter"]), "iters", v["cg_mixed2"]["iters"])
P
done
