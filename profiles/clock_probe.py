"""Samples SM clock / power (nvidia-smi, 50 ms) while each stencil variant runs back to back for
about a second: separates 'slower because the kernel is' from 'slower because the power cap
lowered the clock'.  Output: gpurun_out/clock_probe.json"""
import json
import os
import statistics
import subprocess
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu",
                                   "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
        self.mark = 0
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.p.stdout:
            try:
                self.rows.append([float(x) for x in line.split(",")])
            except ValueError:
                pass

    def begin(self):
        self.mark = len(self.rows)

    def end(self):
        r = self.rows[self.mark:]
        if len(r) > 4:
            r = r[2:]          # skip ramp-up samples
        if not r:
            return {}
        return {"sm_mhz": statistics.median(x[0] for x in r), "mem_mhz": statistics.median(x[1] for x in r),
                "power_w": statistics.median(x[2] for x in r), "temp_c": max(x[3] for x in r), "samples": len(r)}


dims = (32, 32, 32, 64)
ctx = api.Context(dims)
ctx.links_synthetic(1234, 0)
s = Sampler()
out = []
for prec, k in ((2, 1), (2, 2), (2, 4), (1, 1), (1, 2), (1, 3), (1, 4), (0, 1)):
    for rep in range(2):
        s.begin()
        n = 4000 if prec != 2 else 2500
        ms = ctx.dslash_time(prec, 2, n) if k == 1 else ctx.dslash_block_time(prec, k, 2, n)
        c = s.end()
    out.append(dict(prec=prec, nrhs=k, ms=ms, **c))
    print(out[-1], flush=True)
vb = [ctx.vec_create() for _ in range(4)]
vx = [ctx.vec_create() for _ in range(4)]
for k in range(4):
    ctx.vec_gaussian(vb[k], 2, 5678 + 101 * k)
for mixed in (0, 1):
    for rep in range(2):
        for k in range(4):
            ctx.vec_zero(vx[k], 2)
        s.begin()
        it, res = ctx.congrad_block_dev(vb, vx, 0.05, 2, 2000, 10, 1e-10, mixed_precision=mixed)
        c = s.end()
    out.append(dict(solve="block4", mixed=mixed, iters=it, seconds=res[0]["device_seconds"], **c))
    print(out[-1], flush=True)
for mixed in (0, 1, 2):
    for rep in range(3):
        ctx.vec_zero(vx[0], 2)
        s.begin()
        it, res = ctx.congrad_dev(vb[0], vx[0], 0.05, 2, 2000, 10, 1e-10, mixed_precision=mixed)
        c = s.end()
    out.append(dict(solve="single", mixed=mixed, iters=it, seconds=res["device_seconds"], **c))
    print(out[-1], flush=True)
s.p.terminate()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/clock_probe.json", "w"), indent=1)
ctx.close()
