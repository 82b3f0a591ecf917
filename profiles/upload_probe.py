"""Host -> device upload of MILC-layout colour vectors through the library (b200ks_vec_upload: pageable numpy arrays
are bounced through pinned buffers by host threads, pinned arrays go direct) against the driver's own pageable and
pinned copies of the same bytes.  One GPU."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

dims = (32, 32, 32, 64)
V = int(np.prod(dims))
ctx = api.Context(dims)
v = ctx.vec_create()
src = np.random.default_rng(1).standard_normal((V, 3, 2))
pinned = torch.from_numpy(src).pin_memory()
dev = torch.empty(V * 6, dtype=torch.float64, device="cuda")
out = {"bytes": int(src.nbytes)}


def timed(f, n=10):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


for name, parity, nbytes in (("both parities", 3, src.nbytes), ("one parity", 2, src.nbytes // 2)):
    t = timed(lambda: ctx.vec_upload(v, src, parity))
    out["library, pageable numpy, " + name] = {"ms": t * 1e3, "GB/s": nbytes / t / 1e9}
    t = timed(lambda: ctx.vec_upload(v, pinned.numpy(), parity))
    out["library, pinned array, " + name] = {"ms": t * 1e3, "GB/s": nbytes / t / 1e9}
t = timed(lambda: dev.copy_(torch.from_numpy(src).view(-1)))
out["driver, pageable (torch copy_)"] = {"ms": t * 1e3, "GB/s": src.nbytes / t / 1e9}
t = timed(lambda: dev.copy_(pinned.view(-1), non_blocking=True))
out["driver, pinned (torch copy_)"] = {"ms": t * 1e3, "GB/s": src.nbytes / t / 1e9}
back = np.zeros_like(src)
t = timed(lambda: ctx.vec_download(v, back, 3))
out["library download, pageable numpy, both parities"] = {"ms": t * 1e3, "GB/s": src.nbytes / t / 1e9}
# host memcpy alone, 1 thread (numpy) for scale
dst = np.empty_like(src)
t0 = time.perf_counter()
for _ in range(5):
    np.copyto(dst, src)
out["host memcpy, one thread (numpy copyto)"] = {"GB/s": src.nbytes * 5 / (time.perf_counter() - t0) / 1e9}
out["host threads"] = os.cpu_count()
print(json.dumps(out, indent=1))
ctx.close()
