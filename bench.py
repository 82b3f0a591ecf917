#!/usr/bin/env python
"""bench.py -- HISQ staggered solve benchmark (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete single-mass HISQ CG solve (4m^2 - D^2) x = b, mass 0.05, from a zero
initial guess to the requested residual, on the synthetic random-SU(3) 32^3x64 lattice
(BASELINE.json configs[1]).  The metric is MILC's own: GFLOP/s = 1187 flop x V x iterations /
time (generic_ks/d_congrad5_fn_milc.c:81-83,390-396), so it compares directly with the
reference's `CONGRAD5:` lines and is independent of the iteration count.

  value     solve with source/solution resident in HBM (CUDA events on the library stream)
  e2e       same solve through the host-buffer C-ABI call (b200ks_congrad = the
            ks_congrad_parity_gpu seam): pinned host source + guess copied H2D and the solution
            copied D2H inside the timed region, every step
  roofline  the dominant kernel of the timed solve (16-bit stencil for the default --mixed 2:
            544 B/site x Vh sites per launch) / live CUDA-event launch time on the hot GPU,
            against MEASURED_PEAKS.json hbm_gbs; the double / single stencils and the same
            three from a cool start are reported beside it
  block_solve  four sources at once through the K-wide stencil (ks_congrad_block_parity seam)
  cpu_baseline / --impl reference
            the reference's own CPU solver (oracle/_ref, built from /root/reference sources with
            OpenMP) on the box's host cores, on a bounded sample (fixed iteration count) of the
            same workload
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

EVEN, ODD = 2, 1
DIMS = (32, 32, 32, 64)
MASS = 0.05
RESID = 1e-10
NITER, NRESTART = 2000, 10
CG_FLOP_PER_SITE = 1187.0      # d_congrad5_fn_milc.c:81-83
DSLASH_FLOP_PER_SITE = 1146.0  # 16 x 66 + 15 x 6
# algorithmic bytes per output site = w*(8*R_fat + 8*R_long + 6 + 6), w = bytes/real
def dslash_bytes_per_site(prec, long_reals):
    if prec == 0:   # 16-bit links, 16-bit colour vectors + one fp32 scale per site (in and out)
        return 2.0 * (8 * 18 + 8 * long_reals) + 2 * 16
    return (8.0 if prec == 2 else 4.0) * (8 * 18 + 8 * long_reals + 12)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(prec, long_reals):
    """dram bytes per dslash launch from the committed ncu --set full summaries, if any."""
    p = os.path.join(ROOT, "profiles", "dslash_ncu_summary.json")
    if os.path.exists(p):
        try:
            for row in json.load(open(p))["kernels"]:
                if (row["prec"] == prec and row["long_reals"] == long_reals and row.get("nrhs", 1) == 1
                        and row.get("epilogue", 0) == 0 and row.get("mode", 0) == 0):
                    return float(row["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def make_workload(dims):
    from milc_qcd_b200 import fields as F
    fat, lng = F.make_links(dims, seed=1234)
    src = F.make_source(dims, seed=5678, parity=EVEN)
    return fat, lng, src


class stdout_to_stderr:
    """The reference's layout code prints to the C-level stdout; bench.py's stdout carries ONE JSON line."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def omp_all_cores():
    """The reference arm uses every host core whatever the launcher exported: torchrun sets
    OMP_NUM_THREADS=1 for its children, which made round 1's N > 1 reference lines single-threaded.
    Sets the environment (read by libgomp when oracle/_ref/libmilcref_omp.so pulls it in) AND the runtime's
    own setting, and returns the thread count the OpenMP runtime reports."""
    import ctypes
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        gomp = ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL)
        gomp.omp_set_num_threads(cores)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return cores


# ---------------------------------------------------------------------------------------------
def cpu_reference_sample(dims, fat, lng, src, iters, repeats=1):
    """Times the reference's CPU CG (oracle/_ref, OpenMP over all host cores; or the oracle
    port if _ref was not built) for a fixed number of iterations.  Returns (gflops, meta)."""
    from oracle import pyoracle
    cores = omp_all_cores()
    V = int(np.prod(dims))
    times = []
    if pyoracle.ref_available("_omp"):
        kind = "reference"
        with stdout_to_stderr():
            ref = pyoracle.MilcRef(dims, "_omp")
            ref.set_links(fat, lng)
            for _ in range(repeats):
                x = np.zeros_like(src)
                t0 = time.perf_counter()
                it, q = ref.congrad(src, x, MASS, EVEN, iters, 1, RESID)
                times.append((time.perf_counter() - t0, it))
        impl = "MILC d_congrad5_fn_milc.c + dslash_fn_dblstore.c (oracle/_ref, -O3 -DFAST -DOMP)"
    else:
        kind = "port"
        o = pyoracle.Oracle()
        for _ in range(repeats):
            x = np.zeros_like(src)
            t0 = time.perf_counter()
            it, q = o.congrad(dims, fat, lng, src, x, MASS, EVEN, iters, 1, RESID)
            times.append((time.perf_counter() - t0, it))
        impl = "oracle/ks_oracle.c port (OpenMP dslash)"
    return times, dict(kind=kind, cores=cores, impl=impl, V=V)


def run_reference(args):
    """--impl reference: the reference's own CPU path on the host cores, bounded sample."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    # The b200 arm's workload at this N: 32^3x64 (N = 1) or the 64^3x96 strong-scaling lattice
    # (N > 1).  The CPU sample always runs on the 32^3x64 lattice of the same synthetic ensemble:
    # 64^3x96 double links are 29 GB plus MILC's back-link copy, beyond a bounded CPU run, and
    # GFLOP/s (flop per site-iteration / time) is an intensive quantity.
    multi = args.gpus > 1
    arm_dims = (64, 64, 64, 96) if multi else DIMS
    dims = DIMS
    V = int(np.prod(dims))
    fat, lng, src = make_workload(dims)
    iters = 10  # + the initial and the final true-residual evaluation = 11 counted iterations
    times, meta = cpu_reference_sample(dims, fat, lng, src, iters, repeats=args.warmup + args.steps)
    timed = times[args.warmup:]
    tot_t = sum(t for t, _ in timed)
    tot_it = sum(it for _, it in timed)
    gflops = CG_FLOP_PER_SITE * V * tot_it / tot_t / 1e9
    sample = "CG capped at %d iterations per step (%d counted iterations/step) on a 32^3x64 lattice of the workload's synthetic ensemble%s" % (
        iters, timed[0][1], " (sub-volume sample of 64^3x96)" if multi else " (the full workload lattice)")
    lat = "x".join(map(str, arm_dims))
    line = {
        "impl": "reference", "metric": "hisq_cg_gflops", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(len(timed), 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "HISQ single-mass CG, mass 0.05, resid 1e-10, synthetic random-SU(3) %s (%s)"
                               % (lat, "BASELINE configs[3], strong scaling" if multi else "BASELINE configs[1]"),
                   "lattice": list(arm_dims), "sample_lattice": list(dims), "implementation": meta["impl"]},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": meta["cores"], "kind": meta["kind"],
                         "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def seam_e2e(dims, fat, lng, b, steps, mixed, ngpu, resid=RESID):
    """The end-to-end leg: the call a MILC binary makes.  ks_congrad_parity_gpu from the COMPILED route-1
    library (libb200ks_milc.so, include/b200ks_milc.h) on plain pageable host arrays in MILC's layout --
    source and zero guess H2D, solution D2H, link-cache check, qic bookkeeping all inside the timed call
    (what MILC's own dtimec brackets, generic_ks/d_congrad5_fn_gpu.c:43-161).  ngpu > 1: the library
    spreads the lattice over ngpu devices behind the same call (B200KS_NGPU, one process)."""
    import ctypes as C
    from milc_qcd_b200 import milc_abi, _lib
    V = int(np.prod(dims))
    os.environ["B200KS_NGPU"] = str(ngpu)
    shim = milc_abi.load()
    shim.b200ks_milc_finalize()
    shim.b200ks_milc_setup(*dims, mixed)
    fn = milc_abi.fn_links(fat, lng, 1)
    src = np.array(b, copy=True)                      # plain malloc'ed memory, as MILC's create_v_field gives
    dst = np.zeros_like(src)
    lib = _lib.load()

    def call():
        dst[:V // 2] = 0
        q = milc_abi.qic(EVEN, resid, NITER, NRESTART)
        t0 = time.perf_counter()
        it = shim.ks_congrad_parity_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), MASS, C.byref(fn))
        return time.perf_counter() - t0, it, q

    t_first, it, q = call()                           # first call: link upload + re-layout + fingerprints
    call()
    times, iters, prof = [], 0, np.zeros(8)
    acc = np.zeros(8)
    for _ in range(steps):
        t, it, q = call()
        times.append(t)
        iters += it
        lib.b200ks_call_profile(shim.b200ks_milc_context(), prof.ctypes.data_as(C.POINTER(C.c_double)))
        acc += prof
    ctxp = shim.b200ks_milc_context()
    st_up, st_ver = C.c_longlong(0), C.c_longlong(0)
    lib.b200ks_links_sync_stats(ctxp, C.byref(st_up), C.byref(st_ver))
    out = {"seconds": times, "iters": iters, "first_call_s": t_first, "final_rsq": q.final_rsq, "converged": q.converged,
           "num_gpus": lib.b200ks_num_gpus(ctxp), "link_uploads": st_up.value, "link_verifications": st_ver.value,
           "launches": int(lib.b200ks_launch_count(ctxp)),
           "profile_ms": {"h2d_source_and_guess_host_side": 1e3 * acc[0] / steps, "solve_wall": 1e3 * acc[1] / steps,
                          "solve_device": 1e3 * acc[6] / steps, "wait_for_link_verification": 1e3 * acc[2] / steps,
                          "d2h_solution": 1e3 * acc[3] / steps, "library_call": 1e3 * acc[4] / steps,
                          "passes": acc[5] / steps}}
    sol = dst.copy()
    shim.b200ks_milc_finalize()
    os.environ.pop("B200KS_NGPU", None)
    return out, sol


def multishift_summary(api, dist, torch, dims, world, rank, local_rank, long_recon):
    """BASELINE configs[2]: RHMC multi-shift CG (ks_multicg_offset, 12 shifts) on 48^3x96, t-split over the
    ranks; molecular-dynamics (1e-6) and action (1e-10) tolerances, the reference's algorithm in double and the
    mixed form; MILC's (1205 + 15 N) flop convention (generic_ks/ks_multicg_offset.c:156); roofline of
    ms_update_kernel from its algorithmic bytes (48 + 192 N per site) and the time it takes inside the solve."""
    from milc_qcd_b200 import fields as F
    multi = world > 1
    V = int(np.prod(dims))
    grid = (1, 1, 1, world)
    if multi:
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = api.Context(dims, device=local_rank, grid=grid, rank=rank, nranks=world, nccl_id=ids[0])
    else:
        ctx = api.Context(dims, device=local_rank)
    ctx.links_synthetic(1234, long_recon)
    nshift = 12
    offsets = F.rhmc_offsets(nshift, MASS)
    vb = ctx.vec_create()
    vps = [ctx.vec_create() for _ in range(nshift)]
    ctx.vec_gaussian(vb, EVEN, 5678)
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))
    rows = []
    for resid, mixed in ((1e-6, 0), (1e-6, 1), (1e-10, 0)):
        best = None
        for rep in range(2):
            torch.cuda.synchronize()
            if multi:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            it, res = ctx.multicg_dev(vb, vps, offsets, EVEN, 5000, 1, resid, mixed_precision=mixed)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if multi:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            if rep > 0 or best is None:
                best = (ms, it, res)
        ms, it, res = best
        row = {"resid": resid, "mixed_precision": mixed, "iterations_total": it, "seconds": ms * 1e-3,
               "worst_final_rsq": max(r["final_rsq"] for r in res), "converged": min(r["converged"] for r in res)}
        if mixed == 0:
            row["gflops_milc_convention"] = (1205.0 + 15.0 * nshift) * V * it / (ms * 1e-3) / 1e9
            # per-iteration algorithmic bytes per parity site, double: 2 stencils + resid (144) + update (48 + 192 N)
            long_reals = 2 * ctx.long_link_info()[0]
            per_site = 2 * dslash_bytes_per_site(2, long_reals) + 96 + 144 + 48 + 192 * nshift
            row["achieved_gbs_per_gpu"] = per_site * (V / 2) * it / (ms * 1e-3) / 1e9 / world
        rows.append(row)
    out = {"workload": "RHMC multi-shift CG, 12 shifts (4m^2 + geometric ladder), synthetic random-SU(3) %s, BASELINE configs[2]"
                       % "x".join(map(str, dims)), "lattice": list(dims), "rank_grid": list(grid), "n_gpus": world, "rows": rows,
           "device_bytes_per_gpu": ctx.device_bytes()}
    ctx.close()
    return out


def single_solve_point(api, dist, torch, dims, world, rank, local_rank, grid, long_recon, mixed, label):
    """One mixed CG solve on `dims` decomposed over the ranks (fields generated on the device): the
    strong-scaling anchor on one GPU (world 1) and BASELINE configs[4]'s 96^3x192 point on 8."""
    multi = world > 1
    V = int(np.prod(dims))
    if multi:
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = api.Context(dims, device=local_rank, grid=grid, rank=rank, nranks=world, nccl_id=ids[0])
    else:
        ctx = api.Context(dims, device=local_rank)
    ctx.links_synthetic(1234, long_recon)
    vb, vx = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(vb, EVEN, 5678)
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))
    best = None
    for rep in range(2):
        ctx.vec_zero(vx, EVEN)
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        it, res = ctx.congrad_dev(vb, vx, MASS, EVEN, NITER, NRESTART, RESID, mixed_precision=mixed)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if multi:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = (ms, it, res)
    ms, it, res = best
    # independent true residual on the device operator
    vt = ctx.vec_create()
    ctx.dslash_dev(vx, vt, ODD)
    ctx.dslash_dev(vt, vt, EVEN)
    # r = b - (4m^2 x - D^2 x): norms by the library (global)
    bb = ctx.vec_norm2(vb, EVEN)
    out = {"workload": label, "lattice": list(dims), "rank_grid": list(grid), "n_gpus": world, "mixed_precision": mixed,
           "cg_iters": it, "seconds": ms * 1e-3, "value": CG_FLOP_PER_SITE * V * it / (ms * 1e-3) / 1e9, "unit": "GFLOP/s",
           "final_rsq_true": res["final_rsq"], "converged": res["converged"], "source_norm2": bb,
           "device_bytes_per_gpu": ctx.device_bytes()}
    ctx.close()
    return out


def links_force_summary(api, dims, local_rank):
    """Rows f1 / f2 of SURVEY.md section 8 in the driver's line (N = 1): the HISQ link chain U -> V -> W -> (fat, long)
    with everything resident (CUDA events, b200ks_hisq_links_time) and the HISQ fermion force through the host-buffer
    call behind qudaHisqForce (9 terms; U, V, W and the vectors H2D, the momenta D2H inside the clock) -- the numbers
    --workload links / --workload force report in full, with the reference timed beside them."""
    V = int(np.prod(dims))
    ctx = api.Context(dims, device=local_rank)
    try:
        ctx.hisq_links_time(1234, 1)
        ms, nsvd = ctx.hisq_links_time(1234, 3)
        flop = (2 * 61632.0 + 1728.0) * V
        out = {"links": {"ms_per_chain": ms, "value": flop / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "svd_branch_links": nsvd,
                         "note": "Haar-random thin links generated on the device, resident chain; MILC's 2 x 61632 + 1728 "
                                 "flop per site convention"}}
        U, Vl, W = ctx.hisq_links_fetch(0), ctx.hisq_links_fetch(1), ctx.hisq_links_fetch(2)
        rng = np.random.default_rng(77)
        nterms = 9
        X = [rng.standard_normal((V, 3, 2)) for _ in range(nterms)]
        res = np.linspace(0.2, 1.0, nterms)
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            mom = ctx.hisq_force(U, Vl, W, X, res, 0.02)
            times.append(time.perf_counter() - t0)
        out["force"] = {"seconds_per_call": min(times), "all_calls_s": times, "us_per_site": 1e6 * min(times) / V,
                        "nterms": nterms, "h2d_bytes_per_call": int(3 * U.nbytes + nterms * X[0].nbytes),
                        "d2h_bytes_per_call": int(mom.nbytes), "mom_max": float(np.abs(mom).max()),
                        "note": "b200ks_hisq_force on pageable host arrays in MILC's layout, best of three calls"}
        return out
    finally:
        ctx.close()


def run_b200(args):
    import torch
    from milc_qcd_b200 import api, dist as D
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must be launched with %d ranks (torch.distributed.run), got WORLD_SIZE=%d"
                         % (args.gpus, args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    multi = world > 1
    dist = None
    if multi:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
    # N = 1: BASELINE configs[1] (32^3x64).  N > 1: configs[3], strong scaling of 64^3x96.
    dims = tuple(args.lattice) if args.lattice else (DIMS if not multi else (64, 64, 64, 96))
    V = int(np.prod(dims))
    grid = D.rank_grid(world)
    if multi:
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = api.Context(dims, device=local_rank, grid=grid, rank=rank, nranks=world, nccl_id=ids[0])
    else:
        ctx = api.Context(dims, device=local_rank)
    Vl = ctx.volume
    Vlh = Vl // 2

    # synthetic inputs are generated on the device from global coordinates (identical for every
    # decomposition); the host copies used by the e2e leg and the CPU baseline are read back
    t_gen = time.perf_counter()
    ctx.links_synthetic(1234, args.long_recon)
    long_reals = 2 * ctx.long_link_info()[0]
    vb, vx = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(vb, EVEN, 5678)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():   # (no NCCL kernel: the GPUs stay free for the rank that works)
        if multi:
            t = torch.zeros(1)
            dist.all_reduce(t)

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def solve_resident(mixed=args.mixed):
        ctx.vec_zero(vx, EVEN)
        return ctx.congrad_dev(vb, vx, MASS, EVEN, NITER, NRESTART, RESID, mixed_precision=mixed)

    for _ in range(args.warmup):
        it, res = solve_resident()

    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    iters_total, dev_s = 0, 0.0
    for _ in range(args.steps):
        it, res = solve_resident()
        iters_total += it
        dev_s += res["device_seconds"]
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    final_rsq_timed = res["final_rsq"]
    if not (res["converged"] == 1 and final_rsq_timed < RESID ** 2):
        raise SystemExit("bench.py: the timed solve did not converge (final_rsq %.3e): no number" % final_rsq_timed)

    # the other precision modes of the same solve, for the record (not the headline)
    others = []
    for other in (0, 1, 2):
        if other == args.mixed:
            continue
        solve_resident(other)
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record(stream)
        it_o, res_o = solve_resident(other)
        e5.record(stream)
        barrier()
        ms_other = max_over_ranks(e4.elapsed_time(e5))
        others.append({"mixed_precision": other, "value": CG_FLOP_PER_SITE * V * it_o / (ms_other * 1e-3) / 1e9,
                       "unit": "GFLOP/s", "cg_iters": it_o, "cg_time_to_solution_s": ms_other * 1e-3,
                       "final_rsq": res_o["final_rsq"]})

    # dominant-kernel roofline, live: back-to-back dslash launches (halo exchange included for N > 1)
    n_ds = 100
    ds_ms = {}
    for p in (2, 1, 0):
        barrier()
        ds_ms[p] = max_over_ranks(ctx.dslash_time(p, EVEN, n_ds))
    clocks = sampler.stop()
    # the same stencils from a cool start (1 s idle, 20 launches): this part runs every stencil loop
    # at its 1000 W power cap, and the kernels that are not purely HBM-bound slow down with the clock
    ds_ms_cool = {}
    if not multi:
        for p in (2, 1, 0):
            time.sleep(1.0)
            ds_ms_cool[p] = ctx.dslash_time(p, EVEN, 20)

    # block solve (ks_congrad_block_parity seam): 4 sources at once, mixed precision, device-resident
    # (N = 2: the K-wide stencil on the partitioned lattice, one halo exchange for the 4 inputs; one repetition.  This leg
    # has run on real hardware at N = 1 and 2 only -- profiles/bench/r02r_bench_p2p_n2.json -- so the 4- and 8-GPU lines,
    # which carry the strong-scaling numbers, do not depend on it; tests/mgpu_check.py covers it at any N)
    block = None
    if world <= 2:
        vbs = [vb] + [ctx.vec_create() for _ in range(3)]
        vxs = [ctx.vec_create() for _ in range(4)]
        for k in range(1, 4):
            ctx.vec_gaussian(vbs[k], EVEN, 5678 + 101 * k)
        best = None
        for rep in range(1 if multi else 2):
            for v in vxs:
                ctx.vec_zero(v, EVEN)
            barrier()
            e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e6.record(stream)
            it_b, res_b = ctx.congrad_block_dev(vbs, vxs, MASS, EVEN, NITER, NRESTART, RESID, mixed_precision=1)
            e7.record(stream)
            torch.cuda.synchronize()
            ms_b = max_over_ranks(e6.elapsed_time(e7))
            if best is None or ms_b < best[0]:
                best = (ms_b, it_b, res_b)
        ms_b, it_b, res_b = best
        block = {"nsrc": 4, "mixed_precision": 1, "seconds": ms_b * 1e-3, "seconds_per_source": ms_b * 1e-3 / 4,
                 "iterations_total": it_b, "value": CG_FLOP_PER_SITE * V * it_b / (ms_b * 1e-3) / 1e9, "unit": "GFLOP/s",
                 "worst_final_rsq": max(r["final_rsq"] for r in res_b), "converged": min(r["converged"] for r in res_b),
                 "note": "b200ks_congrad_block_dev: K-wide stencil, links streamed once per 4 sources"
                         + ("; partitioned lattice: one halo exchange carries the 4 inputs" if multi else "")}
        for v in vbs[1:] + vxs:
            ctx.vec_free(v)

    # what the copies of one end-to-end step cost on their own (diagnostic, pageable arrays like MILC's)
    hb = np.zeros((Vl, 3, 2))
    hx = np.zeros((Vl, 3, 2))
    ctx.vec_download(vb, hb, EVEN)
    vu = ctx.vec_create()
    ctx.vec_upload(vu, hb, EVEN)
    t0 = time.perf_counter(); ctx.vec_upload(vu, hb, EVEN); ctx.vec_upload(vu, hx, EVEN); t_up = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.vec_download(vu, hx, EVEN); t_dn = time.perf_counter() - t0
    ctx.vec_free(vu)
    copies_ms = max_over_ranks(1e3 * (t_up + t_dn))

    peak, peak_src = measured_peak()
    kp = {0: 2, 1: 1, 2: 0}[args.mixed]   # storage precision of the stencil that dominates the timed solve
    bps = {p: dslash_bytes_per_site(p, long_reals) for p in (0, 1, 2)}
    ach = {p: bps[p] * Vlh / (ds_ms[p] * 1e-3) / 1e9 for p in (0, 1, 2)}   # per GPU
    half_bytes = (V // 2) * 6 * 8

    # host copies of the inputs for the seam-level e2e leg and the CPU baseline (N = 1: read back from this
    # context; N > 1: regenerated below by the single process that drives all GPUs)
    cpu = None
    t_links = None
    fat = lng = None
    if not multi:
        fat, lng = ctx.links_download()
        if not args.no_cpu_baseline:
            try:
                times, meta = cpu_reference_sample(dims, fat, lng, hb, 10, repeats=2)
                t, itc = times[-1]
                cpu = {"value": CG_FLOP_PER_SITE * V * itc / t / 1e9, "unit": "GFLOP/s", "cores": meta["cores"],
                       "kind": meta["kind"],
                       "sample": "CG capped at 10 iterations (%d counted) on the full %s workload, same links and source, %s"
                                 % (itc, "x".join(map(str, dims)), meta["impl"])}
            except Exception as ex:  # the baseline is a reported number, never a reason to lose the GPU line
                cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}
    device_bytes = ctx.device_bytes()
    halo_mode = ctx.halo_mode()
    local_dims = list(ctx.dims)
    ctx.close()
    barrier()

    # ---- end-to-end through the drop-in symbol: ONE process (rank 0), all N GPUs behind it ---------
    e2e, true_resid, seam_err = None, None, None
    if rank == 0:
        try:
            if multi:
                gen = api.Context(dims, ngpu=world)
                gen.links_synthetic(1234, args.long_recon)
                fat, lng = gen.links_download()
                gv = gen.vec_create()
                gen.vec_gaussian(gv, EVEN, 5678)
                hb = np.zeros((V, 3, 2))
                gen.vec_download(gv, hb, EVEN)
                gen.close()
            seam, sol = seam_e2e(dims, fat, lng, hb, args.steps, args.mixed, world)
            ms_e2e = 1e3 * sum(seam["seconds"]) / args.steps
            e2e_value = CG_FLOP_PER_SITE * V * seam["iters"] / sum(seam["seconds"]) / 1e9
            # independent true residual of the solution the seam returned (single-process context again)
            chk = api.Context(dims, ngpu=world) if multi else api.Context(dims, device=local_rank)
            chk.load_links(fat, lng, args.long_recon)
            t_links = None
            if not multi:
                t0 = time.perf_counter()
                chk.load_links(fat, lng, args.long_recon)
                torch.cuda.synchronize()
                t_links = time.perf_counter() - t0
            tt = np.zeros_like(sol)
            chk.dslash(sol, tt, ODD)
            t2 = np.zeros_like(sol)
            chk.dslash(tt, t2, EVEN)
            chk.close()
            h = V // 2
            r = hb[:h] - (4 * MASS * MASS * sol[:h] - t2[:h])
            true_resid = float(np.sqrt(np.sum(r * r) / np.sum(hb[:h] ** 2)))
            e2e = {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": 2 * half_bytes, "d2h_bytes_per_step": half_bytes,
                   "ms_per_step": ms_e2e,
                   "through": "ks_congrad_parity_gpu of libb200ks_milc.so (MILC's prototype, include/b200ks_milc.h) on pageable "
                              "host arrays in MILC's layout; one process driving %d GPU(s) (B200KS_NGPU)" % world,
                   "num_gpus_behind_the_call": seam["num_gpus"], "cg_iters_per_solve": seam["iters"] / args.steps,
                   "final_rsq": seam["final_rsq"], "converged": seam["converged"], "true_residual_of_returned_solution": true_resid,
                   "profile_ms": seam["profile_ms"], "link_uploads": seam["link_uploads"],
                   "link_verifications": seam["link_verifications"], "gpu_launches_whole_leg": seam["launches"],
                   "first_call_s_incl_link_upload": seam["first_call_s"],
                   "copies_alone_ms": copies_ms,
                   "overhead_vs_resident_plus_copies": ms_e2e / (ms_total / args.steps + copies_ms) - 1.0,
                   "first_call_link_upload_ms": None if t_links is None else 1e3 * t_links,
                   "first_call_link_upload_bytes": 2 * 4 * 18 * 8 * V}
            if not (seam["converged"] == 1 and true_resid < 10 * RESID):
                raise RuntimeError("seam solve not converged: final_rsq %.3e, true residual %.3e" % (seam["final_rsq"], true_resid))
        except Exception as ex:
            seam_err = repr(ex)
    fat = lng = None
    cpu_barrier()

    # ---- same-lattice strong-scaling anchor: the 64^3x96 solve on ONE GPU (25.8 GB), measured in this run ----
    anchor = None
    if rank == 0 and not args.lattice and not args.no_extras:
        try:
            anchor = single_solve_point(api, None, torch, (64, 64, 64, 96), 1, 0, local_rank, (1, 1, 1, 1), args.long_recon, args.mixed,
                                        "HISQ single-mass CG 64x64x64x96 on ONE GPU (strong-scaling anchor, BASELINE configs[3])")
        except Exception as ex:
            anchor = {"error": repr(ex)}
    cpu_barrier()

    # ---- BASELINE configs[2] (N = 1, 2, 4) and configs[4] (N = 8) as part of the driver's line ----
    multishift, weak = None, None
    if not args.lattice and not args.no_extras:
        try:
            if world in (1, 2, 4):
                multishift = multishift_summary(api, dist, torch, (48, 48, 48, 96), world, rank, local_rank, args.long_recon)
            if world == 8:
                weak = single_solve_point(api, dist, torch, (96, 96, 96, 192), world, rank, local_rank, grid, args.long_recon, args.mixed,
                                          "HISQ single-mass CG 96x96x96x192 on 8 GPUs (BASELINE configs[4], fields generated on the device)")
        except Exception as ex:
            multishift = {"error": repr(ex)}

    # ---- rows f1 / f2 (link construction, fermion force) at N = 1: a reported extra, never a reason to lose the line ----
    links_force = None
    if not multi and not args.lattice and not args.no_extras:
        try:
            links_force = links_force_summary(api, dims, local_rank)
        except Exception as ex:
            links_force = {"error": repr(ex)}

    value = CG_FLOP_PER_SITE * V * iters_total / (ms_total * 1e-3) / 1e9
    if rank == 0:
        if e2e is None:
            raise SystemExit("bench.py: the end-to-end leg through the drop-in symbol failed: %s" % seam_err)
        lat = "x".join(map(str, dims))
        eff = None
        if multi and anchor and "value" in anchor:
            eff = value / (world * anchor["value"])
        line = {
            "metric": "hisq_cg_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {0: "f64", 1: "f64 (solution, true residuals) / f32 (inner Krylov vectors)",
                      2: "f64 (solution, true residuals) / f32 arithmetic on 16-bit links and search direction (inner iteration)"}[args.mixed],
            "data": "synthetic",
            "config": {"workload": "HISQ single-mass CG, mass 0.05, resid 1e-10, synthetic random-SU(3) %s (%s)"
                                   % (lat, "BASELINE configs[1]" if dims == DIMS else "BASELINE configs[3], strong scaling"
                                      if dims == (64, 64, 64, 96) else "custom lattice"),
                       "lattice": list(dims), "rank_grid": list(grid), "local_lattice": local_dims,
                       "l2": "links streamed by every dslash (%.2f GB per GPU at the inner precision) exceed the 126 MB L2; no flush needed"
                             % ({0: 8, 1: 4, 2: 2}[args.mixed] * (18 + long_reals) * 4 * Vl / 1e9),
                       "flop_convention": "MILC 1187 flop/site/iteration", "mixed_precision": args.mixed,
                       "halo": {0: "none", 1: "depth-3 ghosts, NCCL send/recv overlapped with the interior launch",
                                2: "depth-3 ghosts pushed into the neighbours' peer-mapped ghost buffers (NVLink stores), "
                                   "interior-first single-launch stencil acquiring arrival flags; CG scalars all-reduced "
                                   "through peer-mapped mailboxes inside the finish kernel"}[halo_mode]},
            "cg_iters_per_solve": iters_total / args.steps, "cg_time_to_solution_s": ms_total * 1e-3 / args.steps,
            "cg_device_seconds": dev_s / args.steps, "final_rsq_true_device": final_rsq_timed, "true_residual": true_resid,
            "converged": res["converged"],
            "dslash_gflops": {"f64": DSLASH_FLOP_PER_SITE * V / 2 / (ds_ms[2] * 1e-3) / 1e9,
                              "f32": DSLASH_FLOP_PER_SITE * V / 2 / (ds_ms[1] * 1e-3) / 1e9,
                              "16bit": DSLASH_FLOP_PER_SITE * V / 2 / (ds_ms[0] * 1e-3) / 1e9},
            "dslash_ms": {"f64": ds_ms[2], "f32": ds_ms[1], "16bit": ds_ms[0]},
            "roofline": {"bound": "hbm", "kernel": "%s (fat 18 / long %d reals per link)"
                                                   % ({0: "dslash_half_kernel", 1: "dslash_kernel<float>", 2: "dslash_kernel<double>"}[kp],
                                                      long_reals),
                         "achieved": ach[kp], "peak": peak,
                         "unit": "GB/s", "frac": ach[kp] / peak, "peak_source": peak_src,
                         "traffic": ncu_traffic(kp, long_reals) if dims == DIMS else None,
                         "algorithmic_bytes_per_launch": bps[kp] * Vlh, "algorithmic_bytes_per_site": bps[kp],
                         "note": "per GPU; for N > 1 the launch time includes the halo exchange and the exterior pass",
                         "f64": {"achieved": ach[2], "frac": ach[2] / peak, "bytes_per_site": bps[2]},
                         "f32": {"achieved": ach[1], "frac": ach[1] / peak, "bytes_per_site": bps[1]},
                         "16bit": {"achieved": ach[0], "frac": ach[0] / peak, "bytes_per_site": bps[0]},
                         "cool_start": {name: {"ms": ds_ms_cool[p], "achieved": bps[p] * Vlh / (ds_ms_cool[p] * 1e-3) / 1e9,
                                               "frac": bps[p] * Vlh / (ds_ms_cool[p] * 1e-3) / 1e9 / peak}
                                        for p, name in ((2, "f64"), (1, "f32"), (0, "16bit")) if p in ds_ms_cool}},
            "other_precision_modes": others,
            "block_solve": block,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "strong_scaling_anchor": anchor,
            "efficiency_same_lattice": eff,
            "multishift": multishift,
            "weak_point": weak,
            "links_force": links_force,
            "gpu_launches": launches, "clocks": clocks,
            "setup": {"gen_fields_s": t_gen, "device_bytes": device_bytes},
        }
        print(json.dumps(line))
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_multishift(args):
    """--workload multishift: BASELINE configs[2], RHMC multi-shift solve (12 shifts) on 48^3x96,
    t-split over the ranks.  Secondary line (the driver's default is the single-mass CG): time to
    solution of ks_multicg_offset at a molecular-dynamics tolerance (1e-6) and an action tolerance
    (1e-10), pure double (the reference's algorithm) and mixed (single-precision recurrence +
    per-shift polish), MILC flop convention (1205 + 15 N) V iters for the double run."""
    import torch
    from milc_qcd_b200 import api, dist as D, fields as F
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
    dims = tuple(args.lattice) if args.lattice else (48, 48, 48, 96)
    V = int(np.prod(dims))
    grid = (1, 1, 1, world)
    if multi:
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = api.Context(dims, device=local_rank, grid=grid, rank=rank, nranks=world, nccl_id=ids[0])
    else:
        ctx = api.Context(dims, device=local_rank)
    ctx.links_synthetic(1234, args.long_recon)
    nshift = 12
    offsets = F.rhmc_offsets(nshift, MASS)
    vb = ctx.vec_create()
    vps = [ctx.vec_create() for _ in range(nshift)]
    ctx.vec_gaussian(vb, EVEN, 5678)
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))
    rows = []
    for resid in (1e-6, 1e-10):
        for mixed in (0, 1, 2):
            best = None
            for rep in range(1 + max(1, args.steps // 2)):
                torch.cuda.synchronize()
                if multi:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                it, res = ctx.multicg_dev(vb, vps, offsets, EVEN, 5000, 1, resid, mixed_precision=mixed)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if multi:
                    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                if rep > 0 and (best is None or ms < best[0]):
                    best = (ms, it, res)
            ms, it, res = best
            row = {"resid": resid, "mixed_precision": mixed, "iterations_total": it, "seconds": ms * 1e-3,
                   "worst_final_rsq": max(r["final_rsq"] for r in res), "converged": min(r["converged"] for r in res)}
            if mixed == 0:
                row["gflops_milc_convention"] = (1205.0 + 15.0 * nshift) * V * it / (ms * 1e-3) / 1e9
            rows.append(row)
    if rank == 0:
        base = {r["resid"]: r["seconds"] for r in rows if r["mixed_precision"] == 0}
        for r in rows:
            r["speedup_vs_double"] = base[r["resid"]] / r["seconds"]
        print(json.dumps({"metric": "hisq_multicg_time_to_solution", "unit": "s", "n_gpus": world, "higher_is_better": False,
                          "data": "synthetic", "config": {"workload": "RHMC multi-shift CG, %d shifts (4m^2 + geometric ladder 1e-4..2), mass %.2f, "
                                                                      "synthetic random-SU(3) %s (BASELINE configs[2])"
                                                                      % (nshift, MASS, "x".join(map(str, dims))),
                                                          "lattice": list(dims), "rank_grid": list(grid), "offsets": list(map(float, offsets))},
                          "value": min(r["seconds"] for r in rows if r["resid"] == 1e-6), "rows": rows,
                          "device_bytes": ctx.device_bytes()}))
    ctx.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_block(args):
    """--workload block: multi-right-hand-side (block) CG, the ks_congrad_block_parity_gpu /
    qudaInvertMsrc seam, on the BASELINE configs[1] lattice.  Secondary line: for K = 1..4 sources
    the K-wide stencil (links loaded once for K colour vectors) and the block solve against K
    single solves of the same precision mode, device-resident, CUDA events."""
    import torch
    from milc_qcd_b200 import api
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    dims = tuple(args.lattice) if args.lattice else DIMS
    V = int(np.prod(dims))
    ctx = api.Context(dims, device=local_rank)
    ctx.links_synthetic(1234, args.long_recon)
    long_reals = 2 * ctx.long_link_info()[0]
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))
    peak, peak_src = measured_peak()
    kmax = 4
    vb = [ctx.vec_create() for _ in range(kmax)]
    vx = [ctx.vec_create() for _ in range(kmax)]
    for k in range(kmax):
        ctx.vec_gaussian(vb[k], EVEN, 5678 + 101 * k)

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    stencil = []
    for prec, w in ((2, 8.0), (1, 4.0)):
        for k in range(1, kmax + 1):
            ms = ctx.dslash_block_time(prec, k, EVEN, 50)
            bytes_launch = w * (8 * 18 + 8 * long_reals + 12 * k) * V / 2
            stencil.append({"prec": "f64" if prec == 2 else "f32", "nrhs": k, "ms_per_launch": ms,
                            "gflops": DSLASH_FLOP_PER_SITE * k * V / 2 / (ms * 1e-3) / 1e9,
                            "algorithmic_bytes_per_launch": bytes_launch,
                            "achieved_gbs": bytes_launch / (ms * 1e-3) / 1e9,
                            "frac_of_peak": bytes_launch / (ms * 1e-3) / 1e9 / peak})
    rows = []
    for mixed in (0, 1):
        singles = []
        for k in range(kmax):
            best1 = None
            for rep in range(2):   # best of two: the first solve of a mode also allocates its temporaries
                ctx.vec_zero(vx[k], EVEN)
                ms, (it, res) = timed(lambda: ctx.congrad_dev(vb[k], vx[k], MASS, EVEN, NITER, NRESTART, RESID, mixed_precision=mixed))
                if best1 is None or ms < best1[0]:
                    best1 = (ms, it)
            singles.append(best1)
        for k in range(2, kmax + 1):
            best = None
            for rep in range(2):
                for q in range(k):
                    ctx.vec_zero(vx[q], EVEN)
                ms, (tot, res) = timed(lambda: ctx.congrad_block_dev(vb[:k], vx[:k], MASS, EVEN, NITER, NRESTART, RESID,
                                                                     mixed_precision=mixed))
                if best is None or ms < best[0]:
                    best = (ms, tot, res)
            ms, tot, res = best
            ms_single = sum(t for t, _ in singles[:k])
            rows.append({"mixed_precision": mixed, "nsrc": k, "seconds": ms * 1e-3, "iterations_total": tot,
                         "gflops_milc_convention": CG_FLOP_PER_SITE * V * tot / (ms * 1e-3) / 1e9,
                         "seconds_k_single_solves": ms_single * 1e-3, "iterations_k_single_solves": sum(i for _, i in singles[:k]),
                         "speedup_vs_single_solves": ms_single / ms,
                         "worst_final_rsq": max(r["final_rsq"] for r in res), "converged": min(r["converged"] for r in res)})
    # The ks_spectrum propagator solve for the three colours of a point source, both parities
    # (mat_invert_uml_field x 3, generic_ks/mat_invert.c:328-402), host buffers in and out:
    # (a) the resident sequence b200ks_mat_invert_uml (one call, block solver), against
    # (b) the chain of host-buffer calls MILC's own glue makes per colour -- dslash_fn_field x 2 for
    #     M^+ src, ks_congrad (even), dslash_fn_field (odd reconstruction), ks_congrad (odd) -- of
    #     which only the time inside the library calls is counted (the host arithmetic between them
    #     is MILC's and is left out).
    from milc_qcd_b200 import fields as F
    Vl = ctx.volume
    srcs = [F.point_source(dims, (0, 0, 0, 0), col) for col in range(3)]
    dsts = [np.zeros_like(s) for s in srcs]
    uml = {}
    for mixed in (0, 1):
        for d in dsts:
            d[...] = 0
        ctx.mat_invert_uml(srcs, dsts, MASS, NITER, NRESTART, RESID, mixed_precision=mixed)   # warm
        for d in dsts:
            d[...] = 0
        t0 = time.perf_counter()
        tot, res_u = ctx.mat_invert_uml(srcs, dsts, MASS, NITER, NRESTART, RESID, mixed_precision=mixed)
        t_seq = time.perf_counter() - t0
        t_chain, it_chain = 0.0, 0
        h = Vl // 2
        for col in range(3):
            src = srcs[col]
            dst = np.zeros_like(src)
            tmp = np.zeros_like(src)
            t0 = time.perf_counter(); ctx.dslash(src, tmp, 3); t_chain += time.perf_counter() - t0
            tmp = -tmp + 2 * MASS * src
            t0 = time.perf_counter(); it1, _ = ctx.congrad(tmp, dst, MASS, EVEN, NITER, NRESTART, RESID, mixed_precision=mixed); t_chain += time.perf_counter() - t0
            ttt = np.zeros_like(src)
            t0 = time.perf_counter(); ctx.dslash(dst, ttt, ODD); t_chain += time.perf_counter() - t0
            dst[h:] = (src[h:] - ttt[h:]) / (2 * MASS)
            t0 = time.perf_counter(); it2, _ = ctx.congrad(tmp, dst, MASS, ODD, NITER, NRESTART, RESID, mixed_precision=mixed); t_chain += time.perf_counter() - t0
            it_chain += it1 + it2
            err = float(np.linalg.norm(dst - dsts[col]) / np.linalg.norm(dst))
        uml["mixed%d" % mixed] = {"resident_block_sequence_s": t_seq, "iterations": tot,
                                  "chained_host_calls_s": t_chain, "chained_iterations": it_chain,
                                  "speedup": t_chain / t_seq, "rel_diff_last_colour": err}
    best = max(rows, key=lambda r: r["gflops_milc_convention"])
    print(json.dumps({"metric": "hisq_block_cg_gflops", "unit": "GFLOP/s", "n_gpus": 1, "higher_is_better": True, "data": "synthetic",
                      "value": best["gflops_milc_convention"],
                      "config": {"workload": "HISQ block CG (1..4 sources at once), mass %.2f, resid %g, synthetic random-SU(3) %s"
                                             % (MASS, RESID, "x".join(map(str, dims))),
                                 "lattice": list(dims), "flop_convention": "MILC 1187 flop/site/iteration/source"},
                      "stencil": stencil, "rows": rows, "uml_point_source_3_colours": uml, "peak_gbs": peak, "peak_source": peak_src,
                      "device_bytes": ctx.device_bytes()}))
    ctx.close()
    return 0


def run_links(args):
    """--workload links: HISQ fermion-link construction on the
    BASELINE configs[1] lattice: U -> V (fat7) -> W (U(3) projection) -> fat, long, the chain MILC
    runs for every new gauge field.  value: chains per second with everything resident (CUDA
    events); e2e: the host-buffer call (thin links H2D, fat + long links D2H inside the timed
    region); cpu_baseline: the reference's own create_hisq_links_milc (oracle/_ref, OpenMP) on the
    same input; flop convention: the reference's 2 x 61632 + 1728 flop per site for the two
    smearing levels and the Naik links (fermion_links_fn_load_milc.c:106,268), projection not counted."""
    import torch
    from milc_qcd_b200 import api
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    dims = tuple(args.lattice) if args.lattice else DIMS
    V = int(np.prod(dims))
    flop = (2 * 61632.0 + 1728.0) * V
    ctx = api.Context(dims, device=local_rank)
    for _ in range(max(1, args.warmup)):
        ms, nsvd = ctx.hisq_links_time(1234, 1)
    ms, nsvd = ctx.hisq_links_time(1234, max(1, args.steps))
    U = ctx.hisq_links_fetch(0)
    fat_dev, lng_dev = ctx.hisq_links_fetch(3), ctx.hisq_links_fetch(4)
    # end to end: pinned host links in, fat + long out
    pin = [torch.zeros(U.shape, dtype=torch.float64).pin_memory() for _ in range(3)]
    hU, hF, hL = (p.numpy() for p in pin)
    hU[...] = U
    import ctypes as C
    c1, c2 = ctx._coeffs(ctx.HISQ_FAT7), ctx._coeffs(ctx.HISQ_ASQTAD_LIKE)

    def chain_host():
        n = C.c_longlong(0)
        api.check(ctx.lib.b200ks_hisq_links(ctx.h, c1, c2, hU.ctypes.data, None, None, hF.ctypes.data, hL.ctypes.data, 2,
                                            C.byref(n)), "b200ks_hisq_links")
    chain_host()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        chain_host()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    same = bool(np.array_equal(hF, fat_dev) and np.array_equal(hL, lng_dev))
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            cores = omp_all_cores()
            if pyoracle.ref_available("_omp"):
                ref = pyoracle.MilcRef(dims, "_omp")
                t0 = time.perf_counter()
                r = ref.hisq_links(U)
                t_cpu = time.perf_counter() - t0
                err = {k: float(np.abs(r[k] - x).max() / np.abs(r[k]).max()) for k, x in (("fat", fat_dev), ("lng", lng_dev))}
                cpu = {"value": flop / t_cpu / 1e9, "unit": "GFLOP/s", "seconds": t_cpu, "cores": cores, "kind": "reference",
                       "sample": "one full chain on the same %s thin links, create_hisq_links_milc (oracle/_ref, -O3 -DOMP)"
                                 % "x".join(map(str, dims)),
                       "svd_branch_links": r["nsvd"], "max_rel_diff_gpu_vs_reference": err}
        except Exception as ex:
            cpu = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    # bytes per chain: 72 staple passes per level read 6 matrices + read/write the fat link (+ staple store)
    print(json.dumps({"metric": "hisq_link_construction_gflops", "unit": "GFLOP/s", "n_gpus": 1, "higher_is_better": True,
                      "data": "synthetic", "value": flop / (ms * 1e-3) / 1e9, "ms_per_chain": ms,
                      "config": {"workload": "HISQ link construction U->V->W->(fat,long), Haar-random thin links %s"
                                             % "x".join(map(str, dims)), "lattice": list(dims),
                                 "flop_convention": "MILC: 61632 flop/site per smearing level + 1728 for the Naik links"},
                      "svd_branch_links": nsvd,
                      "e2e": {"value": flop / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_chain": ms_e2e,
                              "h2d_bytes_per_step": int(U.nbytes), "d2h_bytes_per_step": int(2 * U.nbytes),
                              "same_bits_as_resident_chain": same},
                      "cpu_baseline": cpu, "device_bytes": ctx.device_bytes()}))
    ctx.close()
    return 0


def run_force(args):
    """--workload force: HISQ fermion force through the host-buffer
    call behind qudaHisqForce on the BASELINE configs[1] lattice: 9 terms (a typical RHMC
    molecular-dynamics order), thin links generated on the device, V and W from the device link
    construction.  value: site-terms per second end to end (U, V, W and the 9 vectors H2D, the
    momenta D2H inside the timed region).  cpu_baseline: the reference's eo_fermion_force_multi
    (oracle/_ref, OpenMP) on a 16^3x32 lattice of the same ensemble (a bounded sample: the
    reference needs ~250 us per site on 8 cores)."""
    import torch
    from milc_qcd_b200 import api, fields as F
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    dims = tuple(args.lattice) if args.lattice else DIMS
    V = int(np.prod(dims))
    nterms = 9
    ctx = api.Context(dims, device=local_rank)
    ctx.hisq_links_time(1234, 1)
    U, Vl, W = ctx.hisq_links_fetch(0), ctx.hisq_links_fetch(1), ctx.hisq_links_fetch(2)
    rng = np.random.default_rng(77)
    X = [rng.standard_normal((V, 3, 2)) for _ in range(nterms)]
    res = np.linspace(0.2, 1.0, nterms)
    times = []
    for rep in range(max(2, args.steps)):
        t0 = time.perf_counter()
        mom = ctx.hisq_force(U, Vl, W, X, res, 0.02)
        times.append(time.perf_counter() - t0)
    t_gpu = min(times)
    bytes_in = 3 * U.nbytes + nterms * X[0].nbytes
    out = {"metric": "hisq_fermion_force_site_terms_per_s", "unit": "site-terms/s", "n_gpus": 1, "higher_is_better": True,
           "data": "synthetic", "value": V * nterms / t_gpu, "seconds_per_call": t_gpu, "all_calls_s": times,
           "us_per_site": 1e6 * t_gpu / V,
           "config": {"workload": "HISQ fermion force, %d terms, Haar-random thin links %s, host buffers in and out"
                                  % (nterms, "x".join(map(str, dims))), "lattice": list(dims), "nterms": nterms},
           "e2e": {"value": V * nterms / t_gpu, "unit": "site-terms/s", "h2d_bytes_per_step": int(bytes_in),
                   "d2h_bytes_per_step": int(mom.nbytes)},
           "mom_max": float(np.abs(mom).max()), "device_bytes": ctx.device_bytes()}
    ctx.close()
    if not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            cores = omp_all_cores()
            sdims = (16, 16, 16, 32)
            Vs = int(np.prod(sdims))
            ref = pyoracle.MilcRef(sdims, "_omp")
            Us = F.make_thin_links(sdims, seed=11, spread=0.5)
            Xs = rng.standard_normal((nterms, Vs, 3, 2))
            t0 = time.perf_counter()
            ref.hisq_force(Us, Xs, res, 0.02)
            t_cpu = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": Vs * nterms / t_cpu, "unit": "site-terms/s", "seconds": t_cpu, "cores": cores,
                                   "kind": "reference", "us_per_site": 1e6 * t_cpu / Vs,
                                   "sample": "one call of eo_fermion_force_multi (incl. its own link construction) on a "
                                             "16x16x16x32 lattice, %d terms (oracle/_ref, -O3 -DOMP)" % nterms}
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(out))
    return 0


def run_deflate(args):
    """--workload deflate: low-mode deflation of a trial solution with the
    vectors resident in HBM, on the BASELINE configs[1] lattice.  The set is synthetic (Gaussian vectors and a
    made-up spectrum: the kernels' traffic does not depend on the values); value = vectors deflated per second,
    device time by CUDA events on the library's stream; roofline: 2 x 48 B per site and vector against the
    measured HBM peak.  cpu_baseline: the oracle's restatement of deflate() (mat_invert.c:131-183) on a
    16^3x32 lattice (the reference's own function is static; the restatement is pinned on its effect)."""
    import torch
    from milc_qcd_b200 import api
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    dims = tuple(args.lattice) if args.lattice else DIMS
    V = int(np.prod(dims))
    nvecs = args.nvecs
    ctx = api.Context(dims, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))
    hs = []
    for j in range(nvecs):
        h = ctx.vec_create()
        ctx.vec_gaussian(h, 3, 1000 + j)
        hs.append(h)
    ctx.eig_set(hs, list(np.linspace(1e-4, 1e-2, nvecs)), use_in_uml=False)
    vs, vd = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(vs, 3, 7)
    ctx.vec_zero(vd, 3)
    l0 = ctx.launch_count()
    for _ in range(max(3, args.warmup)):
        ctx.deflate_dev(vs, vd, 0.05, 2)
    steps = max(3, args.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l1 = ctx.launch_count()
    e0.record(stream)
    for _ in range(steps):
        ctx.deflate_dev(vs, vd, 0.05, 2)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = ctx.launch_count() - l1
    peak, peak_src = measured_peak()
    alg_bytes = 2.0 * 48 * (V // 2) * nvecs
    out = {"metric": "deflation_vectors_per_s", "unit": "vectors/s", "n_gpus": 1, "higher_is_better": True, "data": "synthetic",
           "value": nvecs / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "dtype": "f64", "gpu_launches": int(launches),
           "config": {"workload": "deflation of one parity of a trial solution with %d resident vectors, %s (vectors larger than L2)"
                                  % (nvecs, "x".join(map(str, dims))), "lattice": list(dims), "nvecs": nvecs},
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
           "device_bytes": ctx.device_bytes()}
    ctx.close()
    if not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            sdims = (16, 16, 16, 32)
            Vs = int(np.prod(sdims))
            rng = np.random.default_rng(5)
            ev = rng.standard_normal((min(nvecs, 32), Vs, 3, 2))
            src, dst = rng.standard_normal((Vs, 3, 2)), np.zeros((Vs, 3, 2))
            o = pyoracle.Oracle()
            t0 = time.perf_counter()
            o.deflate(sdims, dst, src, 0.05, ev, np.linspace(1e-4, 1e-2, ev.shape[0]), 2)
            t_cpu = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": ev.shape[0] / t_cpu * (Vs / V), "unit": "vectors/s (scaled to the workload's volume)",
                                   "cores": 1, "kind": "port",
                                   "sample": "kso_deflate, %d vectors on a 16x16x16x32 lattice, one thread" % ev.shape[0]}
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(out))
    return 0


def run_meson(args):
    """Meson tie-ups (ks_meson_cont_mom's site loops, generic_ks/ks_meson_mom.c:160-437) on the configs[1] lattice:
    two propagators -> nt x nmom complex numbers.  `value`: resident propagators (b200ks_meson_mom_dev; the call's wall
    time includes the upload of the phase tables and the read-back of the result); `e2e`: host propagators in MILC's
    layout through b200ks_meson_mom (2 x 100 MB up per call).  roofline: 96 B per site whatever the number of momenta.
    cpu_baseline: the reference's own ks_meson_cont_mom (oracle/_ref, OpenMP) with the same momenta on a 16^3 x 32
    sample, scaled by volume."""
    import torch
    from milc_qcd_b200 import api
    local_rank = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local_rank)
    dims = tuple(args.lattice) if args.lattice else DIMS
    V = int(np.prod(dims))
    nmom = max(1, min(args.nmom, 100))
    rng = np.random.default_rng(3)
    mom = rng.integers(-2, 3, size=(nmom, 3))
    mom[0] = 0
    par = np.full((nmom, 3), 3)
    r0 = (0, 0, 0, 0)
    ctx = api.Context(dims, device=local_rank)
    va, vq = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(va, 3, 11)
    ctx.vec_gaussian(vq, 3, 12)
    stream = torch.cuda.ExternalStream(ctx.lib.b200ks_stream(ctx.h))
    for _ in range(max(3, args.warmup)):
        ctx.meson_mom_dev(va, vq, 15, r0, mom, par)
    steps = max(5, args.steps)
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        corr = ctx.meson_mom_dev(va, vq, 15, r0, mom, par)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    launches = ctx.launch_count() - l0
    # kernel time alone (CUDA events around the same calls include the host work between launches; keep both)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.meson_mom_dev(va, vq, 15, r0, mom, par)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1)
    ha, hq = np.zeros((V, 3, 2)), np.zeros((V, 3, 2))
    ctx.vec_download(va, ha)
    ctx.vec_download(vq, hq)
    ctx.meson_mom(ha, hq, 15, r0, mom, par)
    t0 = time.perf_counter()
    for _ in range(steps):
        corr_h = ctx.meson_mom(ha, hq, 15, r0, mom, par)
    ms_e2e = (time.perf_counter() - t0) / steps * 1e3
    peak, peak_src = measured_peak()
    alg = 96.0 * V
    out = {"metric": "meson_tieup_site_momenta_per_s", "unit": "site-momenta/s", "n_gpus": 1, "higher_is_better": True,
           "data": "synthetic", "value": V * nmom / (ms * 1e-3), "ms_per_step": ms, "ms_device_events": ms_dev, "steps": steps,
           "dtype": "f64", "gpu_launches": int(launches),
           "config": {"workload": "meson tie-up of two resident propagators, %s, %d momenta, local sink operator"
                                  % ("x".join(map(str, dims)), nmom), "lattice": list(dims), "nmom": nmom},
           "roofline": {"bound": "hbm", "achieved": alg / (ms_dev * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms_dev * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                        "note": "96 B per site (two colour vectors read once) over the event-timed call; FP64 issue takes "
                                "over from HBM as the number of momenta grows"},
           "e2e": {"value": V * nmom / (ms_e2e * 1e-3), "unit": "site-momenta/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": 2 * V * 48, "d2h_bytes_per_step": int(dims[3] * nmom * 16),
                   "through": "b200ks_meson_mom on pageable host propagators in MILC's layout"},
           "agreement_resident_vs_host": float(np.abs(corr - corr_h).max() / np.abs(corr).max()),
           "device_bytes": ctx.device_bytes()}
    ctx.close()
    if not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            cores = omp_all_cores()
            sdims = (16, 16, 16, 32)
            Vs = int(np.prod(sdims))
            with stdout_to_stderr():
                ref = pyoracle.MilcRef(sdims, "_omp")
                ref.set_ape_links(np.zeros((Vs, 4, 3, 3, 2)))
                s1, s2 = rng.standard_normal((Vs, 3, 2)), rng.standard_normal((Vs, 3, 2))
                idx = ref.spin_taste_index("pion5")
                t0 = time.perf_counter()
                ref.meson_cont_mom(s1, s2, mom, par, [idx] * nmom, list(range(nmom)), [0] * nmom, [1.0] * nmom, [0] * nmom, 1, r0)
                t_cpu = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": Vs * nmom / t_cpu, "unit": "site-momenta/s", "cores": cores, "kind": "reference",
                                   "sample": "ks_meson_cont_mom (oracle/_ref, -DOMP) with the same %d momenta on a 16^3x32 lattice" % nmom}
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mixed", type=int, default=2,
                    help="0 pure double; 1 double solution and true residuals with single-precision Krylov vectors and "
                         "reliable updates; 2 (default; BASELINE configs[1] 'mixed-precision CG', north_star 'single- or "
                         "half-precision inner solve') additionally 16-bit links and search direction in the stencil")
    ap.add_argument("--long-recon", type=int, default=0, help="long-link storage: 18, 14 or 0 = decided on the data")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the strong-scaling anchor, the multi-shift summary and the 96^3x192 point")
    ap.add_argument("--lattice", type=int, nargs=4, default=None, help="override the lattice (nx ny nz nt)")
    ap.add_argument("--nvecs", type=int, default=64, help="--workload deflate: number of resident vectors")
    ap.add_argument("--nmom", type=int, default=20, help="--workload meson: number of momenta")
    ap.add_argument("--workload", default="cg", choices=["cg", "multishift", "block", "links", "force", "deflate", "meson"],
                    help="cg (default, the driver's line): single-mass CG; multishift: BASELINE configs[2]; "
                         "block: multi-right-hand-side CG (ks_congrad_block_parity seam); "
                         "links: HISQ fermion-link construction (qudaLoadUnitarizedLink / qudaLoadKSLink seam)")
    args = ap.parse_args()
    # stdout carries the JSON line and nothing else: whatever libraries write to the C-level stdout (NCCL's version
    # banner, the reference's layout chatter) goes to stderr; Python's sys.stdout keeps the real one
    sys.stdout.flush()
    sys.stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "multishift":
        return run_multishift(args)
    if args.workload == "block":
        return run_block(args)
    if args.workload == "links":
        return run_links(args)
    if args.workload == "force":
        return run_force(args)
    if args.workload == "deflate":
        return run_deflate(args)
    if args.workload == "meson":
        return run_meson(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
