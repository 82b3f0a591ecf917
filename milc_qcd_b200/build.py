"""Builds the in-tree native libraries with explicit nvcc/gcc commands (sm_100a only).

    python -m milc_qcd_b200.build            # build everything that is out of date
    python -m milc_qcd_b200.build --force

Outputs (git-ignored, shipped to the GPU box by gpurun):
    milc_qcd_b200/libb200ks.so        CUDA kernels + the C ABI (include/b200ks.h) + the quda* symbols
                                      MILC's own GPU glue binds (include/quda_milc_interface.h)
    milc_qcd_b200/libb200ks_milc.so   MILC-named solver symbols (include/b200ks_milc.h), PRECISION=2
    milc_qcd_b200/libb200ks_milc_f.so same, PRECISION=1
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")

NVCC_FLAGS = [
    "--threads", "3",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    out = os.path.join(HERE, "libb200ks.so")
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs += [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include")))]
    if not (force or _newer(out, srcs)):
        return out
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp", ".c"))]
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"),
                                                                           "-o", out] + units
    subprocess.check_call(cmd)
    return out


def build_milc_shim(force=False):
    """Route-1 boundary (MILC-named symbols), double and single MILC_PRECISION builds."""
    outs = []
    src = os.path.join(HERE, "csrc_milc", "milc_shim.c")
    deps = [src, os.path.join(ROOT, "include", "b200ks_milc.h"), os.path.join(ROOT, "include", "b200ks.h")]
    for prec, name in ((2, "libb200ks_milc.so"), (1, "libb200ks_milc_f.so")):
        out = os.path.join(HERE, name)
        if force or _newer(out, deps):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu99", "-Wall", "-DMILC_PRECISION=%d" % prec,
                                   "-I", os.path.join(ROOT, "include"), "-o", out, src,
                                   "-L", HERE, "-lb200ks", "-Wl,-rpath,$ORIGIN"])
        outs.append(out)
    return outs


def build_all(force=False, verbose=False):
    return [build_cuda(force, verbose)] + build_milc_shim(force)


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print("built", p)
