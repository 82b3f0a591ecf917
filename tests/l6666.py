"""BASELINE configs[0] as written: HISQ single-mass CG solves (masses 0.03, 0.05 and 0.07 with a Naik epsilon) on
the shipped 6^4 lattice binary_samples/lat.sample.l6666.hisq, "as in the ks_spectrum sample test".  The reference
ships no golden for this combination (its ks_spectrum samples use l8888), so -- SURVEY.md section 8c -- the input is the
6^4 variant of ks_spectrum/test/ks_spectrum_hisq.nd.2.sample-in and the golden is the reference's own CPU build run
on it (tests/golden/make_golden_l6666.py, committed output).  Test infrastructure shared by the CPU and GPU tests."""
import os
import shutil
import subprocess

from conftest import ROOT
import milc_regress as R

REF = os.path.join(ROOT, "oracle", "_ref")
APPS = os.path.join(REF, "apps")
SAMPLES = os.path.join(REF, "samples")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ks_spectrum_hisq.nd.l6666.corrfile.golden")


def make_input():
    src = open(os.path.join(SAMPLES, "ks_spectrum", "test", "ks_spectrum_hisq.nd.2.sample-in")).read()
    for d in ("nx", "ny", "nz", "nt"):
        src = src.replace("\n%s 8\n" % d, "\n%s 6\n" % d)
    assert "lat.sample.l8888" in src
    src = src.replace("lat.sample.l8888", "lat.sample.l6666.hisq")
    return src.replace("ks_spectrum_hisq.nd.2.corrfile_t0.test-out", "corr.l6666.test-out")


def run(app, workdir, env=None):
    work = os.path.join(str(workdir), "ks_spectrum", "test")
    os.makedirs(work, exist_ok=True)
    bs = os.path.join(str(workdir), "binary_samples")
    if not os.path.exists(bs):
        os.symlink(os.path.join(SAMPLES, "binary_samples"), bs)
    out = os.path.join(work, "corr.l6666.test-out")
    if os.path.exists(out):
        os.remove(out)
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(APPS, app)], input=make_input(), capture_output=True, text=True, cwd=work, timeout=900, env=e)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    corr = [ln for ln in open(out).read().splitlines() if not ln.startswith("date:") and not ln.startswith("JobID:")]
    return corr, p.stdout.splitlines()


def compare(got, want, rel=2e-5, floor=1e-9):
    """Field by field like diffn3.pl; numeric fields to the relative size of the reference's own corrfile tolerances
    (2e-5 on O(1) correlators in ks_spectrum_hisq.nd.2.corrfile_t0.errtol).  Returns (discrepancies, worst rel diff)."""
    bad, worst = [], 0.0
    if len(got) != len(want):
        return ["line counts differ: %d vs %d" % (len(got), len(want))], float("inf")
    for n, (a, b) in enumerate(zip(got, want)):
        fa, fb = a.split(), b.split()
        if len(fa) != len(fb):
            bad.append("line %d: %r vs %r" % (n + 1, a, b))
            continue
        for x, y in zip(fa, fb):
            if R._is_number(x) and R._is_number(y):
                vx, vy = R._val(x), R._val(y)
                d = abs(vx - vy)
                if max(abs(vx), abs(vy)) > 1e-6:
                    worst = max(worst, d / max(abs(vx), abs(vy)))
                if d > rel * max(abs(vx), abs(vy)) + floor:
                    bad.append("line %d: %s vs %s" % (n + 1, x, y))
            elif x != y:
                bad.append("line %d: %r vs %r" % (n + 1, x, y))
    return bad, worst
