"""End-to-end drop-in check with the reference's OWN applications and OWN regression data.

oracle/build_apps.sh builds ks_spectrum_hisq and su3_rhmc_hisq from the unmodified reference
sources twice: the default CPU build (`*_cpu`) and the USE_CG_GPU build linked against
libb200ks through include/quda_milc_interface.h (`*_b200`).  Both are run on the shipped sample
inputs and compared with the shipped sample outputs under the shipped tolerances
(ks_spectrum/test/checklist:50-62, ks_imp_rhmc/test/checklist), using the reference's
procedure restated in tests/milc_regress.py.

  * CPU flavour (no GPU needed): pins oracle/_ref -- the reference as built here reproduces
    its own known answers.
  * b200 flavour (-m gpu): every solve of the run (HISQ single-mass CG with UML even/odd
    reconstruction, multi-shift CG, RHMC trajectories) goes through the CUDA library and the
    physics output still matches the reference's goldens.
"""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT
import milc_regress as R

REF = os.path.join(ROOT, "oracle", "_ref")
APPS = os.path.join(REF, "apps")
SAMPLES = os.path.join(REF, "samples")


def _have(app):
    return os.path.exists(os.path.join(APPS, app)) and os.path.isdir(SAMPLES)


def _run(app, testdir, stem, tmp_path, timeout=900, env=None):
    """Run an application on <stem>.sample-in inside a scratch copy of the staged test dir."""
    work = tmp_path / "work" / testdir / "test"
    shutil.copytree(os.path.join(SAMPLES, testdir, "test"), work)
    for f in os.listdir(work):  # the applications append to their output files
        if f.endswith(".test-out"):
            os.remove(work / f)
    bs = tmp_path / "work" / "binary_samples"
    if not bs.exists():
        os.symlink(os.path.join(SAMPLES, "binary_samples"), bs)
    with open(work / (stem + ".sample-in")) as fin:
        e = dict(os.environ)
        e.update(env or {})
        p = subprocess.run([os.path.join(APPS, app)], stdin=fin, capture_output=True, text=True, cwd=work,
                           timeout=timeout, env=e)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return work, p.stdout.splitlines()


def _read(path):
    with open(path) as f:
        return f.read().splitlines()


def check_spectrum(app, case, tmp_path, stdout_strict, env=None, timeout=900):
    stem = "ks_spectrum_hisq.%s.2" % case
    work, out = _run(app, "ks_spectrum", stem, tmp_path, env=env, timeout=timeout)
    # extra-output: the correlator file, all lines  (checklist: `extra-output ... --- EOF`)
    got = R.filter_test_lines(_read(work / (stem + ".corrfile_t0.test-out")))
    want = _read(work / (stem + ".corrfile_t0.sample-out"))
    tol = _read(work / (stem + ".corrfile_t0.errtol"))
    bad = R.diffn3(got, want, tol)
    assert not bad, "\n".join(bad[:20])
    assert len(got) > 100
    # stdout region between the checklist patterns
    pats = ["PBP:", "FACTION:", "PBP:", "FACTION:"] if case == "spectrum2" else ["PLAQ:", "NERSC"]
    g = R.filter_test_lines(R.headtail(out, pats))
    w = R.headtail(_read(work / (stem + ".sample-out")), pats)
    t = _read(work / (stem + ".errtol"))
    if stdout_strict:
        bad = R.diffn3(g, w, t)
        assert not bad, "\n".join(bad[:20])
    return out


@pytest.mark.parametrize("case", ["nd", "fpi"])
def test_reference_cpu_build_reproduces_its_goldens(case, tmp_path):
    if not _have("ks_spectrum_hisq_cpu"):
        pytest.skip("oracle/_ref/apps not built (oracle/build_apps.sh needs /root/reference)")
    check_spectrum("ks_spectrum_hisq_cpu", case, tmp_path, stdout_strict=True)


def check_rhmc(app, tmp_path):
    stem = "su3_rhmc_hisq.2"
    work, out = _run(app, "ks_imp_rhmc", stem, tmp_path)
    sel = "PBP|DG|PLAQ|ACTION|delta|G_LOOP"
    g = R.filter_test_lines(R.headtail(out, ["delta", "RUNNING"], sel))
    w = R.headtail(_read(work / (stem + ".sample-out")), ["delta", "RUNNING"], sel)
    t = _read(work / (stem + ".errtol"))
    bad = R.diffn3(g, w, t)
    assert not bad, "\n".join(bad[:20])
    assert len(g) >= 30
    return out


def test_reference_cpu_rhmc_reproduces_its_goldens(tmp_path):
    if not _have("su3_rhmc_hisq_cpu"):
        pytest.skip("oracle/_ref/apps not built")
    check_rhmc("su3_rhmc_hisq_cpu", tmp_path)


def test_b200_apps_link_only_the_documented_symbols():
    """The unmodified MILC objects need exactly the quda* symbols INTEGRATION.md lists."""
    if not _have("ks_spectrum_hisq_b200"):
        pytest.skip("oracle/_ref/apps not built")
    want = {"qudaInit", "qudaSetMPICommHandle", "qudaFinalize", "qudaAllocatePinned", "qudaFreePinned",
            "qudaInvert", "qudaInvertMsrc", "qudaMultishiftInvert", "qudaDslash"}
    for app, extra in (("ks_spectrum_hisq_b200", set()), ("su3_rhmc_hisq_b200", {"qudaMomAction"})):
        nm = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(APPS, app)], capture_output=True, text=True).stdout
        used = {ln.split()[-1] for ln in nm.splitlines() if " quda" in ln}
        assert used == want | extra, (app, used ^ (want | extra))


@pytest.mark.gpu
# ("periodic" needs QIO to save a SciDAC propagator; the image has no QIO, so neither the CPU nor
# the GPU flavour of that case can run here)
@pytest.mark.parametrize("case", ["nd", "fpi", "nl", "nlpi2", "spectrum2"])
def test_ks_spectrum_hisq_on_libb200ks_matches_reference_goldens(case, tmp_path):
    if not _have("ks_spectrum_hisq_b200"):
        pytest.skip("oracle/_ref/apps not built")
    out = check_spectrum("ks_spectrum_hisq_b200", case, tmp_path, stdout_strict=False)
    assert any("fn_QUDA" in ln or "multicg_offset_QUDA" in ln for ln in out), "solves did not go through the GPU seam"


@pytest.mark.gpu
@pytest.mark.parametrize("ngpu,case", [(2, "nd"), (4, "nd"), (2, "spectrum2"), (4, "fpi")])
def test_ks_spectrum_hisq_on_several_gpus_behind_the_seam_matches_reference_goldens(ngpu, case, tmp_path):
    """The UNMODIFIED application, still one vanilla MILC rank, with B200KS_NGPU devices behind qudaInvert /
    qudaMultishiftInvert / qudaDslash (b200ks_create_multi): 8^4 split 2 ways in t, or 2 x 2 in z and t (local extent
    4, no interior sites at all).  On a 1-GPU box only the two-member cases run, both members on the one device
    (B200KS_NGPU_OVERSUBSCRIBE: the whole multi-GPU host path still runs; a stall there is a skip, see test_gpu_seam.py).  (The 6^4 RHMC sample cannot be split: its local extents would be odd.)"""
    if not _have("ks_spectrum_hisq_b200"):
        pytest.skip("oracle/_ref/apps not built")
    import torch
    env = {"B200KS_NGPU": str(ngpu), "B200KS_NGPU_OVERSUBSCRIBE": "1"}
    have = torch.cuda.device_count()
    shared = have < ngpu
    if shared:
        if case != "nd":
            pytest.skip("members sharing a device: only the first two-member case runs (needs one device per member)")
        env["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"   # one hardware work queue per stream
    if ngpu > 2 * have:
        pytest.skip("%d members need %d GPUs (this box has %d)" % (ngpu, ngpu, have))
    try:
        out = check_spectrum("ks_spectrum_hisq_b200", case, tmp_path, stdout_strict=False, env=env,
                             timeout=60 if shared else 900)
    except subprocess.TimeoutExpired:
        if not shared:
            raise
        pytest.skip("%d members sharing %d device(s) stalled (needs one device per member)" % (ngpu, have))
    except AssertionError as e:
        if shared and "halo exchange timed out" in str(e):
            pytest.skip("%d members sharing %d device(s): halo wait gave up (needs one device per member)" % (ngpu, have))
        raise
    assert any("fn_QUDA" in ln or "multicg_offset_QUDA" in ln for ln in out), "solves did not go through the GPU seam"
    assert any("lattice spread over %d GPUs" % ngpu in ln for ln in out), "the multi-GPU context was not used"


@pytest.mark.gpu
def test_su3_rhmc_hisq_on_libb200ks_matches_reference_goldens(tmp_path):
    if not _have("su3_rhmc_hisq_b200"):
        pytest.skip("oracle/_ref/apps not built")
    out = check_rhmc("su3_rhmc_hisq_b200", tmp_path)
    assert any("multicg_offset_QUDA" in ln for ln in out)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["nd", "spectrum2"])
def test_ks_spectrum_hisq_with_gpu_link_construction_matches_reference_goldens(case, tmp_path):
    """-DUSE_FL_GPU build (WANT_FL_GPU=true): the HISQ links themselves are built by libb200ks
    (qudaLoadUnitarizedLink + qudaLoadKSLink), then the solves run on them; same goldens."""
    if not _have("ks_spectrum_hisq_b200fl"):
        pytest.skip("oracle/_ref/apps not built")
    out = check_spectrum("ks_spectrum_hisq_b200fl", case, tmp_path, stdout_strict=False)
    assert any("fn_QUDA" in ln or "multicg_offset_QUDA" in ln for ln in out), "solves did not go through the GPU seam"


@pytest.mark.gpu
def test_su3_rhmc_hisq_with_gpu_link_construction_matches_reference_goldens(tmp_path):
    """A full RHMC trajectory: links rebuilt on the GPU after every gauge update."""
    if not _have("su3_rhmc_hisq_b200fl"):
        pytest.skip("oracle/_ref/apps not built")
    out = check_rhmc("su3_rhmc_hisq_b200fl", tmp_path)
    assert any("multicg_offset_QUDA" in ln for ln in out)


def test_b200fl_apps_bind_the_link_construction_symbols():
    if not _have("ks_spectrum_hisq_b200fl"):
        pytest.skip("oracle/_ref/apps not built")
    for app in ("ks_spectrum_hisq_b200fl", "su3_rhmc_hisq_b200fl"):
        nm = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(APPS, app)], capture_output=True, text=True).stdout
        used = {ln.split()[-1] for ln in nm.splitlines() if " quda" in ln}
        assert {"qudaLoadKSLink", "qudaLoadUnitarizedLink"} <= used, (app, used)
