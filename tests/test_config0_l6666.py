"""BASELINE configs[0] literally (see tests/l6666.py): the CPU build pins the committed golden; the same
application linked against libb200ks (every solve through the GPU seam) must reproduce it -- on one GPU and,
behind the unchanged call surface, on two and four."""
import os

import pytest

import l6666


def _golden():
    return open(l6666.GOLDEN).read().splitlines()


def _have(app):
    return os.path.exists(os.path.join(l6666.APPS, app)) and os.path.isdir(l6666.SAMPLES)


def test_reference_cpu_build_reproduces_the_committed_l6666_golden(tmp_path):
    if not _have("ks_spectrum_hisq_cpu"):
        pytest.skip("oracle/_ref/apps not built (oracle/build_apps.sh needs /root/reference)")
    corr, out = l6666.run("ks_spectrum_hisq_cpu", tmp_path)
    bad, worst = l6666.compare(corr, _golden(), rel=1e-9)
    assert not bad, "\n".join(bad[:20])
    assert len(corr) > 500 and any("mass 0.05" in ln or "0.05" in ln for ln in out)


@pytest.mark.gpu
@pytest.mark.parametrize("app", ["ks_spectrum_hisq_b200", "ks_spectrum_hisq_b200fl"])
def test_l6666_hisq_spectrum_on_libb200ks_matches_the_cpu_golden(app, tmp_path):
    if not _have(app):
        pytest.skip("oracle/_ref/apps not built")
    corr, out = l6666.run(app, tmp_path)
    bad, worst = l6666.compare(corr, _golden())
    print("worst relative difference to the CPU golden: %.2e" % worst)
    assert not bad, "\n".join(bad[:20])
    assert any("fn_QUDA" in ln or "multicg_offset_QUDA" in ln for ln in out), "solves did not go through the GPU seam"
