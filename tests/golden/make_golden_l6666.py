#!/usr/bin/env python
"""Generates tests/golden/ks_spectrum_hisq.nd.l6666.corrfile.golden: the reference's own CPU build of
ks_spectrum_hisq (oracle/_ref/apps/ks_spectrum_hisq_cpu, built from /root/reference by oracle/build_apps.sh) run on
the 6^4 variant of its nd.2 sample input with binary_samples/lat.sample.l6666.hisq -- BASELINE configs[0].
    python tests/golden/make_golden_l6666.py
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import l6666  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    corr, out = l6666.run("ks_spectrum_hisq_cpu", d)
with open(l6666.GOLDEN, "w") as f:
    f.write("\n".join(corr) + "\n")
print("wrote %s: %d lines; CPU solves: %d" % (l6666.GOLDEN, len(corr), sum("CONGRAD5" in ln for ln in out)))
