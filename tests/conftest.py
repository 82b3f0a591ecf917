import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


_FIELD_CACHE = {}


def fields_for(dims, seed=1234):
    """Seeded synthetic HISQ-like links + an all-parity source for a lattice size."""
    from milc_qcd_b200 import fields as F
    key = (tuple(dims), seed)
    if key not in _FIELD_CACHE:
        fat, lng = F.make_links(dims, seed=seed)
        src = F.make_source(dims, seed=seed + 1, parity=F.EVENANDODD)
        _FIELD_CACHE[key] = (fat, lng, src)
    return _FIELD_CACHE[key]
