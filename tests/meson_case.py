"""The correlator table shared by tests/golden/make_golden_meson.py and tests/test_meson.py: what a ks_spectrum input
with several sink operators, momenta and reflection parities turns into (ks_spectrum/setup.c builds the same tables)."""
import numpy as np

DIMS = (4, 6, 4, 8)
R0 = [1, 0, 3, 2]
MOM = [[0, 0, 0], [1, 0, 0], [0, 1, 1], [1, 1, 1], [2, 0, 1]]
PAR = [[3, 3, 3], [2, 3, 3], [3, 1, 2], [1, 1, 1], [3, 2, 1]]       # EVENANDODD = 3, EVEN = 2, ODD = 1
LOCAL = ["pion5", "pion05", "rhox", "rhoy", "rhoz", "rhox0", "rhoy0", "rhoz0", "rhoi", "G5-G5", "GXT-GXT", "G1-G1", "GT-GT"]
SHIFTED = ["pioni5", "rhozs", "rhoxsfn", "rhotsfn", "G5X-GY"]         # one-link (APE links), FN currents, gamma-gamma one-link
NPROP = 7


def sources(seed=5):
    rng = np.random.default_rng(seed)
    V = int(np.prod(DIMS))
    return rng.standard_normal((V, 3, 2)), rng.standard_normal((V, 3, 2))


def table(index_of, names):
    """spin_taste, p_index, phase, factor, corr_index, corr_table for three correlators per operator."""
    spin_taste, p_index, phase, factor, corr_index = [], [], [], [], []
    m = 0
    for k, nm in enumerate(names):
        for j in range(3):
            spin_taste.append(index_of[nm])
            p_index.append((k + j) % len(MOM))
            phase.append((k + j) % 4)
            factor.append(0.5 + 0.1 * j + k)
            corr_index.append(m % NPROP)
            m += 1
    corr_table = []
    for c in range(len(spin_taste)):
        if c == 0 or spin_taste[c] != spin_taste[c - 1]:
            corr_table.append([])
        corr_table[-1].append(c)
    return spin_taste, p_index, phase, factor, corr_index, corr_table
