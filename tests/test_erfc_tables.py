"""The piecewise-polynomial erfc tables of the CUDA kernels (graspa_b200/csrc/erfc_table.inc: degree 12, arguments up to 6.06, shared by
every kernel; erfc_table10.inc: degree 10 on [0, 3.4375), the short table of k_wc_energy_lt), evaluated on the CPU exactly as the kernels
do (interval = nearest multiple of 1/8, Horner in double) against 50-digit erfc.  The reference calls the C library's erfc
(maths.cuh:496-500 CoulombReal); the tolerance BASELINE.json states for energies is 1e-10 relative."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, suffix):
    text = open(os.path.join(ROOT, "graspa_b200", "csrc", name)).read()
    deg = int(re.search(rf"#define GBK_ERFC{suffix}_DEG (\d+)", text).group(1))
    nint = int(re.search(rf"#define GBK_ERFC{suffix}_NINT (\d+)", text).group(1))
    xmax = float(re.search(rf"#define GBK_ERFC{suffix}_XMAX ([0-9.eE+-]+)", text).group(1))
    body = text.split("{", 1)[1].rsplit("}", 1)[0]
    vals = np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()])
    assert vals.size == (deg + 1) * nint
    return deg, nint, xmax, vals.reshape(deg + 1, nint)


@pytest.mark.parametrize("name,suffix,tol", [("erfc_table.inc", "", 1e-15), ("erfc_table10.inc", "10", 6e-16)])
def test_erfc_table_against_50_digit_erfc(name, suffix, tol):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    deg, nint, xmax, coef = _load(name, suffix)
    assert abs(xmax - ((nint - 1) / 8.0 + 1.0 / 16.0)) < 1e-6
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.random(3000) * xmax, np.linspace(0.0, xmax, 1001), [xmax]])
    y = xs * 8.0
    k = np.rint(y); t = y - k; ki = k.astype(int)
    assert ki.max() <= nint - 1
    acc = coef[deg, ki].copy()
    for j in range(deg - 1, -1, -1):
        acc = acc * t + coef[j, ki]
    worst = 0.0
    for x, a in zip(xs, acc):
        ref = mp.erfc(mp.mpf(float(x)))
        worst = max(worst, float(abs((mp.mpf(float(a)) - ref) / ref)))
    assert worst < tol, worst


def test_short_table_covers_the_default_ewald_setup():
    """The reference's Ewald set-up (read_data.cpp:693-697): tol = sqrt|ln(precision r_cut)|, alpha = sqrt|ln(precision r_cut tol)| / r_cut.
    With the examples' `EwaldPrecision 1e-6` alpha * r_cut is 3.1-3.2 for cutoffs of 10-16 A -- inside the short table; a precision of
    1e-8 gives 3.8 and falls back to the full table (engine.cu: P.erfc10_ok)."""
    _, _, xmax10, _ = _load("erfc_table10.inc", "10")
    _, _, xmax12, _ = _load("erfc_table.inc", "")

    def alpha_rc(precision, rc):
        tol = np.sqrt(abs(np.log(precision * rc)))
        return np.sqrt(abs(np.log(precision * rc * tol)))

    for rc in (10.0, 12.0, 12.8, 14.0, 16.0):
        assert alpha_rc(1e-6, rc) < xmax10 < xmax12
        assert xmax10 < alpha_rc(1e-8, rc) < xmax12
