"""DMP calling on the GPU (K6 + device sort/scan) through the C ABI, against the fixture generated from the reference's
multiple_testing.py and against the NumPy oracle.  Integer counts -> bit-exact statistics."""
import os
import sys

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_site_statistics_bit_exact_vs_golden(built):
    from hygeia_b200 import dmp
    g = golden("dmp_small.npz")
    r = dmp.site_statistics(g["merged"], g["control"], g["case"], 6, test_regime_combinations=True)
    assert np.array_equal(r["split_probs"], g["split_probs"])
    assert np.array_equal(r["null_stats"], g["null_stats"])
    assert np.array_equal(r["control_freqs"], g["control_freqs"])
    assert np.array_equal(r["case_freqs"], g["case_freqs"])
    off = ~np.eye(6, dtype=bool)
    assert np.array_equal(r["pair_stats"][:, off], g["pair_stats"][:, off])


@pytest.mark.parametrize("T,P,R", [(1, 1, 2), (129, 25, 6), (1000, 50, 6), (257, 203, 6), (300, 580, 8), (300, 2000, 6), (5, 7, 3), (640, 200, 6), (333, 250, 6), (100, 75, 6), (97, 1, 6), (64, 3, 2), (40, 254, 6)])
def test_site_statistics_shapes_and_ragged_rows(built, T, P, R):
    """Row lengths that are not multiples of 4 / 16, tiles that end inside a 16-byte vector, more than 252 particles (count
    fields are flushed), a single site."""
    import dmp_oracle
    from hygeia_b200 import dmp
    rng = np.random.default_rng(T * 1000 + P)
    merged = rng.integers(0, 2, size=(T, P)); control = rng.integers(0, R, size=(T, P)); case = rng.integers(0, R, size=(T, P))
    want = dmp_oracle.site_statistics(merged, control, case, R, pairs=True)
    got = dmp.site_statistics(merged, control, case, R, test_regime_combinations=True)
    for k in ("split_probs", "null_stats", "control_freqs", "case_freqs", "pair_stats"):
        assert np.array_equal(got[k], want[k]), k


def test_fdr_procedures_vs_golden(built):
    from hygeia_b200 import dmp
    g = golden("dmp_small.npz")
    for a in (0.01, 0.05, 0.5):
        k, Qk, thr = dmp.FDR_procedure(g["null_stats"], a)
        wk, wQ, wt = g[f"fdr_{a}"]
        assert k == int(wk) and thr == wt and abs(Qk - wQ) <= 1e-12 * max(1.0, abs(wQ))
        idx, Nk = dmp.weighted_FDR_procedure(g["null_stats"], a, g["w_fp"], g["w_fn"])
        want = g[f"wfdr_idx_{a}"]
        assert len(idx) == len(want) and abs(Nk - g[f"wfdr_Nk_{a}"][0]) < 1e-9
        sym = np.setxor1d(idx, want)
        assert len(sym) == 0 or np.ptp(g[f"wfdr_ranking_{a}"][sym]) == 0.0


def test_fdr_edge_cases_vs_oracle(built):
    import dmp_oracle
    from hygeia_b200 import dmp
    rng = np.random.default_rng(9)
    t = rng.random(5000)
    assert dmp.FDR_procedure(t, 1e-9) == (0, 0.0, 0.0)                      # threshold below the smallest statistic
    k, Qk, thr = dmp.FDR_procedure(t, 5.0)                                   # everything selected
    assert k == 5000 and thr == 1.01 and abs(Qk - t.mean()) < 1e-12
    k, Qk, thr = dmp.FDR_procedure(np.array([0.25]), 0.3)
    assert (k, Qk, thr) == (1, 0.25, 1.01)
    for a in (0.05, 0.2):
        k, Qk, thr = dmp.FDR_procedure(t, a)
        k2, Q2, t2 = dmp_oracle.FDR_procedure(t, a)
        assert k == int(k2) and thr == float(t2) and abs(Qk - Q2) < 1e-12
        w = 0.5 + rng.random(5000)
        idx, Nk = dmp.weighted_FDR_procedure(t, a, np.ones(5000), w)
        i2, N2 = dmp_oracle.weighted_FDR_procedure(t, a, np.ones(5000), w)
        assert np.array_equal(idx, i2) and abs(Nk - N2) < 1e-12               # no ties: same ORDER, not only the same set
    idx, Nk = dmp.weighted_FDR_procedure(np.array([0.9, 0.8]), 0.05, np.ones(2), np.ones(2))
    i2, N2 = dmp_oracle.weighted_FDR_procedure(np.array([0.9, 0.8]), 0.05, np.ones(2), np.ones(2))
    assert len(idx) == 0 and len(i2) == 0 and abs(Nk - N2) < 1e-12          # nothing selected: Nsums[-1], as the reference


def test_device_sort_scan_count_at_size_with_ties_and_signs(built):
    """The procedures' sort / scan / count are this repo's kernels (stable LSD radix sort on the order-preserving image of the
    doubles, three-kernel scan): 300 000 statistics from a set of 76 values -- thousands of exact ties -- with ranking values of both
    signs; the selected indices must be those of a STABLE argsort, in order."""
    from hygeia_b200 import dmp
    rng = np.random.default_rng(12)
    n = 300_000
    t = rng.integers(0, 76, size=n) / 75.0
    wfn = rng.integers(1, 4, size=n).astype(np.float64)
    wfp = np.ones(n)
    a = 0.3
    ranking = wfp * (t - a) / (wfn * (1 - t) + wfp * np.abs(t - a))
    order = np.argsort(ranking, kind="stable")
    nsums = np.cumsum((wfp * (t - a))[order])
    s = int(np.sum(nsums <= 0))
    idx, Nk = dmp.weighted_FDR_procedure(t, a, wfp, wfn)
    assert (ranking < 0).any() and (ranking > 0).any() and s > 1000
    assert len(idx) == s or abs(nsums[min(len(idx), n - 1)]) < 1e-6      # the count may differ only where a partial sum is 0 up to rounding
    m = min(s, len(idx))
    assert np.array_equal(idx[:m], order[:m])
    assert abs(Nk - nsums[len(idx) - 1]) < 1e-7
    k, Qk, thr = dmp.FDR_procedure(t, 0.2)
    st = np.sort(t)
    qs = np.cumsum(st) / np.arange(1, n + 1)
    k2 = int(np.sum(qs <= 0.2))
    assert abs(k - k2) <= 2 and abs(Qk - qs[k - 1]) < 1e-9 and thr == st[k]
