"""DMP calling, CPU side: the NumPy restatement (oracle/dmp_oracle.py) against the committed fixture generated from the
reference's own multiple_testing.py, and -- when /root/reference is present (build container) -- against that module live."""
import os
import sys

import numpy as np
import pytest

from conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dmp_oracle  # noqa: E402

REF_DIR = "/root/reference/src/two_group"


def test_oracle_matches_golden_fixture():
    g = golden("dmp_small.npz")
    st = dmp_oracle.site_statistics(g["merged"], g["control"], g["case"], 6, pairs=True)
    for k in ("split_probs", "null_stats", "control_freqs", "case_freqs", "pair_stats"):
        assert np.array_equal(st[k], g[k]), k
    for a in (0.01, 0.05, 0.5):
        k, Qk, thr = dmp_oracle.FDR_procedure(g["null_stats"], a)
        assert [float(k), float(Qk), float(thr)] == list(g[f"fdr_{a}"])
        idx, Nk = dmp_oracle.weighted_FDR_procedure(g["null_stats"], a, g["w_fp"], g["w_fn"])
        want = g[f"wfdr_idx_{a}"]
        assert len(idx) == len(want) and abs(Nk - g[f"wfdr_Nk_{a}"][0]) < 1e-9
        sym = np.setxor1d(idx, want)                      # the reference's argsort is unstable: equal up to ties of the ranking
        assert len(sym) == 0 or np.ptp(g[f"wfdr_ranking_{a}"][sym]) == 0.0


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "multiple_testing.py")), reason="reference tree not present")
def test_oracle_matches_reference_module_live():
    sys.path.insert(0, REF_DIR)
    import multiple_testing as ref
    rng = np.random.default_rng(3)
    for n in (1, 7, 1000):
        t = np.round(rng.random(n), 3)                    # ties on purpose
        for a in (0.0005, 0.05, 0.3, 2.0):
            k, Qk, thr = ref.FDR_procedure(t, a)
            k2, Q2, t2 = dmp_oracle.FDR_procedure(t, a)
            assert int(np.squeeze(k)) == int(k2) and float(Qk) == float(Q2) and float(thr) == float(t2)
        w = 0.5 + rng.random(n)
        t = rng.random(n)                                  # no ties: the selected set is unique
        for a in (0.05, 0.3):
            i1, N1 = ref.weighted_FDR_procedure(t, a, np.ones(n), w)
            i2, N2 = dmp_oracle.weighted_FDR_procedure(t, a, np.ones(n), w)
            assert np.array_equal(i1, i2) and N1 == N2
