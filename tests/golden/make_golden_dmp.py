"""Generate tests/golden/dmp_small.npz from the REFERENCE's own multiple_testing.py (imported from /root/reference, build
container only) and from the NumPy expressions of aggregate_results.py / get_dmps.py restated in oracle/dmp_oracle.py.

    python tests/golden/make_golden_dmp.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/src/two_group")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import multiple_testing as ref  # noqa: E402  the reference module itself
import dmp_oracle  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    T, P, R = 3000, 200, 6          # 8 seeds x 25 backward trajectories
    # trajectories with structure: stretches of agreement / difference, as the two-group sampler produces
    base = rng.integers(0, R, size=T)
    control = np.repeat(base[:, None], P, 1)
    flip = rng.random((T, P)) < 0.08
    control[flip] = rng.integers(0, R, size=int(flip.sum()))
    case = control.copy()
    dm = np.zeros(T, bool)
    for start in rng.integers(0, T - 60, size=12):
        dm[start:start + int(rng.integers(5, 60))] = True
    diff = dm[:, None] & (rng.random((T, P)) < 0.97)
    case[diff] = (case[diff] + 1 + rng.integers(0, R - 1, size=int(diff.sum()))) % R
    merged = (~diff).astype(np.int8)
    merged[rng.random((T, P)) < 0.02] ^= 1
    st = dmp_oracle.site_statistics(merged, control, case, R, pairs=True)
    positions = 10000 + np.cumsum(1 + rng.geometric(0.01, size=T))
    out = dict(merged=merged.astype(np.int8), control=control.astype(np.int8), case=case.astype(np.int8), positions=positions,
               **{k: v for k, v in st.items()})
    # the reference's FDR procedures on the null statistics (get_dmps.py:111-113,153-155)
    diffs = np.full(T, np.nan)
    pos = positions.astype(np.float64)
    d1 = np.full(T, np.nan); d1[1:] = pos[1:] - pos[:-1]
    d2 = np.full(T, np.nan); d2[2:] = pos[2:] - pos[:-2]
    d3 = np.full(T, np.nan); d3[3:] = pos[3:] - pos[:-3]
    diffs = 1 / 3 * (d1 + d2 + d3)                                           # get_dmps.py:81-82
    w_fn = 1. / np.where(np.isnan(diffs), 1e+5, diffs)                        # get_dmps.py:108
    w_fp = np.ones(T)                                                         # get_dmps.py:102
    out["w_fn"], out["w_fp"] = w_fn, w_fp
    for a in (0.01, 0.05, 0.5):
        k, Qk, thr = ref.FDR_procedure(st["null_stats"], a)
        out[f"fdr_{a}"] = np.array([float(np.squeeze(k)), float(Qk), float(thr)])
        idx, Nk = ref.weighted_FDR_procedure(st["null_stats"], a, w_fp, w_fn)
        out[f"wfdr_idx_{a}"] = np.sort(idx)
        out[f"wfdr_Nk_{a}"] = np.array([Nk])
        # the restatement agrees with the reference module
        k2, Q2, t2 = dmp_oracle.FDR_procedure(st["null_stats"], a)
        assert (int(np.squeeze(k)), float(Qk), float(thr)) == (int(k2), float(Q2), float(t2))
        # weighted procedure: the reference's argsort is unstable, so the selected SET is defined up to ties of the ranking
        i2, N2 = dmp_oracle.weighted_FDR_procedure(st["null_stats"], a, w_fp, w_fn)
        t = st["null_stats"]
        ranking = w_fp * (t - a) / (w_fn * (1 - t) + w_fp * abs(t - a))
        out[f"wfdr_ranking_{a}"] = ranking
        sym = np.setxor1d(i2, idx)
        assert len(i2) == len(idx) and abs(N2 - Nk) < 1e-9 * max(1.0, abs(Nk)), (len(i2), len(idx), N2, Nk)
        if len(sym):
            assert np.ptp(ranking[sym]) == 0.0 and ranking[sym][0] == ranking[idx].max(), "sets differ beyond ties"
    np.savez_compressed(os.path.join(HERE, "dmp_small.npz"), **out)
    print("wrote dmp_small.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
