"""How much does cutting a chain into halo'ed segments change the single-group results?  (CPU study, oracle only.)

A segment run is the standard algorithm started from the R-particle initial system at site t0 - halo_left (uniforms still
indexed by the global site) and stopped once every site below t1 is finalised.  Compared with the whole-chain run of the same
oracle on: posterior rows (max abs), regime calls, log Z increments.

    python tools/segment_study.py --T 120000 --S 4 --segment 20000 --halo 2000
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _oracle import Oracle  # noqa: E402
from hygeia_b200 import model, philox, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=120000)
    ap.add_argument("--S", type=int, default=4)
    ap.add_argument("--lam", type=float, default=30.0)
    ap.add_argument("--segment", type=int, default=20000)
    ap.add_argument("--halo", type=int, nargs="+", default=[500, 2000, 5000])
    ap.add_argument("--right", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=11)
    a = ap.parse_args()

    ch = synthetic.make_chain(a.T, a.S, seed=a.seed, lam=a.lam)
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    o = Oracle()
    alpha, beta = model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA)
    lo = o.emission(alpha, beta, ch["n_total"], ch["n_meth"])
    u = philox.uniforms_by_site(3, 0, a.T)
    full = o.run(vartheta, theta, u, logobs=lo)
    print(f"whole chain: T={a.T} S={a.S} {full['seconds']:.1f} s  max pending {full['n_pending'].max()}  "
          f"max lag {(full['finalised_at'] - np.arange(a.T)).max()}")
    P = full["regime_probs"][:, 1:]
    for H in a.halo:
        worst, calls_diff, worst_dz, forced = 0.0, 0, 0.0, 0
        right_used = []
        for t0 in range(0, a.T, a.segment):
            t1 = min(a.T, t0 + a.segment)
            b = max(0, t0 - H)
            e = min(a.T, t1 + a.right)
            seg = o.run(vartheta, theta, u[b:], logobs=lo[b:e])
            fin = seg["finalised_at"][t0 - b:t1 - b] + b        # global step at which the owned sites were finalised
            # the run could have stopped at the last of those steps
            right_used.append(int(max(0, fin.max() - (t1 - 1))))
            if e < a.T:
                forced += int((fin >= e - 1).sum())
            p = seg["regime_probs"][t0 - b:t1 - b, 1:]
            d = np.abs(p - P[t0:t1]).max()
            worst = max(worst, d)
            calls_diff += int((p.argmax(1) != P[t0:t1].argmax(1)).sum())
            inc = seg["logz"][t0 - b:t1 - b] - (seg["logz"][t0 - b - 1] if t0 > b else 0.0)
            inc_full = full["logz"][t0:t1] - (full["logz"][t0 - 1] if t0 > 0 else 0.0)
            if t0 > 0:
                worst_dz = max(worst_dz, np.abs(inc - inc_full).max())
            same_fin = float((fin == full["finalised_at"][t0:t1]).mean())
        print(f"halo {H:6d}: max |dp| = {worst:.3e}   differing regime calls = {calls_diff}   max |d logZ increment| = {worst_dz:.3e}   "
              f"forced = {forced}   right halo used: max {max(right_used)} mean {np.mean(right_used):.0f}   last-seg same finalisation step {same_fin:.4f}")


if __name__ == "__main__":
    main()
