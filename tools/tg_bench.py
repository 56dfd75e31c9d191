"""K4/K5 (two-group filter + backward simulation): microseconds per site and chain with every SM busy (two resident CTAs per SM).

    python tools/tg_bench.py [--sites 4000] [--chains 148] [--samples 8]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=4000)
    ap.add_argument("--chains", type=int, default=296, help="default: two resident CTAs on each of the 148 SMs")
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--backward", type=int, default=25)
    a = ap.parse_args()
    from _tg_case import make_case
    from hygeia_b200.two_group import TwoGroupSession
    T = a.sites
    c = make_case(T, a.samples, seed=5, d_max=T + 10)
    s = TwoGroupSession(0)
    s.set_emission_model(c["mu"], c["sigma"], c["u"])
    s.add_dataset(c["nt_c"], c["nm_c"])
    s.add_dataset(c["nt_k"], c["nm_k"])
    s.emission()
    s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], c["u"], 50, a.backward, rho_control=c["model"].rho_c,
                          rho_case=c["model"].rho_k, t_max=T)
    specs = [dict(control_dataset=0, case_dataset=1, T=T, seed=k, chain_id=k) for k in range(a.chains)]
    ms = []
    for _ in range(1 + a.reps):
        out = s.run(specs)
        ms.append(s.ms_two_group)
    ms = ms[1:]
    avg = float(np.mean(ms))
    print(json.dumps({"kernel": "tg_kernel (K4/K5)", "sites": T, "chains": a.chains, "ms": ms,
                      "us_per_site_per_chain": 1000.0 * avg / T / max(1, -(-a.chains // 296)),
                      "us_per_site_per_sm": 1000.0 * avg * 148 / (T * a.chains),
                      "site_chains_per_s": T * a.chains / avg * 1e3,
                      "log_norm_chain0": float(out[0]["log_normalizing_constant"])}))


if __name__ == "__main__":
    main()
