#!/usr/bin/env python
"""K2 diagnostics on the GPU: per-site latency and status words for a batch of identical-shape chains.

    python tools/k2_diag.py --S 32 --T 100000 --chains 1 [--full-sort] [--lam 30] [--pmiss 0.05] [--no-smoothing]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hygeia_b200 import model, synthetic
from hygeia_b200.single_group import Session, make_run_args, STATUS_WORDS

p = argparse.ArgumentParser()
p.add_argument("--S", type=int, default=32); p.add_argument("--T", type=int, default=100000); p.add_argument("--chains", type=int, default=1)
p.add_argument("--lam", type=float, default=30.0); p.add_argument("--pmiss", type=float, default=0.05)
p.add_argument("--full-sort", action="store_true"); p.add_argument("--no-smoothing", action="store_true"); p.add_argument("--reps", type=int, default=2)
p.add_argument("--pe", action="store_true")
a = p.parse_args()
vartheta, _ = model.get_known_parameters(); theta = model.default_theta()
ch = synthetic.make_chain(a.T, a.S, seed=99, lam=a.lam, p_missing=a.pmiss)
s = Session(0)
s.set_vartheta(vartheta); s.set_theta(theta, a.T)
ds = s.add_dataset(ch["n_total"], ch["n_meth"])
outs = [dict(regime_probs=np.full((a.T, 7), np.nan), logz=np.zeros(a.T)) for _ in range(a.chains)]
s.set_chains([dict(dataset=ds, seed=k, chain_id=0, positions=ch["positions"], **o) for k, o in enumerate(outs)])
s.emission()
ra = make_run_args(resample_full_sort=a.full_sort, use_online_marginal_smoothing=not a.no_smoothing, use_online_parameter_estimation=a.pe)
ms = []
for _ in range(a.reps):
    s.filter(ra); s.sync(); ms.append(s.timings()["ms_filter"])
st = s.download()
tot = np.sum(np.array(st), axis=0)
units, seg, ctas = s.filter_units(with_segment_sites=True)
print(json.dumps({"S": a.S, "T": a.T, "chains": a.chains, "full_sort": a.full_sort, "smoothing": not a.no_smoothing, "pe": a.pe, "lam": a.lam,
                  "ms": ms, "us_per_site_per_cta": 1000.0 * min(ms) * min(ctas, a.chains) / (a.T * a.chains) if ctas else None,
                  "ctas": ctas, "status_sum": dict(zip(STATUS_WORDS, [int(x) for x in tot])),
                  "acc": float((outs[0]["regime_probs"][:, 1:].argmax(1) == ch["regimes"]).mean()) if not a.no_smoothing else None}))
