"""Timing of one chain in parameter-estimation mode (K3) vs plain filtering, for a sweep of chain lengths (tuning aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hygeia_b200 import model, synthetic  # noqa: E402
from hygeia_b200.single_group import Session, make_run_args  # noqa: E402

S = 8
vartheta, _ = model.get_known_parameters()
theta = model.default_theta()
s = Session(0)
s.set_vartheta(vartheta)
for T in [int(x) for x in sys.argv[1:]] or [2000, 8000]:
    for pe in (False, True):
        s.clear()
        s.set_theta(theta, T)
        ch = synthetic.make_chain(T, S, seed=7)
        s.add_dataset(ch["n_total"], ch["n_meth"])
        tr = np.zeros((T, 36))
        pr = np.zeros((T, 7))
        s.set_chains([dict(dataset=0, seed=1, chain_id=0, regime_probs=pr, theta_trace=tr if pe else None)])
        s.emission()
        for rep in range(2):
            s.filter(make_run_args(use_online_parameter_estimation=pe))
            s.sync()
        tm = s.timings()
        s.download()
        print(f"T={T} pe={pe}: {tm['ms_filter']:.1f} ms  {tm['ms_filter']*1e3/T:.2f} us/site", flush=True)
s.close()
