"""One small pass over every kernel (K1 emission, K2 recursion, K3 parameter mode, K4/K5 two-group) for ncu captures:

    python tools/profile_all.py                      # must exit 0 without ncu first
    ncu --set full --clock-control none --import-source on -k regex:"sg_emission|sg_filter|tg_kernel" -o gpurun_out/x python tools/profile_all.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _tg_case import make_case  # noqa: E402
from hygeia_b200 import model, synthetic  # noqa: E402
from hygeia_b200.single_group import Session, make_run_args  # noqa: E402
from hygeia_b200.two_group import TwoGroupSession  # noqa: E402

S = 32
lens = synthetic.chromosome_lengths(int(os.environ.get("HYG_PROFILE_SITES", "3000000")))
vartheta, _ = model.get_known_parameters()
theta = model.default_theta()
s = Session(0)
s.set_vartheta(vartheta)
s.set_theta(theta, int(max(lens)))
probs = []
for i, T in enumerate(lens):
    ch = synthetic.make_chain(int(T), S, seed=100 + i)
    s.add_dataset(ch["n_total"], ch["n_meth"])
    probs.append(np.zeros((int(T), 7)))
s.set_chains([dict(dataset=i, seed=1, chain_id=i, regime_probs=probs[i]) for i in range(len(lens))])
s.emission()
s.filter(make_run_args())
s.download()
print("single-group", s.timings())
# K3: one chain in parameter-estimation mode
s.clear()
ch = synthetic.make_chain(20000, S, seed=7)
s.add_dataset(ch["n_total"], ch["n_meth"])
tr = np.zeros((20000, 36))
s.set_chains([dict(dataset=0, seed=1, chain_id=0, theta_trace=tr)])
s.emission()
s.filter(make_run_args(use_online_parameter_estimation=True))
s.download()
print("parameter mode", s.timings())
s.close()

c = make_case(3000, 50, seed=5, d_max=4096)
t = TwoGroupSession(0)
t.set_emission_model(c["mu"], c["sigma"], 3)
t.add_dataset(c["nt_c"], c["nm_c"])
t.add_dataset(c["nt_k"], c["nm_k"])
t.emission()
t.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], 3, 50, 25, t_max=3000)
out = t.run([dict(control_dataset=0, case_dataset=1, T=3000, seed=1, chain_id=i) for i in range(148)])
print("two-group ms", t.ms_two_group)
t.close()
