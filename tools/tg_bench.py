"""Timing of the two-group path on one GPU: n_chains segments of T sites, S samples per group (tuning aid, not bench.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _tg_case import make_case  # noqa: E402
from hygeia_b200.two_group import TwoGroupSession, summarise  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50
n_chains = int(sys.argv[3]) if len(sys.argv) > 3 else 148
c = make_case(T, S, seed=5, d_max=4096)
s = TwoGroupSession(0)
s.set_emission_model(c["mu"], c["sigma"], 3)
s.add_dataset(c["nt_c"], c["nm_c"])
s.add_dataset(c["nt_k"], c["nm_k"])
s.emission()
s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], 3, 50, 25, t_max=T)
specs = [dict(control_dataset=0, case_dataset=1, T=T, seed=1, chain_id=i) for i in range(n_chains)]
for rep in range(2):
    t0 = time.time()
    out = s.run(specs)
    wall = time.time() - t0
    ms = s.ms_two_group
    print(f"rep {rep}: T={T} S={S} chains={n_chains} device {ms:.1f} ms  wall {wall*1e3:.1f} ms  "
          f"{ms*1e3/T:.2f} us/site/chain-wave  {T*n_chains/ms*1e3:.3e} site-chains/s  {T*n_chains*2*S/ms*1e3:.3e} site.sample.chain/s", flush=True)
split, reg = summarise(out[0]["trajectories"], 6)
print("control acc", (reg[:, :6].argmax(1) == c["regimes"]).mean(), "case acc", (reg[:, 6:].argmax(1) == c["reg_case"]).mean(),
      "split acc", ((split > 0.5) == (c["regimes"] != c["reg_case"])).mean())
