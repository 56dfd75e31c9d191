import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from hygeia_b200 import model, philox, synthetic
from hygeia_b200.single_group import Session
from _oracle import Oracle
from test_gpu_single_group import run_chain
T,S=6000,4
ch = synthetic.make_chain(T,S,seed=555)
vt,_=model.get_known_parameters(); th=model.default_theta()
u = philox.uniforms_by_site(21,3,T)
want = Oracle().run(vt, th, u, ch["n_total"], ch["n_meth"], ch["positions"], want_ancestors=True)
s=Session(0)
got = run_chain(s, vt, th, ch["n_total"], ch["n_meth"], ch["positions"], seed=21, chain_id=3)
d = np.where(got["k_kept"]!=want["k_kept"])[0]
print('n diff K', len(d), d[:10], got["k_kept"][d[:10]], want["k_kept"][d[:10]])
print('logz rel', np.max(np.abs(got["logz"]-want["logz"])/np.abs(want["logz"])), 'dp', np.nanmax(np.abs(got["regime_probs"]-want["regime_probs"])))
print('ncurr eq', (got["n_curr"]==want["n_curr"]).all(), 'fin eq', (got["finalised_at"]==want["finalised_at"]).mean(), 'drew', (got["drew_uniform"]==want["drew_uniform"]).mean())
t0=d[0]
print('around', t0, got["k_kept"][t0-2:t0+3], want["k_kept"][t0-2:t0+3], got["logz"][t0-1:t0+2]-want["logz"][t0-1:t0+2])
