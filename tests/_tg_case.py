"""Seeded two-group (case/control) test cases shared by the emulation and the GPU parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import tg_oracle  # noqa: E402
from hygeia_b200 import model, synthetic  # noqa: E402


def make_case(T, S, seed=5, R=6, u=3, omega_case=0.8, d_max=None):
    """Counts for a control group and a case group that differs on the stretch [T/3, T/2); the oracle's model object."""
    rng = np.random.default_rng(seed)
    mu, sigma = model.DEFAULT_MU[:R], model.DEFAULT_SIGMA[:R]
    regimes = synthetic.simulate_regimes(T, rng) % R
    reg_case = regimes.copy()
    reg_case[T // 3: T // 2] = (regimes[T // 3: T // 2] + 2) % R
    nt_c, nm_c = synthetic.simulate_counts(regimes, S, rng, mu=mu, sigma=sigma)
    nt_k, nm_k = synthetic.simulate_counts(reg_case, S, rng, mu=mu, sigma=sigma)
    alpha, beta = model.beta_parameters(mu, sigma)
    theta = np.concatenate([np.linspace(-0.5, 0.5, R * (R - 1)), np.log(np.asarray(model.DEFAULT_OMEGA[:R]) / (1 - np.asarray(model.DEFAULT_OMEGA[:R])))])
    logP, om_logit = tg_oracle.control_params_from_theta(theta, R)
    om_c = 1.0 / (1.0 + np.exp(-om_logit))
    m = tg_oracle.TwoGroupModel(R, logP, om_c, np.full(R, omega_case), np.full(R, 2.0), np.full(R, 2.0), u,
                                d_max=(T + 10) if d_max is None else d_max)
    return dict(T=T, S=S, R=R, u=u, regimes=regimes, reg_case=reg_case, nt_c=nt_c, nm_c=nm_c, nt_k=nt_k, nm_k=nm_k,
                alpha=alpha, beta=beta, theta=theta, logP=logP, omega_control=om_c, omega_case=np.full(R, omega_case), model=m,
                mu=mu, sigma=sigma)
