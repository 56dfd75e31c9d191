"""Deterministic synthetic WGBS count data from the reference's generative model.

Follows SURVEY.md section 8(d): regimes and sojourn times from the change-point
model of the reference (Model.h:62-75; singleGroup.h:485-557 -- new segment with
probability rho(d, r), regime from row r of P, sojourn law = u + NegBin(kappa, omega)),
coverage n ~ Poisson(lam) with a fraction ``p_missing`` of entries set to 0,
methylated counts x ~ BetaBinomial(n, alpha_r, beta_r), positions = 10000 +
cumulative (1 + Geometric(1/100)) gaps.  Counts are uint16, layout [S][T]
(site index fastest) -- the layout the device emission kernel streams.
"""
from __future__ import annotations

import numpy as np

from . import model as _model

# relative lengths of the 22 synthetic "chromosomes" (a synthetic partition)
CHROM_WEIGHTS = (2.30, 2.20, 1.65, 1.50, 1.50, 1.45, 1.50, 1.25, 1.20, 1.35, 1.30, 1.25,
                 0.75, 0.85, 0.85, 1.05, 1.15, 0.70, 1.05, 0.75, 0.40, 0.60)


def chromosome_lengths(total_sites: int, weights=CHROM_WEIGHTS):
    w = np.asarray(weights, dtype=np.float64)
    t = np.floor(w / w.sum() * total_sites).astype(np.int64)
    t[0] += total_sites - t.sum()
    return [int(x) for x in t]


def simulate_regimes(T, rng, u=3, kappa=2.0, omega=_model.DEFAULT_OMEGA, p=None):
    """Regime path of length T: segments of length u + NegBin(kappa, 1-omega_r) (support 0,1,...)."""
    R = len(omega)
    if p is None:
        p = _model.default_p(R)
    regimes = np.empty(T, dtype=np.uint8)
    t = 0
    r = int(rng.integers(R))
    while t < T:
        # numpy's negative_binomial(n, p) counts failures before n successes with success prob p;
        # the reference's law (misc.h:673-693) has pmf ~ omega^k (1-omega)^kappa, i.e. success prob 1-omega.
        length = u + int(rng.negative_binomial(kappa, 1.0 - omega[r]))
        regimes[t:t + length] = r
        t += length
        r = int(rng.choice(R, p=p[r]))
    return regimes


def simulate_counts(regimes, S, rng, lam=30.0, p_missing=0.05, mu=_model.DEFAULT_MU, sigma=_model.DEFAULT_SIGMA):
    """(n_total, n_meth) as uint16 [S][T]."""
    alpha, beta = _model.beta_parameters(mu, sigma)
    T = regimes.shape[0]
    n = rng.poisson(lam, size=(S, T))
    n[rng.random((S, T)) < p_missing] = 0
    pm = rng.beta(alpha[regimes][None, :].repeat(S, 0), beta[regimes][None, :].repeat(S, 0))
    x = rng.binomial(n, pm)
    return np.ascontiguousarray(n, dtype=np.uint16), np.ascontiguousarray(x, dtype=np.uint16)


def simulate_positions(T, rng):
    return (10000 + np.cumsum(1 + rng.geometric(1.0 / 100.0, size=T))).astype(np.uint32)


def make_chain(T, S, seed, lam=30.0, p_missing=0.05, u=3):
    """One chain: dict(positions[T] u32, n_total[S][T] u16, n_meth[S][T] u16, regimes[T] u8)."""
    rng = np.random.default_rng(seed)
    regimes = simulate_regimes(T, rng, u=u)
    n_total, n_meth = simulate_counts(regimes, S, rng, lam=lam, p_missing=p_missing)
    return dict(positions=simulate_positions(T, rng), n_total=n_total, n_meth=n_meth, regimes=regimes)
