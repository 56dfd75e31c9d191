"""Host-side parameter plumbing of the single-group model.

Mirrors the R helpers the reference CLI calls before it crosses into C++
(/root/reference/src/single_group/src/r/model_functions.R):

* ``get_known_parameters``                -> model_functions.R:36-63
* ``convert_model_parameters_to_theta``   -> model_functions.R:65-78
* ``convert_theta_to_model_parameters``   -> model_functions.R:81-111

including the column-major extraction quirk of ``p[p != -1]``:
the P matrix read back from ``p.csv`` is effectively transposed on the way in.
"""
from __future__ import annotations

import numpy as np

DEFAULT_MU = (0.95, 0.05, 0.80, 0.20, 0.50, 0.50)            # nextflow.config:5
DEFAULT_SIGMA = (0.05, 0.05, 0.1, 0.1, 0.1, 0.2886751)       # nextflow.config:6
DEFAULT_OMEGA = (0.995, 0.975, 0.950, 0.925, 0.900, 0.900)   # bin/estimate_parameters_and_regimes:33-37
DEFAULT_KAPPA = (2.0,) * 6                                   # bin/estimate_parameters_and_regimes:28-32


def beta_parameters(mu, sigma):
    """alpha_r = mu*nu, beta_r = (1-mu)*nu, nu = mu(1-mu)/sigma^2 - 1 (model_functions.R:44-47)."""
    mu = np.asarray(mu, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    nu = mu * (1 - mu) / sigma ** 2 - 1
    return mu * nu, (1 - mu) * nu


def get_known_parameters(mu=DEFAULT_MU, sigma=DEFAULT_SIGMA, u=3, is_kappa_fixed=True, kappa=None):
    """vartheta = (u, R, alpha, beta, isKappaFixed[, kappa]) and dim(theta)."""
    alpha, beta = beta_parameters(mu, sigma)
    R = len(alpha)
    if kappa is None:
        kappa = np.full(R, 2.0)
    if is_kappa_fixed:
        vartheta = np.concatenate([[u, R], alpha, beta, [1.0], np.asarray(kappa, dtype=np.float64)])
        dim_theta = R * R
    else:
        vartheta = np.concatenate([[u, R], alpha, beta, [0.0]])
        dim_theta = R * (R + 1)
    return np.ascontiguousarray(vartheta, dtype=np.float64), dim_theta


def default_p(R=6):
    """Initial transition matrix of the CLI: 1/(R-1)=1/5 off the diagonal (bin/...:243-247)."""
    p = np.full((R, R), 1.0 / 5.0)
    np.fill_diagonal(p, 0.0)
    return p


def convert_model_parameters_to_theta(p, omega):
    """theta = (log p[p != -1] in COLUMN-major order, logit(omega)) (model_functions.R:65-78)."""
    p = np.array(p, dtype=np.float64, copy=True)
    R = p.shape[0]
    np.fill_diagonal(p, -1.0)
    flat = p.flatten(order="F")          # R stores matrices column-major
    off = flat[flat != -1.0]
    assert off.size == R * (R - 1)
    omega = np.asarray(omega, dtype=np.float64)
    return np.concatenate([np.log(off), np.log(omega) - np.log(1.0 - omega)])


def convert_theta_to_model_parameters(theta, R=6):
    """theta -> (P row-wise softmax with zero diagonal, omega) (model_functions.R:81-111)."""
    theta = np.asarray(theta, dtype=np.float64)
    p = np.zeros((R, R))
    for rr in range(R):
        blk = theta[rr * (R - 1):(rr + 1) * (R - 1)]
        m = blk.max()
        lz = m + np.log(np.exp(blk - m).sum())
        p[rr, np.arange(R) != rr] = np.exp(blk - lz)
    omega = 1.0 / (1.0 + np.exp(-theta[R * (R - 1):R * R]))
    return p, omega


def default_theta(R=6, omega=DEFAULT_OMEGA):
    return convert_model_parameters_to_theta(default_p(R), omega)
