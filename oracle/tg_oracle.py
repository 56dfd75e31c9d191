"""oracle/tg_oracle.py -- TEST INFRASTRUCTURE, not product code.

NumPy restatement of Hygeia's TWO-GROUP inference path (case/control model, deterministic-proposal particle filter with
optimal finite-state resampling, backward simulation).  Reference (paths relative to /root/reference/src/two_group):

    run_inference_two_groups.py:110-167,233-240,283-322        parameters, test functions, outputs
    hygeia/filter_and_smoother_algorithm.py:141-288,334-446    first step, filter step, padding, backward simulation
    hygeia/case_control_regime_model.py:19-23,80-231           beta parameters, merged-state law, rho, emission
    hygeia/case_control_distributions.py:138-151,246-291       control / case transition log-probabilities
    hygeia/case_control_proposal_mappings.py:11-134,175-216    the 2R + R^2 = 48 proposals per ancestor
    hygeia/resampling_functions.py:7-69                        OptimalFiniteState, SystematicResampling
    hygeia/smoothing_functions.py:46-59                        backward kernel from the transition matrix

PARITY UNPINNED.  The reference runs on TensorFlow 2.3 / TFP 0.11 (src/two_group/requirements.txt:74,78), which cannot be
installed here (no network); it ships no tests or golden vectors for this path.  This file restates the published
algorithm from the reference's own call sites.  Deliberate, documented differences from the reference's arithmetic:
  * everything is fp64 (the reference evaluates the model in fp32 with fp64 weights);
  * TFP's NegativeBinomial log_prob / log_survival_function are replaced by an fp64 evaluation of the same hazard
    rho(d) = pmf(d-u) / P(X >= d-u) (see hazard_table; the reference's "0.1 where not finite" fp32 artefact is not reproduced);
  * random draws (the systematic-resampling uniform of each filter step, the categorical draws of the backward pass and
    the phantom initial regime) come from Philox keyed by (seed, chain) and indexed by site, and categorical sampling is
    inverse-CDF in particle order -- TensorFlow's own stream cannot be reproduced.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hygeia_b200 import philox  # noqa: E402  (host copy of the counter-based generator the kernels use)

NEG_INF = -np.inf
TAG_FILTER, TAG_BACKWARD, TAG_PHANTOM = 0x46494C54, 0x42414B57, 0x5048414E   # "FILT", "BAKW", "PHAN"


def tg_uniform(seed, chain, tag, index):
    """One uniform in [0,1): philox(key = seed, counter = (index_lo, index_hi, chain, tag))."""
    index = np.asarray(index, dtype=np.uint64)
    x0, x1, _, _ = philox.philox4x32_10(index & np.uint64(0xFFFFFFFF), index >> np.uint64(32), np.uint64(chain & 0xFFFFFFFF),
                                        np.uint64(tag), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    hi = x0.astype(np.uint64) >> np.uint64(5)
    lo = x1.astype(np.uint64) >> np.uint64(6)
    return (hi * np.float64(67108864.0) + lo) * np.float64(1.0 / 9007199254740992.0)


def control_params_from_theta(theta, R):
    """get_estimated_control_group_param (run_inference_two_groups.py:76-89): log P (row-normalised exp(theta)), omega logits."""
    p = np.zeros((R, R))
    i = 0
    for r in range(R):
        for r1 in range(R):
            if r != r1:
                p[r, r1] = np.exp(theta[i])
                i += 1
        p[r, :] /= p[r, :].sum()
    with np.errstate(divide="ignore"):
        logp = np.log(p)
    return logp, np.asarray(theta[-R:], dtype=np.float64)


def hazard_table(omega, kappa, u, d_max):
    """rho[r][d] for d = 0..d_max: pmf(d-u) / P(X >= d-u), X ~ NB(kappa, omega) (case_control_regime_model.py:111-168);
    0 below u.  Evaluated by the backward recurrence of the inverse hazard g(k) = P(X >= k) / pmf(k):
        g(k) = 1 + pmf(k+1)/pmf(k) * g(k+1) = 1 + omega (k + kappa)/(k + 1) * g(k+1),   g(inf) = 1/(1 - omega),
    started far enough beyond d_max that the start-up error (contracted by ~omega per step) is below 1e-18.  It never
    leaves the finite range, so the reference's "0.1 where not finite" branch (an fp32 underflow of TFP's survival function
    at sojourns of several hundred sites) is not reproduced."""
    R = len(omega)
    rho = np.zeros((R, d_max + 1))
    for r in range(R):
        om, ka = float(omega[r]), float(kappa[r])
        n_extra = int(min(np.ceil(-41.5 / np.log(om)), 5e7))
        g = 1.0 / (1.0 - om)
        for k in range(d_max - u + n_extra - 1, -1, -1):
            g = 1.0 + (om * (k + ka) / (k + 1.0)) * g
            if k <= d_max - u:
                v = 1.0 / g
                rho[r, k + u] = v if np.isfinite(v) else 0.1
    return rho


def reference_hazard_table(omega, kappa, u, d_max):
    """rho[r][d], d = 0..d_max, AS THE REFERENCE EVALUATES IT (case_control_regime_model.py:111-168): exp(log_prob(d - u) -
    log_survival_function(d - u - 1)) of tfd.NegativeBinomial(kappa, probs = omega) in fp32, and the fixed value 0.1 wherever
    that is not finite.  TFP's survival function is log1p(-cdf); in fp32 the cdf rounds to 1 once 1 - cdf < 2^-25, so from that
    sojourn on (d = 94 for omega = 0.8, 197 for 0.9, ... 4071 for 0.995 at kappa = 2) EVERY regime's hazard is 0.1.  fp64
    special functions (SciPy) rounded to fp32 where TFP holds fp32 values; pinned to the reference's own module run on
    oracle/shim_tf by tests/test_two_group_reference.py."""
    from scipy import special
    omega = np.asarray(omega, dtype=np.float64); kappa = np.asarray(kappa, dtype=np.float64)
    R = omega.shape[0]
    rho = np.zeros((R, d_max + 1), dtype=np.float64)
    d = np.arange(d_max + 1, dtype=np.float64)
    f32 = np.float32
    for r in range(R):
        p = f32(omega[r]); tc = np.float64(f32(kappa[r]))
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            lg = np.float64(f32(np.log(np.float64(p)) - np.log1p(-np.float64(p))))
            x = d - u
            logh = (tc * (-np.logaddexp(0.0, lg)) + x * (-np.logaddexp(0.0, -lg)) - special.betaln(1.0 + x, tc) - np.log(tc + x)).astype(f32)
            cdf = special.betainc(tc, 1.0 + (x - 1.0), 1.0 / (1.0 + np.exp(lg))).astype(f32)
            logsf = np.log1p(-cdf).astype(f32)
            logh = np.where(d >= u, logh, f32(-np.inf))
            logsf = np.where(d > u, logsf, f32(0.0))
            val = np.where(logh == -np.inf, f32(0.0), np.exp((logh - logsf).astype(f32)).astype(f32))
            val = np.where(np.isfinite(val), val, f32(0.1))
        rho[r] = val.astype(np.float64)
    return rho


class TwoGroupModel:
    def __init__(self, R, log_p_control, omega_control, omega_case, kappa_control, kappa_case, u,
                 p_merged=((0.9, 0.1), (0.01, 0.99)), d_max=4096):
        self.R, self.u = R, int(u)
        self.logP = np.asarray(log_p_control, dtype=np.float64).copy()
        np.fill_diagonal(self.logP, NEG_INF)
        # tf.nn.softmax of the log-probabilities with a -inf diagonal (case_control_regime_model.py:90-95)
        m = self.logP.max(1, keepdims=True)
        self.logP = self.logP - (m + np.log(np.exp(self.logP - m).sum(1, keepdims=True)))
        self.rho_c = hazard_table(omega_control, kappa_control, self.u, d_max)
        self.rho_k = hazard_table(omega_case, kappa_case, self.u, d_max)
        with np.errstate(divide="ignore"):
            self.logPm = np.log(np.asarray(p_merged, dtype=np.float64))
        self.d_max = d_max

    # --- transition log-probability of `nxt` given `prv`; states are dicts of int arrays, broadcast together ---
    def log_trans(self, prv, nxt, step=1):
        R, u = self.R, self.u
        m, dc, rc, dk, rk = prv["m"], prv["dc"], prv["rc"], prv["dk"], prv["rk"]
        m2, dc2, rc2, dk2, rk2 = nxt["m"], nxt["dc"], nxt["rc"], nxt["dk"], nxt["rk"]
        with np.errstate(divide="ignore", invalid="ignore"):
            # merged indicator (case_control_regime_model.py:80-87)
            if step == 0:
                lm = np.where(m2 == 1, 0.0, NEG_INF)
            else:
                free = np.minimum(dk, dc) >= u
                lm = np.where(free, self.logPm[m, m2], np.where(m2 == m, 0.0, NEG_INF))
            # control (case_control_distributions.py:138-151)
            rho_c = np.ones_like(dc, dtype=np.float64) if step == 0 else self.rho_c[rc, np.minimum(dc, self.d_max)]
            lc = np.where(dc2 == 1, np.log(rho_c) + self.logP[rc, rc2],
                          np.log(1.0 - rho_c) + np.where((dc2 == dc + 1) & (rc2 == rc), 0.0, NEG_INF))
            # case (case_control_distributions.py:246-291), first matching rule
            rho_k = np.ones_like(dk, dtype=np.float64) if step == 0 else self.rho_k[rk, np.minimum(dk, self.d_max)]
            rule1 = np.where((rc2 == rk2) & (dc2 == dk2), 0.0, NEG_INF)
            # uniform over regimes different from the new control regime
            rule2 = np.where(rk2 != rc2, -np.log(R - 1.0), NEG_INF) + np.where(dk2 == 1, 0.0, NEG_INF)
            # uniform over regimes not in {new control regime, previous case regime}
            n_excl = np.where(rc2 == rk, 1.0, 2.0)
            unif2 = np.where((rk2 != rc2) & (rk2 != rk), -np.log(R - n_excl), NEG_INF)
            rule3 = np.where(dk2 == 1, 0.0, NEG_INF) + unif2
            rule4 = np.where(dk2 == 1, np.log(rho_k) + unif2,
                             np.log(1.0 - rho_k) + np.where((dk2 == dk + 1) & (rk2 == rk), 0.0, NEG_INF))
            lk = np.where(m2 == 1, rule1,
                          np.where((m == 1) & (dc2 != 1), rule2,
                                   np.where((rc2 == rk) & (m == 0), rule3, rule4)))
            out = lm + lc + lk
        return np.where(np.isnan(out), NEG_INF, out)

    # --- the 2R + R^2 proposals of each ancestor, stacked [48, M'] (case_control_proposal_mappings.py:11-134,175-207) ---
    def propose(self, anc):
        R = self.R
        m, dc, rc, dk, rk = (np.asarray(anc[k]) for k in ("m", "dc", "rc", "dk", "rk"))
        Mp = m.shape[0]
        I = 2 * R + R * R
        out = {k: np.zeros((I, Mp), dtype=np.int64) for k in ("m", "dc", "rc", "dk", "rk")}
        out["m"][0], out["dc"][0], out["rc"][0], out["dk"][0], out["rk"][0] = m, dc + 1, rc, dk + 1, rk
        for idx in range(1, R):                                # control jumps, skipping the case regime
            reg = np.where(idx <= rk, idx - 1, idx)
            out["m"][idx], out["dc"][idx], out["rc"][idx], out["dk"][idx], out["rk"][idx] = 0, 1, reg, dk + 1, rk
        for idx in range(R, 2 * R - 1):                        # case jumps, skipping the control regime
            reg = np.where(idx < R + rc, idx - R, idx - R + 1)
            out["m"][idx], out["dc"][idx], out["rc"][idx], out["dk"][idx], out["rk"][idx] = 0, dc + 1, rc, 1, reg
        dmerge = np.where(m == 0, dc + 1, 0)                    # merge (durations 0 = impossible if already merged)
        out["m"][2 * R - 1], out["dc"][2 * R - 1], out["rc"][2 * R - 1], out["dk"][2 * R - 1], out["rk"][2 * R - 1] = 1, dmerge, rc, dmerge, rc
        p = 2 * R
        for i in range(R):                                      # both change: control regime i, case regime j
            for j in range(R):
                out["m"][p], out["dc"][p], out["rc"][p], out["dk"][p], out["rk"][p] = int(i == j), 1, i, 1, j
                p += 1
        return out

    def initial_particles(self):
        R = self.R
        i, j = np.meshgrid(np.arange(R), np.arange(R), indexing="ij")
        i, j = i.ravel(), j.ravel()
        return dict(m=(i == j).astype(np.int64), dc=np.ones(R * R, np.int64), rc=i.astype(np.int64), dk=np.ones(R * R, np.int64), rk=j.astype(np.int64))


def logsumexp(x):
    m = np.max(x)
    if not np.isfinite(m):
        return m
    return m + np.log(np.sum(np.exp(x - m)))


def optimal_finite_state(logw_norm, M, u):
    """resampling_functions.py:7-52 on normalised log-weights.  Returns (parents[M] or None, log_c, K)."""
    n = logw_norm.shape[0]
    order = np.argsort(-logw_norm, kind="stable")
    sw = logw_norm[order]
    rcs = np.cumsum(np.exp(sw)[::-1])[::-1]
    a, b, log_c = 0, -1, -1.0
    while a != b and a < n and a < M:
        with np.errstate(divide="ignore"):
            log_c_new = np.log(float(M - a)) - np.log(rcs[a])
        k_new = a + int(np.sum(log_c_new + sw[a:] > 0))
        a, b, log_c = k_new, a, log_c_new
    K = b
    if not K < n:
        K, log_c = n, NEG_INF
    if not np.isfinite(log_c):
        return None, log_c, K
    L = M - K
    parents = np.zeros(M, dtype=np.int64)
    parents[:K] = order[:K]
    res = sw[K:]
    res = res - logsumexp(res)
    T = (np.arange(L) + u) / L
    Q = np.cumsum(np.exp(res))
    i = j = 0
    idx = np.zeros(L, dtype=np.int64)
    while j < L and i < res.shape[0]:
        if T[j] <= Q[i]:
            idx[j] = i
            j += 1
        else:
            i += 1
    parents[K:] = order[K + idx]
    return parents, log_c, K


def run(model: TwoGroupModel, lo_control, lo_case, M=50, n_backward=25, seed=0, chain=0):
    """filter_and_smoother_algorithm.run: filter over all sites, then backward simulation.
    lo_control / lo_case: T x R emission tables (sum over the samples of each group of the beta-binomial log-density).
    Returns dict(traj_m [T,B], traj_control [T,B,2], traj_case [T,B,2], log_norm, split_probs, regime_probs, taps)."""
    R = model.R
    T = lo_control.shape[0]
    I = 2 * R + R * R
    hist = []
    taps = dict(n_particles=np.zeros(T, np.int64), K=np.full(T, -1, np.int64), n_finite=np.zeros(T, np.int64), log_c=np.zeros(T))
    # ---- first step (filter_and_smoother_algorithm.py:141-172; case_control_regime_model.py:233-244) ----
    r_ph = int(np.floor(tg_uniform(seed, chain, TAG_PHANTOM, 0) * R))
    parts = model.initial_particles()
    phantom = dict(m=np.ones(1, np.int64), dc=np.zeros(1, np.int64), rc=np.full(1, r_ph), dk=np.zeros(1, np.int64), rk=np.full(1, r_ph))
    lt = model.log_trans(phantom, parts, step=0)
    w = lt + lo_control[0][parts["rc"]] + lo_case[0][parts["rk"]]
    w = np.where(np.isfinite(lt), w, NEG_INF)
    hist.append((parts, w))
    taps["n_particles"][0] = w.shape[0]
    taps["n_finite"][0] = int(np.isfinite(w).sum())
    # ---- filter steps (filter_and_smoother_algorithm.py:176-288) ----
    for t in range(1, T):
        pw, pp = w, parts
        finite = np.where(pw > NEG_INF)[0]
        F = finite.shape[0]
        lse = logsumexp(pw)
        logW = pw - lse
        use_unbiased = False
        if F > M:
            u = float(tg_uniform(seed, chain, TAG_FILTER, t))
            parents, log_c, K = optimal_finite_state(logW, M, u)
            taps["K"][t] = K
            if parents is None:   # log c infinite: multinomial draws by inverse CDF, unbiased weights (resampling_functions.py:42-47)
                use_unbiased = True
                cdf = np.cumsum(np.exp(logW))
                us = tg_uniform(seed, chain, TAG_FILTER, (np.arange(M, dtype=np.uint64) + 1) * np.uint64(1 << 32) + np.uint64(t))
                parents = np.minimum(np.searchsorted(cdf, us * cdf[-1], side="left"), pw.shape[0] - 1)
                log_c = 0.0
        else:
            parents, log_c = finite, 0.0
        taps["log_c"][t] = log_c
        anc = {k: v[parents] for k, v in pp.items()}
        prop = model.propose(anc)                                   # [48, M']
        Mp = parents.shape[0]
        anc_b = {k: np.broadcast_to(v[None, :], (I, Mp)) for k, v in anc.items()}
        lt = model.log_trans(anc_b, prop, step=t)
        lg = np.where(np.isfinite(lt), lt + lo_control[t][prop["rc"]] + lo_case[t][prop["rk"]], NEG_INF)
        if F <= M:
            wn = pw[parents][None, :] + lg
        elif use_unbiased:
            wn = -np.log(float(M)) + lse + lg
        else:
            wn = pw[parents][None, :] + lg - np.minimum(0.0, log_c + logW[parents])[None, :]
        parts = {k: v.reshape(-1) for k, v in prop.items()}         # proposal-major flattening
        w = wn.reshape(-1)
        hist.append((parts, w))
        taps["n_particles"][t] = w.shape[0]
        taps["n_finite"][t] = int(np.isfinite(w).sum())
    log_norm = logsumexp(w)
    # ---- backward simulation (filter_and_smoother_algorithm.py:368-446; smoothing_functions.py:46-59) ----
    B = n_backward
    traj = {k: np.zeros((T, B), dtype=np.int64) for k in ("m", "dc", "rc", "dk", "rk")}

    def categorical(logits, uu):
        m = np.max(logits)
        p = np.exp(logits - m)
        cdf = np.cumsum(p)
        return int(min(np.searchsorted(cdf, uu * cdf[-1], side="left"), logits.shape[0] - 1))

    parts, w = hist[T - 1]
    us = tg_uniform(seed, chain, TAG_BACKWARD, np.arange(B, dtype=np.uint64) * np.uint64(1 << 32) + np.uint64(T - 1))
    for j in range(B):
        i = categorical(np.where(np.isfinite(w), w, NEG_INF), us[j])
        for k in traj:
            traj[k][T - 1, j] = parts[k][i]
    for t in range(T - 2, -1, -1):
        parts, w = hist[t]
        keep = np.where(w > NEG_INF)[0]
        pk = {k: v[keep] for k, v in parts.items()}
        us = tg_uniform(seed, chain, TAG_BACKWARD, np.arange(B, dtype=np.uint64) * np.uint64(1 << 32) + np.uint64(t))
        for j in range(B):
            nxt = {k: np.full(keep.shape[0], traj[k][t + 1, j]) for k in traj}
            lt = model.log_trans(pk, nxt, step=t + 1)
            logits = np.where(np.isfinite(lt), lt + w[keep], NEG_INF)
            i = categorical(logits, us[j])
            for k in traj:
                traj[k][t, j] = pk[k][i]
    # ---- outputs (run_inference_two_groups.py:233-240,294-296) ----
    split_probs = (traj["m"] == 0).mean(1)
    regime_probs = np.concatenate([np.stack([(traj["rc"] == i).mean(1) for i in range(R)], -1),
                                   np.stack([(traj["rk"] == i).mean(1) for i in range(R)], -1)], -1)
    return dict(traj_m=traj["m"], traj_control=np.stack([traj["dc"], traj["rc"]], -1), traj_case=np.stack([traj["dk"], traj["rk"]], -1),
                log_norm=log_norm, split_probs=split_probs, regime_probs=regime_probs, taps=taps, final_weights=w)
