"""Generate tests/golden/tg_reference.npz from the REFERENCE'S OWN two-group Python (run in the build container only).

The reference modules are imported unmodified from /root/reference/src/two_group/hygeia on top of oracle/shim_tf, a
NumPy-backed stand-in for the slice of TensorFlow 2.3 / TFP 0.11 they touch (neither can be installed here).  Recorded:
  * CaseControlProposal.proposal_fn_standard_filter / initial_proposal_fn_standard_filter  (case_control_proposal_mappings.py:11-216)
  * CaseControlRegimeModel.transition_fn(step, prev).log_prob(next) with its three parts    (case_control_regime_model.py:80-193,
                                                                                             case_control_distributions.py:138-151,246-291)
  * the hazards rho(d, r) the model builds, fp32, including where the "0.1 if not finite" branch fires (:111-168)
  * OptimalFiniteState / SystematicResampling with an injected uniform                       (resampling_functions.py:7-69)
  * compute_log_backward_kernel_from_transition_matrix                                       (smoothing_functions.py:46-59)

    python tests/golden/make_golden_tg.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim_tf"))
sys.path.insert(0, "/root/reference/src/two_group")
sys.path.insert(0, ROOT)

# the reference pins numpy 1.18 (requirements.txt:39) and spells infinity np.Inf / np.math.inf, both removed in NumPy 2
import math as _math  # noqa: E402
if not hasattr(np, "Inf"):
    np.Inf = np.inf
if not hasattr(np, "math"):
    np.math = _math

import tensorflow as tf  # noqa: E402  (the stand-in)
from hygeia import (case_control_proposal_mappings, case_control_regime_model, resampling_functions,  # noqa: E402
                    smoothing_functions, filter_and_smoother_algorithm)

R, U = 6, 3
MU = np.array([0.95, 0.05, 0.8, 0.2, 0.5, 0.5], np.float32)
SIGMA = np.array([0.05, 0.05, 0.1, 0.1, 0.1, 0.2886751], np.float32)
OMEGA_CONTROL = np.array([0.995, 0.975, 0.95, 0.925, 0.9, 0.9])


def T(x, dt=None):
    return tf._t(np.asarray(x), dt)


def build_model(theta_p):
    """The constructor arguments exactly as run_inference_two_groups.py:118-167 derives them (fp32)."""
    f32 = np.float32
    p_softmax = theta_p.astype(f32)                                    # log P of the control regimes (get_estimated_control_group_param)
    omega_logit_control = np.log(OMEGA_CONTROL / (1 - OMEGA_CONTROL)).astype(f32)
    inv_logit = lambda x: (np.exp(x) / (1 + np.exp(x))).astype(f32)    # noqa: E731  (:147-148)
    omega_case = (0.8 * np.ones(R)).astype(f32)
    omega_inv_logit_case = inv_logit(omega_case)                        # (:149)  the reference applies inv_logit to omega_case itself
    omega_control = inv_logit(omega_logit_control)
    omega_inv_logit_control = inv_logit(omega_control)                  # (:151)
    merge_log_prob, split_log_prob = np.log(0.1), np.log(0.01)
    p_merged = np.array([[np.log(1 - np.exp(merge_log_prob)), merge_log_prob], [split_log_prob, np.log(1 - np.exp(split_log_prob))]], f32)
    kappa = (2 * np.ones(R)).astype(f32)
    model = case_control_regime_model.CaseControlRegimeModel(
        n_methylation_regimes=R, mu_true=T(MU), sigma_true=T(SIGMA), P_softmax_control=T(p_softmax), P_softmax_merged=T(p_merged),
        omega_inv_logit_control=T(omega_inv_logit_control), omega_inv_logit_case=T(omega_inv_logit_case), minimum_duration=U,
        kappa_control=T(kappa), kappa_case=T(kappa), n_total_reads_control=None, n_total_reads_case=None)
    return model, dict(omega_inv_logit_control=omega_inv_logit_control, omega_inv_logit_case=omega_inv_logit_case, p_merged=p_merged)


def states(rng, n, dmax=40):
    m = rng.integers(0, 2, n)
    dc = rng.integers(1, dmax, n); rc = rng.integers(0, R, n)
    dk = np.where(m == 1, dc, rng.integers(1, dmax, n)); rk = np.where(m == 1, rc, rng.integers(0, R, n))
    return dict(merged_state=m.astype(np.int32), control_state=np.stack([dc, rc], -1).astype(np.int32), case_state=np.stack([dk, rk], -1).astype(np.int32))


def end_to_end(T_sites, S, data_seed, seed, chain):
    """The reference's whole filter + backward simulation (filter_and_smoother_algorithm.run, exactly as
    run_inference_two_groups.py:261-276 calls it) on the stand-in, with the random draws injected: the systematic-resampling
    uniform of filter step t, the phantom initial regime and the categorical draws of the backward pass come from the same
    Philox counters, by inverse CDF in particle order, as oracle/tg_oracle.py and the CUDA path use."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tg_oracle
    from _tg_case import make_case
    fsa = filter_and_smoother_algorithm
    f32 = np.float32
    c = make_case(T_sites, S, seed=data_seed)
    inv_logit = lambda x: (np.exp(x) / (1 + np.exp(x))).astype(f32)   # noqa: E731
    p_softmax = c["logP"].copy(); p_softmax[~np.isfinite(p_softmax)] = 0.0
    omega_logit_control = np.log(c["omega_control"] / (1 - c["omega_control"])).astype(f32)
    kappa = (2 * np.ones(R)).astype(f32)
    p_merged = np.array([[np.log(0.9), np.log(0.1)], [np.log(0.01), np.log(0.99)]], f32)
    model = case_control_regime_model.CaseControlRegimeModel(
        n_methylation_regimes=R, mu_true=T(np.asarray(c["mu"], f32)), sigma_true=T(np.asarray(c["sigma"], f32)),
        P_softmax_control=T(p_softmax.astype(f32)), P_softmax_merged=T(p_merged),
        omega_inv_logit_control=T(inv_logit(inv_logit(omega_logit_control))), omega_inv_logit_case=T(inv_logit((0.8 * np.ones(R)).astype(f32))),
        minimum_duration=U, kappa_control=T(kappa), kappa_case=T(kappa),
        n_total_reads_control=T(c["nt_c"].T.astype(f32)), n_total_reads_case=T(c["nt_k"].T.astype(f32)))
    prop = case_control_proposal_mappings.CaseControlProposal(R)
    obs = dict(control=T(c["nm_c"].T.astype(f32)), case=T(c["nm_k"].T.astype(f32)))
    st = {"step": 0, "bt": T_sites - 1}
    orig = fsa._filter_one_step

    def wrapped(step, *a, **k):
        st["step"] = int(np.asarray(step))
        return orig(step, *a, **k)
    fsa._filter_one_step = wrapped
    tf.random.uniform_hook = lambda shp: np.full(shp, float(tg_oracle.tg_uniform(seed, chain, tg_oracle.TAG_FILTER, st["step"])))

    def inv_cdf(row, u):
        p = np.exp(row - np.max(row)); cdf = np.cumsum(p)
        return int(min(np.searchsorted(cdf, u * cdf[-1], side="left"), row.shape[0] - 1))

    def cat_init(logits, n):
        u = float(tg_oracle.tg_uniform(seed, chain, tg_oracle.TAG_PHANTOM, 0))
        return np.array([[inv_cdf(logits[b], u)] * n for b in range(logits.shape[0])])

    def cat_backward(logits, n):
        t = st["bt"]; st["bt"] -= 1
        nd = n if logits.shape[0] == 1 else logits.shape[0]
        us = tg_oracle.tg_uniform(seed, chain, tg_oracle.TAG_BACKWARD, np.arange(nd, dtype=np.uint64) * np.uint64(1 << 32) + np.uint64(t))
        if logits.shape[0] == 1:
            return np.array([[inv_cdf(logits[0], us[j]) for j in range(n)]])
        return np.array([[inv_cdf(logits[j], us[j])] for j in range(nd)])
    tf.random.categorical_hook = cat_init
    init = model.intitial_state_dist(batch_size=1)
    tf.random.categorical_hook = cat_backward
    try:
        bs, wT = fsa.run(observations=obs, initial_state_prior=init, transition_fn=model.transition_fn, observation_fn=model.observation_fn,
                         proposal_fn=prop.proposal_fn_standard_filter, initial_proposal=prop.initial_proposal_fn_standard_filter,
                         num_particles=50 * (2 * R + R * R), num_resampled_ancestors=50, optimal_resampling=True, multinomial_resampling=False,
                         num_simulations=25)
    finally:
        fsa._filter_one_step = orig
    wT = np.asarray(wT)
    return dict(T=T_sites, S=S, data_seed=data_seed, seed=seed, chain=chain,
                traj_merged=np.asarray(bs.particle["merged_state"]).astype(np.int16),
                traj_control=np.asarray(bs.particle["control_state"]).astype(np.int32), traj_case=np.asarray(bs.particle["case_state"]).astype(np.int32),
                log_norm=np.float64(tg_oracle.logsumexp(np.where(np.isfinite(wT), wT, -np.inf))), n_final_finite=int(np.isfinite(wT).sum()))


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    # ---- proposals ----
    prop = case_control_proposal_mappings.CaseControlProposal(R)
    anc = states(rng, 24)
    anc["control_state"][:4, 0] = [1, 2, 3, 4]; anc["case_state"][:4, 0] = [1, 2, 3, 1]   # around the minimum duration
    pp = prop.proposal_fn_standard_filter({k: T(v) for k, v in anc.items()})
    for k, v in anc.items():
        out[f"anc_{k}"] = v
    for k, v in pp.items():
        out[f"prop_{k}"] = np.asarray(v)
    ip = prop.initial_proposal_fn_standard_filter()
    for k, v in ip.items():
        out[f"init_prop_{k}"] = np.asarray(v)
    # ---- transition log-probabilities of every proposal of every ancestor, steps 0 and 5 ----
    theta_p = rng.normal(size=(R, R))
    model, par = build_model(theta_p)
    out["theta_p"] = theta_p
    for k, v in par.items():
        out[k] = v
    ancT = {k: T(v) for k, v in anc.items()}
    d = model.transition_fn(5, ancT)
    out["trans_step5"] = np.asarray(d.log_prob(pp))
    for k, v in d.log_prob_parts(pp).items():
        out[f"trans_step5_{k}"] = np.asarray(v)
    # step 0 as the filter's first step uses it (filter_and_smoother_algorithm.py:141-172): previous state = the phantom initial
    # state (merged, durations 0, regime r), next states = the R^2 initial proposals
    ph = dict(merged_state=np.ones(R, np.int32), control_state=np.stack([np.zeros(R, np.int32), np.arange(R, dtype=np.int32)], -1),
              case_state=np.stack([np.zeros(R, np.int32), np.arange(R, dtype=np.int32)], -1))
    ipb = {k: T(np.broadcast_to(np.asarray(v), (R * R, R) + np.asarray(v).shape[2:]).copy()) for k, v in ip.items()}
    for k, v in ph.items():
        out[f"phantom_{k}"] = v
    out["trans_step0"] = np.asarray(model.transition_fn(0, {k: T(v) for k, v in ph.items()}).log_prob(ipb))
    # random (mostly impossible) next states against the same ancestors, broadcast [48, 24] like the proposals
    nxt = states(rng, 48 * 24, dmax=6)
    nxt = {k: T(v.reshape((48, 24) + v.shape[1:])) for k, v in nxt.items()}
    for k, v in nxt.items():
        out[f"rand_next_{k}"] = np.asarray(v)
    out["trans_rand_step5"] = np.asarray(model.transition_fn(5, ancT).log_prob(nxt))
    # ---- hazards: rho(d, r) as the model's own code builds them, d = 0..6000 ----
    dgrid = np.arange(0, 6001, dtype=np.int32)
    for group in ("control", "case"):
        rho = np.zeros((R, dgrid.shape[0]), np.float32)
        for r in range(R):
            st = dict(merged_state=T(np.zeros(dgrid.shape[0], np.int32)), control_state=T(np.stack([dgrid, np.full_like(dgrid, r)], -1)),
                      case_state=T(np.stack([dgrid, np.full_like(dgrid, r)], -1)))
            d = model.transition_fn(5, st)
            cs = d.model["control_state"](st["merged_state"])
            if group == "control":
                rho[r] = np.asarray(cs._rho)
            else:
                ks = d.model["case_state"](st["merged_state"], T(np.stack([dgrid + 1, np.full_like(dgrid, r)], -1)))
                rho[r] = np.asarray(ks._rho)
        out[f"rho_{group}"] = rho
    out["rho_d"] = dgrid
    # ---- optimal finite-state resampling with injected uniforms (fp32 log-weights, as _filter_one_step passes them) ----
    for i, (n, conc, u) in enumerate([(300, 0.05, 0.37), (2400, 0.02, 0.81), (120, 0.5, 0.05), (60, 5.0, 0.5)]):
        w = rng.dirichlet(np.ones(n) * conc)
        lw = np.log(np.maximum(w, 1e-300))
        lw = (lw - np.log(np.exp(lw).sum())).astype(np.float32)
        tf.random.uniform_hook = lambda shp, u=u: np.full(shp, u)
        _, parents, log_c, unbiased, _, _ = resampling_functions.OptimalFiniteState(T(lw), 50, [])
        out[f"ofs{i}_logw"] = lw; out[f"ofs{i}_u"] = np.float64(u)
        out[f"ofs{i}_parents"] = np.asarray(parents); out[f"ofs{i}_log_c"] = np.float64(log_c); out[f"ofs{i}_unbiased"] = bool(np.asarray(unbiased))
    # ---- backward kernel ----
    tm = rng.normal(size=(7, 40)).astype(np.float32)
    tm[rng.random(tm.shape) < 0.4] = -np.inf
    pw = rng.normal(size=40) * 5
    pw[rng.random(40) < 0.2] = -np.inf
    out["bk_trans"] = tm; out["bk_prev_logw"] = pw
    out["bk_out"] = np.asarray(smoothing_functions.compute_log_backward_kernel_from_transition_matrix(T(pw), T(tm)))
    # ---- the whole filter + backward simulation, end to end (short: sojourns stay in the exact-hazard region; long: the
    #      case group's hazard reaches the fixed value 0.1 from d = 94 on) ----
    for tag, (Ts, S, ds, sd, chn) in (("e2e_short", (60, 2, 5, 1, 0)), ("e2e_long", (150, 3, 6, 2, 1))):
        e = end_to_end(Ts, S, ds, sd, chn)
        for k, v in e.items():
            out[f"{tag}_{k}"] = v
        print(tag, "log_norm", e["log_norm"], "finite final particles", e["n_final_finite"], "split fraction", float((e["traj_merged"] == 0).mean()))
    np.savez_compressed(os.path.join(HERE, "tg_reference.npz"), **out)
    for g in ("control", "case"):
        rho = out[f"rho_{g}"]
        first = [int(np.argmax((rho[r, 200:] == np.float32(0.1)) & (np.abs(rho[r, 199:-1] - 0.1) > 1e-6))) + 200 if ((rho[r, 200:] == np.float32(0.1)).any()) else -1 for r in range(R)]
        print(g, "rho(d=50):", rho[:, 50], "first d with the fixed value 0.1:", first)
    print("written", len(out), "arrays")


if __name__ == "__main__":
    main()
