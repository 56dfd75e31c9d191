"""Pin the two-group restatement (oracle/tg_oracle.py) to the REFERENCE'S OWN Python.

tests/golden/tg_reference.npz was produced by importing /root/reference/src/two_group/hygeia/*.py unmodified on top of
oracle/shim_tf (a NumPy stand-in for the slice of TensorFlow/TFP they touch; tests/golden/make_golden_tg.py): proposal maps,
transition log-densities, hazards incl. the fp32 "0.1 if not finite" branch, optimal finite-state resampling, backward kernel.
What stays unpinned: TensorFlow's random streams and the last bits of its fp32 special functions."""
import os
import sys

import numpy as np
import pytest

from conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tg_oracle  # noqa: E402

R, U = 6, 3
OMEGA_CONTROL = np.array([0.995, 0.975, 0.95, 0.925, 0.9, 0.9])


@pytest.fixture(scope="module")
def g():
    return golden("tg_reference.npz")


def _as_oracle_state(merged, control, case):
    return dict(m=merged.astype(np.int64), dc=control[..., 0].astype(np.int64), rc=control[..., 1].astype(np.int64),
                dk=case[..., 0].astype(np.int64), rk=case[..., 1].astype(np.int64))


def _model(g, hazard="reference"):
    m = tg_oracle.TwoGroupModel(R, g["theta_p"], OMEGA_CONTROL, np.full(R, 0.8), np.full(R, 2.0), np.full(R, 2.0), U, d_max=6000)
    if hazard == "reference":
        m.rho_c = g["rho_control"].astype(np.float64)
        m.rho_k = g["rho_case"].astype(np.float64)
    return m


def test_proposal_map_is_the_reference(g):
    m = _model(g)
    anc = _as_oracle_state(g["anc_merged_state"], g["anc_control_state"], g["anc_case_state"])
    p = m.propose(anc)
    assert np.array_equal(p["m"], g["prop_merged_state"])
    assert np.array_equal(np.stack([p["dc"], p["rc"]], -1), g["prop_control_state"])
    assert np.array_equal(np.stack([p["dk"], p["rk"]], -1), g["prop_case_state"])
    i = m.initial_particles()
    assert np.array_equal(i["m"], g["init_prop_merged_state"][:, 0])
    assert np.array_equal(np.stack([i["dc"], i["rc"]], -1), g["init_prop_control_state"][:, 0])
    assert np.array_equal(np.stack([i["dk"], i["rk"]], -1), g["init_prop_case_state"][:, 0])


@pytest.mark.parametrize("step,key,nxt,prv", [(0, "trans_step0", "init_prop", "phantom"), (5, "trans_step5", "prop", "anc"),
                                              (5, "trans_rand_step5", "rand_next", "anc")])
def test_transition_log_density_is_the_reference(g, step, key, nxt, prv):
    """Every rule of ControlStateTransition / CaseStateTransition / the merged-indicator law, through the reference's own
    transition_fn(step, prev).log_prob(next); fp32 there, fp64 here.  Step 0 = the filter's first step (phantom initial state
    against the R^2 initial proposals)."""
    m = _model(g)
    anc = _as_oracle_state(g[f"{prv}_merged_state"], g[f"{prv}_control_state"], g[f"{prv}_case_state"])
    nx = _as_oracle_state(g[f"{nxt}_merged_state"], g[f"{nxt}_control_state"], g[f"{nxt}_case_state"])
    if nx["m"].shape[1] == 1:
        nx = {k: np.broadcast_to(v, (v.shape[0], anc["m"].shape[0])) for k, v in nx.items()}
    I, M = nx["m"].shape
    anc_b = {k: np.broadcast_to(v[None, :], (I, M)) for k, v in anc.items()}
    got = m.log_trans(anc_b, nx, step=step)
    want = g[key].astype(np.float64)
    want = np.where(np.isnan(want), -np.inf, want)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))          # the same transitions are possible
    f = np.isfinite(want)
    assert f.sum() > (150 if nxt == "prop" else 5)   # (step 0: each phantom regime reaches all R^2 initial states)
    assert np.allclose(got[f], want[f], rtol=2e-6, atol=2e-6)


def test_hazard_reference_mode_and_where_the_fixed_value_starts(g):
    """case_control_regime_model.py:111-168 builds rho = exp(log_prob - log_survival_function) in fp32 and replaces non-finite
    values by 0.1.  The survival function 1 - cdf underflows in fp32 where 1 - cdf < 2^-25: from there on EVERY regime's hazard
    is 0.1 -- from d = 94 for the case group (true hazard there 0.19), d = 197 ... 4071 for the control regimes."""
    from hygeia_b200.two_group import hazard_table
    want_first = {"control": [4071, 808, 401, 265, 197, 197], "case": [94] * 6}
    for group, omega in (("control", OMEGA_CONTROL), ("case", np.full(R, 0.8))):
        ref = g[f"rho_{group}"].astype(np.float64)
        mine = tg_oracle.reference_hazard_table(omega, np.full(R, 2.0), U, 6000)
        product = hazard_table(omega, np.full(R, 2.0), U, 6000, hazard="reference")      # host-only C-ABI call (no GPU needed)
        # the product's native builder against the restatement: the same fp32 values up to the last bit of expf / log1pf
        assert np.allclose(product, mine, rtol=2.5e-6, atol=0)   # one fp32 ulp of a log-hazard near -20
        assert np.array_equal(product == np.float64(np.float32(0.1)), mine == np.float64(np.float32(0.1)))
        exact = tg_oracle.hazard_table(omega, np.full(R, 2.0), U, 6000)
        fixed = np.float64(np.float32(0.1))
        for r in range(R):
            nf = np.nonzero(ref[r] != fixed)[0]
            first = int(nf.max()) + 1
            assert first == want_first[group][r]
            nf2 = np.nonzero(mine[r] != fixed)[0]
            assert abs(int(nf2.max()) + 1 - first) <= 1                  # the product's table switches at the same sojourn
            # the same construction (the reference's omega went through two fp32 inv_logits and a logit, hence 1e-3 near the end)
            assert np.allclose(mine[r][:int(0.5 * first)], ref[r][:int(0.5 * first)], rtol=1e-3, atol=1e-9)
            late = slice(int(0.5 * first), first - 1)   # 1 - cdf is down to a few fp32 bits here: both are noisy around the truth
            assert np.median(np.abs(mine[r][late] - ref[r][late]) / exact[r][late]) < 0.05
            safe = slice(U, int(0.5 * first))                            # well before the fp32 noise of 1 - cdf sets in
            assert np.allclose(ref[r][safe], exact[r][safe], rtol=1e-3)                     # ... where it is the exact hazard
            assert np.all(ref[r][:U] == 0.0) and np.all(mine[r][:U] == 0.0)


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_optimal_finite_state_is_the_reference(g, i):
    """resampling_functions.py:7-69 with the uniform injected (fp32 log-weights, as _filter_one_step passes them)."""
    lw = g[f"ofs{i}_logw"].astype(np.float64)
    parents, log_c, K = tg_oracle.optimal_finite_state(lw, 50, float(g[f"ofs{i}_u"]))
    assert not bool(g[f"ofs{i}_unbiased"]) and parents is not None
    assert np.array_equal(parents, g[f"ofs{i}_parents"])
    assert abs(log_c - float(g[f"ofs{i}_log_c"])) < 1e-4 * max(1.0, abs(log_c))


def test_backward_kernel_is_the_reference(g):
    """smoothing_functions.py:46-59: log w + log f over the finite entries, normalised per row."""
    tm, pw, want = g["bk_trans"].astype(np.float64), g["bk_prev_logw"], g["bk_out"]
    both = np.isfinite(tm) & np.isfinite(pw)[None, :]
    lwf = np.where(both, tm + pw[None, :], -np.inf)
    got = lwf - np.array([tg_oracle.logsumexp(row) for row in lwf])[:, None]
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    f = np.isfinite(want)
    assert np.allclose(got[f], want[f], rtol=1e-12, atol=1e-12)


def _e2e_case(g, tag, hazard):
    from _oracle import Oracle
    from _tg_case import make_case
    T, S = int(g[f"{tag}_T"]), int(g[f"{tag}_S"])
    c = make_case(T, S, seed=int(g[f"{tag}_data_seed"]))
    if hazard == "reference":
        m = c["model"]
        m.rho_c = tg_oracle.reference_hazard_table(c["omega_control"], np.full(R, 2.0), U, m.d_max)
        m.rho_k = tg_oracle.reference_hazard_table(c["omega_case"], np.full(R, 2.0), U, m.d_max)
    o = Oracle()
    lo_c = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"])
    lo_k = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    return c, lo_c, lo_k


@pytest.mark.parametrize("tag,hazard", [("e2e_short", "exact"), ("e2e_long", "reference")])
def test_whole_filter_and_backward_simulation_is_the_reference(g, tag, hazard):
    """filter_and_smoother_algorithm.run of the reference (50 ancestors x 48 proposals, optimal finite-state resampling, 25
    backward trajectories), executed unmodified on the stand-in with this repo's Philox draws injected, against the
    restatement: the sampled trajectories must be the same, site by site and trajectory by trajectory.  The reference computes
    in fp32, the restatement in fp64, so a draw that lands within fp32 rounding of a CDF step may differ: allow 1% (both
    committed cases agree 100%).  e2e_short (60 sites) keeps every sojourn below 94, where the reference's hazard is the exact
    one up to fp32 noise; e2e_long (150 sites) crosses it, so the restatement needs the reference-mode table
    (tg_oracle.reference_hazard_table; the product builds the same table natively, hyg_tg_reference_hazard_table) -- with the exact hazard its log-evidence is off by 3e-3 there."""
    c, lo_c, lo_k = _e2e_case(g, tag, hazard)
    r = tg_oracle.run(c["model"], lo_c, lo_k, M=50, n_backward=25, seed=int(g[f"{tag}_seed"]), chain=int(g[f"{tag}_chain"]))
    assert abs(r["log_norm"] - float(g[f"{tag}_log_norm"])) <= 2e-6 * abs(r["log_norm"])      # fp32 accumulation over T sites
    assert (r["traj_m"] == g[f"{tag}_traj_merged"]).mean() >= 0.99
    assert (r["traj_control"] == g[f"{tag}_traj_control"]).mean() >= 0.99
    assert (r["traj_case"] == g[f"{tag}_traj_case"]).mean() >= 0.99
    assert r["taps"]["n_finite"][-1] == int(g[f"{tag}_n_final_finite"])
