"""Two-group path without a GPU: the NumPy oracle's behaviour, the host-side helpers, and the CUDA kernel's source run
under the CPU emulation of the CUDA execution model (tests/emu) against the oracle."""
import numpy as np
import pytest

from _emu import Emu, tg_run_emu
from _oracle import Oracle
from _tg_case import make_case, tg_oracle
from hygeia_b200 import two_group


@pytest.fixture(scope="module")
def case300():
    c = make_case(300, 3)
    o = Oracle()
    c["lo_c"] = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"])
    c["lo_k"] = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    c["ref"] = tg_oracle.run(c["model"], c["lo_c"], c["lo_k"], M=50, n_backward=25, seed=1, chain=0)
    return c


def test_oracle_recovers_regimes_and_split(case300):
    c, r = case300, case300["ref"]
    assert (r["regime_probs"][:, :6].argmax(1) == c["regimes"]).mean() > 0.9
    assert (r["regime_probs"][:, 6:].argmax(1) == c["reg_case"]).mean() > 0.9
    assert ((r["split_probs"] > 0.5) == (c["regimes"] != c["reg_case"])).mean() > 0.9
    assert np.isfinite(r["log_norm"])
    # particle budget: R^2 initial particles, then at most M (2R + R^2) proposals per site
    assert r["taps"]["n_particles"][0] == 36
    assert r["taps"]["n_particles"].max() <= 50 * 48


def test_trajectories_respect_the_model_constraints(case300):
    r = case300["ref"]
    u = case300["u"]
    # control group: a regime only changes after a sojourn of at least u sites; durations count up from 1
    d, g = r["traj_control"][:, :, 0], r["traj_control"][:, :, 1]
    change = g[1:] != g[:-1]
    assert (d[:-1][change] >= u).all()
    assert (d[1:][change] == 1).all()
    cont = d[1:] != 1
    assert (d[1:][cont] == d[:-1][cont] + 1).all() and (g[1:][cont] == g[:-1][cont]).all()
    # merged sites: the case group carries the control group's state
    merged = r["traj_m"] == 1
    assert (r["traj_case"][merged] == r["traj_control"][merged]).all()
    assert merged[0].all()    # every trajectory starts merged (filter_and_smoother_algorithm.py:141-172)
    # the merged indicator only flips when both sojourns have lasted at least u sites
    flip = r["traj_m"][1:] != r["traj_m"][:-1]
    dmin = np.minimum(r["traj_control"][:-1, :, 0], r["traj_case"][:-1, :, 0])
    assert (dmin[flip] >= u).all()


def test_kernel_under_emulation_matches_oracle(case300):
    c, r = case300, case300["ref"]
    g = tg_run_emu(Emu(), c["model"], c["lo_c"], c["lo_k"], M=50, B=25, seed=1, chain=0)
    assert abs(g["log_norm"] - r["log_norm"]) <= 1e-10 * abs(r["log_norm"])
    assert (g["taps"][:, 0] == r["taps"]["n_particles"]).all()
    assert (g["taps"][:, 1] == r["taps"]["K"]).all()
    assert (g["taps"][:, 2] == r["taps"]["n_finite"]).all()
    assert (g["traj"][:, :, 0] == r["traj_m"]).all()
    assert (g["traj"][:, :, 1:3] == r["traj_control"]).all()
    assert (g["traj"][:, :, 3:5] == r["traj_case"]).all()


@pytest.mark.parametrize("T,M,B,R", [(1, 50, 25, 6), (2, 50, 25, 6), (40, 7, 5, 6), (60, 20, 32, 4)])
def test_kernel_under_emulation_edge_cases(T, M, B, R):
    c = make_case(T, 2, seed=11, R=R)
    o = Oracle()
    lo_c = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"])
    lo_k = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    r = tg_oracle.run(c["model"], lo_c, lo_k, M=M, n_backward=B, seed=3, chain=2)
    g = tg_run_emu(Emu(), c["model"], lo_c, lo_k, M=M, B=B, seed=3, chain=2)
    assert abs(g["log_norm"] - r["log_norm"]) <= 1e-10 * abs(r["log_norm"])
    assert (g["taps"][:, 1] == r["taps"]["K"]).all()
    assert (g["traj"][:, :, 0] == r["traj_m"]).all()
    assert (g["traj"][:, :, 1:3] == r["traj_control"]).all()
    assert (g["traj"][:, :, 3:5] == r["traj_case"]).all()


def test_kernel_sort_preselection_never_changes_the_result():
    """The resampling sort covers only the heaviest particles of a site and is repeated with more of them when a tooth of the
    systematic comb falls behind the sorted prefix.  One-sample data (flat weights) with the smallest allowed prefix forces second
    and third attempts; the outcome must be the oracle's -- which sorts everything -- at every site, whatever the prefix."""
    c = make_case(260, 1, seed=3)
    o = Oracle()
    lo_c = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"])
    lo_k = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    r = tg_oracle.run(c["model"], lo_c, lo_k, M=50, n_backward=25, seed=1, chain=0)
    seen = set()
    for pre in ((0, 0), (1, 1), (1, 300), (100000, 100000)):
        g = tg_run_emu(Emu(), c["model"], lo_c, lo_k, M=50, B=25, seed=1, chain=0, preselect=pre)
        assert abs(g["log_norm"] - r["log_norm"]) <= 1e-10 * abs(r["log_norm"])
        assert (g["taps"][:, 1] == r["taps"]["K"]).all() and (g["taps"][:, 2] == r["taps"]["n_finite"]).all()
        assert (g["traj"][:, :, 0] == r["traj_m"]).all()
        assert (g["traj"][:, :, 1:3] == r["traj_control"]).all() and (g["traj"][:, :, 3:5] == r["traj_case"]).all()
        seen |= set(np.unique(g["taps"][:, 3]).tolist())
        if pre == (100000, 100000):
            assert set(np.unique(g["taps"][2:, 3]).tolist()) <= {0, 1}       # everything sorted at once
    assert {1, 2, 3} <= seen                                                  # every attempt level was exercised


def test_kernel_sort_in_global_scratch_is_the_same_sort():
    """More than 2048 selected particles do not fit the shared-memory sort arrays; the same code then runs on a global scratch
    area.  Forced here for every site (scratch_from = 1), with and without the pre-selection: identical to the oracle."""
    c = make_case(120, 1, seed=3)
    o = Oracle()
    lo_c = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"])
    lo_k = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    r = tg_oracle.run(c["model"], lo_c, lo_k, M=50, n_backward=25, seed=1, chain=0)
    for pre in ((0, 0), (100000, 100000)):
        g = tg_run_emu(Emu(), c["model"], lo_c, lo_k, M=50, B=25, seed=1, chain=0, preselect=pre, scratch_from=1)
        assert abs(g["log_norm"] - r["log_norm"]) <= 1e-10 * abs(r["log_norm"])
        assert (g["taps"][:, 1] == r["taps"]["K"]).all()
        assert (g["traj"][:, :, 0] == r["traj_m"]).all()
        assert (g["traj"][:, :, 1:3] == r["traj_control"]).all() and (g["traj"][:, :, 3:5] == r["traj_case"]).all()


def test_control_group_parameters_match_oracle():
    rng = np.random.default_rng(0)
    theta = rng.normal(size=36)
    lp, om = two_group.control_group_parameters(theta, 6)
    lp_o, logit_o = tg_oracle.control_params_from_theta(theta, 6)
    assert np.array_equal(lp, lp_o)
    assert np.allclose(om, 1 / (1 + np.exp(-logit_o)), rtol=0, atol=0)
    assert np.allclose(np.exp(lp).sum(1), 1.0)
    assert np.all(np.isneginf(np.diag(lp)))


def test_segment_index_windows():
    # run_inference_two_groups.py:194-219 on a 250-site chromosome, segments of 100 with a buffer of 10
    idx, ret = two_group.segment_index(0, 100, 10, 250)
    assert (idx[0], idx[-1]) == (0, 109) and (ret[0], ret[-1]) == (0, 99)
    idx, ret = two_group.segment_index(1, 100, 10, 250)
    assert (idx[0], idx[-1]) == (90, 209) and (idx[ret][0], idx[ret][-1]) == (100, 199)
    idx, ret = two_group.segment_index(2, 100, 10, 250)
    assert (idx[0], idx[-1]) == (190, 249) and (idx[ret][0], idx[ret][-1]) == (200, 249)
    with pytest.raises(two_group.HygeiaError):
        two_group.segment_index(3, 100, 10, 250)


def test_host_hazard_table_matches_oracle():
    # host function of the C ABI (no device needed): same recurrence as the oracle, and the closed form for kappa = 2
    om, ka = [0.995, 0.975, 0.95, 0.925, 0.9, 0.9], [2.0] * 6
    a = two_group.hazard_table(om, ka, 3, 2000)
    b = tg_oracle.hazard_table(om, ka, 3, 2000)
    assert a.shape == b.shape
    assert np.all(a[:, :3] == 0.0)
    np.testing.assert_allclose(a, b, rtol=1e-14, atol=0)
    # closed form for kappa = 2: pmf(k) = (k+1)(1-w)^2 w^k, P(X >= k) = w^k (1 + k (1-w)) -> rho = (k+1)(1-w)^2 / (1 + k(1-w))
    k = np.arange(0, 1998)
    for r, w in enumerate(om):
        np.testing.assert_allclose(a[r, 3:], (k + 1) * (1 - w) ** 2 / (1 + k * (1 - w)), rtol=1e-13)


def test_summarise_counts_trajectories():
    tr = np.zeros((3, 4, 5), np.int32)
    tr[0, :, 0] = [0, 1, 1, 1]
    tr[:, :, 2] = 2
    tr[1, :2, 4] = 5
    split, reg = two_group.summarise(tr, 6)
    assert np.allclose(split, [(tr[t, :, 0] == 0).mean() for t in range(3)])
    assert np.allclose(reg[:, 2], 1.0) and np.isclose(reg[1, 6 + 5], 0.5) and np.isclose(reg[1, 6 + 0], 0.5)
