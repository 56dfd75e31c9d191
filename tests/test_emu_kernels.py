"""Kernel LOGIC on the GPU-less box: the device code of hygeia_b200/csrc compiled under tests/emu/cuda_emu.h
(an emulation of the CUDA execution model) against the oracle.  The GPU tests (-m gpu) are the parity tests proper."""
import numpy as np
import pytest

from conftest import golden


@pytest.fixture(scope="module")
def emu(built):
    from _emu import Emu
    return Emu()


def test_emission_bit_exact(emu, oracle):
    g = golden("sg_default_s4.npz")
    for grid, block in ((1, 32), (3, 64)):
        lo = emu.sg_emission(g["alpha"], g["beta"], g["n_total"], g["n_meth"], grid=grid, block=block)
        assert np.array_equal(lo, g["ref_strict_logobs"])


def test_emission_table_tiers(emu, oracle):
    """Counts beyond the shared-memory rows come from the L2-resident table (still bit-exact); beyond that, device lgamma."""
    rng = np.random.default_rng(1)
    S, T = 3, 301  # odd T exercises the half-filled last site pair
    n = rng.integers(0, 400, size=(S, T)).astype(np.uint16)
    x = (rng.random((S, T)) * (n + 1)).astype(np.uint16)
    x[0, 5] = n[0, 5] + 1  # impossible count -> -inf (misc.h:636-639)
    from hygeia_b200 import model
    al, be = model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA)
    want = oracle.emission(al, be, n, x)
    got = emu.sg_emission(al, be, n, x, nmax_table=255, nmax_smem=20, grid=2, block=64)
    assert np.all(np.isneginf(got[5])) and np.all(np.isneginf(want[5]))
    small = (n <= 255).all(axis=0)
    small[5] = False
    assert np.array_equal(got[small], want[small])
    big = ~small
    big[5] = False
    assert np.allclose(got[big], want[big], rtol=1e-13)


@pytest.mark.parametrize("case,T", [("sg_default_s4.npz", 260), ("sg_sparse_s1.npz", 250), ("sg_dense_s16.npz", 200)])
def test_filter_matches_golden_reference(emu, case, T):
    g = golden(case)
    lo = g["ref_strict_logobs"][:T]
    r = emu.sg_filter(g["vartheta"], g["theta"], lo, uniforms=g["uniforms"][:T], n_particles=int(g["n_particles"]),
                      epsilon=float(g["epsilon"]))
    # sites finalised before the truncation point are final-step independent
    done = g["ref_strict_finalised_at"][:T] < T - 1
    assert done.sum() > T // 2
    assert np.allclose(r["logz"], g["ref_strict_logz"][:T], rtol=1e-12)
    assert np.array_equal(r["drew_uniform"], g["ref_strict_drew_uniform"][:T])
    assert np.array_equal(r["n_pending"][:-1], g["ref_strict_n_pending"][:T - 1])
    want = g["ref_strict_regime_probs"][:T, 1:]
    assert np.allclose(r["probs"][done], want[done], rtol=1e-6, atol=1e-12)
    assert np.array_equal(r["probs"][done].argmax(1), want[done].argmax(1))


def test_filter_step_level_vs_oracle(emu, oracle, default_model):
    from hygeia_b200 import philox, synthetic
    T, S = 320, 2
    ch = synthetic.make_chain(T, S, seed=31, lam=15.0)
    al, be = default_model["alpha_beta"]
    lo = oracle.emission(al, be, ch["n_total"], ch["n_meth"])
    u = philox.uniforms_by_site(4, 9, T)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u, logobs=lo, tie_order="canonical")
    got = emu.sg_filter(default_model["vartheta"], default_model["theta"], lo, uniforms=None, seed=4, chain_id=9)  # device Philox
    assert np.array_equal(got["k_kept"], want["k_kept"])
    assert np.array_equal(got["n_curr"], want["n_curr"])
    assert np.array_equal(got["finalised_at"], want["finalised_at"])
    assert np.allclose(got["logz"], want["logz"], rtol=1e-12)
    assert np.allclose(got["probs"], want["regime_probs"][:, 1:], rtol=1e-6, atol=1e-12)
    assert got["status"][0] == 0 and got["status"][1] == want["n_pending"].max()


@pytest.mark.parametrize("S,lam,pm,T", [(1, 8.0, 0.2, 1200), (32, 30.0, 0.05, 700)])
def test_filter_packed_sort_equals_exact_sort_and_oracle(emu, oracle, default_model, S, lam, pm, T):
    """One-sample sparse data (exact ties, long lag sets) and 32-sample data (weights underflow to zero): the packed 64-bit
    sort with its exact re-sort on near-ties, the exact sort at every site, and the oracle with the canonical tie order take
    the same decisions at every site."""
    from hygeia_b200 import philox, synthetic
    ch = synthetic.make_chain(T, S, seed=12, lam=lam, p_missing=pm)
    al, be = default_model["alpha_beta"]
    lo = oracle.emission(al, be, ch["n_total"], ch["n_meth"])
    u = philox.uniforms_by_site(12, 0, T)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u, logobs=lo, tie_order="canonical")
    fast = emu.sg_filter(default_model["vartheta"], default_model["theta"], lo, uniforms=u, lcap=128)
    full = emu.sg_filter(default_model["vartheta"], default_model["theta"], lo, uniforms=u, lcap=128, force_full_sort=True)
    sorted_sites = int((want["k_kept"] != -1).sum()) - int((np.diff(want["n_curr"]) == 0).sum() - (want["k_kept"] != -1).sum() < 0)
    assert fast["status"][4] < 0.3 * sorted_sites and full["status"][4] >= (want["k_kept"] >= 0).sum()
    for got in (fast, full):
        for k in ("k_kept", "support_hash", "drew_uniform", "n_pending", "finalised_at"):
            assert np.array_equal(got[k], want[k]), k
        assert np.array_equal(got["tie_flags"] & 2, want["tie_flags"] & 2)
        assert np.allclose(got["logz"], want["logz"], rtol=1e-13)
        assert np.abs(got["probs"] - want["regime_probs"][:, 1:]).max() < 1e-11
        assert got["status"][5] == int((want["tie_flags"] & 2 > 0).sum()) and got["status"][6] == 0
    if S == 1:
        assert fast["status"][5] > 0 and fast["status"][1] > 20      # ties decided fates; the lag set was long


def test_filter_lag_capacity_overflow_is_reported(emu, oracle, default_model):
    g = golden("sg_sparse_s1.npz")
    T = 200
    r = emu.sg_filter(g["vartheta"], g["theta"], g["ref_strict_logobs"][:T], uniforms=g["uniforms"][:T], lcap=4)
    assert r["status"][1] <= 4
    assert r["status"][0] > 0          # forced emissions are counted, never silent
    assert not np.isnan(r["probs"]).any()


def test_parameter_estimation_under_emulation(emu, oracle, default_model):
    """K3 (score recursion + ADAM + on-device table rebuild) against the oracle, which is pinned to the reference."""
    from hygeia_b200 import model, philox, synthetic
    T, S = 330, 2
    ch = synthetic.make_chain(T, S, seed=77)
    al, be = default_model["alpha_beta"]
    lo = oracle.emission(al, be, ch["n_total"], ch["n_meth"])
    u = philox.uniforms_by_site(1, 0, T)
    theta0 = model.default_theta() + 0.2 * np.random.default_rng(3).standard_normal(36)
    for kw in (dict(adam=True), dict(adam=False, normalise=True, lr_factor=0.05)):
        want = oracle.run(default_model["vartheta"], theta0, u, logobs=lo, param_est=True, n_steps_without_update=50, **kw, tie_order="canonical")
        got = emu.sg_filter(default_model["vartheta"], theta0, lo, uniforms=u, param_est=True, n_steps_without_update=50, **kw)
        assert np.abs(want["theta_trace"][-1] - theta0).max() > 1e-3           # theta moved
        assert np.allclose(got["theta_trace"], want["theta_trace"], rtol=1e-9, atol=1e-12)
        assert np.allclose(got["logz"], want["logz"], rtol=1e-12)
        assert np.array_equal(got["k_kept"], want["k_kept"])
        assert np.allclose(got["probs"], want["regime_probs"][:, 1:], rtol=1e-6, atol=1e-12)


def test_parameter_estimation_underflow_regime(emu, oracle, default_model):
    """Informative data (32 samples): all but one regime class underflow in the linear domain around site 2865, where the
    new-segment particles must take the log-domain path -- in the weights AND in the score recursion (regression: phi
    became 0 * inf = NaN).  Components of the score that are mathematically zero are rounding noise in both
    implementations and ADAM turns that noise into steps of up to lr * |noise| / 1e-8, hence the absolute tolerance."""
    from hygeia_b200 import model, philox, synthetic
    T, S = 3100, 32
    ch = synthetic.make_chain(T, S, seed=7)
    al, be = default_model["alpha_beta"]
    lo = oracle.emission(al, be, ch["n_total"], ch["n_meth"])
    u = philox.uniforms_by_site(1, 0, T)
    theta0 = model.default_theta()
    want = oracle.run(default_model["vartheta"], theta0, u, logobs=lo, param_est=True, tie_order="canonical")
    got = emu.sg_filter(default_model["vartheta"], theta0, lo, uniforms=u, param_est=True, lcap=128)
    assert np.isfinite(got["theta_trace"]).all() and np.isfinite(got["logz"]).all()
    assert np.abs(got["theta_trace"] - want["theta_trace"]).max() < 1e-6
    assert np.allclose(got["logz"], want["logz"], rtol=1e-9)
    assert (got["k_kept"] == want["k_kept"]).mean() > 0.999
    assert got["status"][0] == 0      # no forced emissions: the lag set never filled up


@pytest.mark.parametrize("S,lam", [(1, 10.0), (4, 30.0)])
def test_filter_segment_view_mechanics(emu, oracle, default_model, S, lam):
    """Segmented execution (hyg_sg_set_segmentation): a descriptor that covers sites [t_off, t_off + T) and owns a
    sub-range must reproduce, on the owned rows, the plain algorithm run on that slice -- global Philox counters, log Z
    relative to the site before the owned range, early exit once no owned site is pending, unowned rows untouched."""
    from hygeia_b200 import philox, synthetic
    T_full, a, lo_, hi_ = 700, 100, 150, 450            # slice starts at global site 100, owns global sites [250, 550)
    ch = synthetic.make_chain(T_full, S, seed=77, lam=lam)
    al, be = default_model["alpha_beta"]
    lo = oracle.emission(al, be, ch["n_total"], ch["n_meth"])
    u = philox.uniforms_by_site(5, 2, T_full)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u[a:], logobs=lo[a:], tie_order="canonical")
    got = emu.sg_filter(default_model["vartheta"], default_model["theta"], lo[a:], uniforms=None, seed=5, chain_id=2,
                        t_off=a, own=(lo_, hi_), last_segment=False)
    steps = int(got["status"][3])
    assert hi_ <= steps < T_full - a, "the run must stop early, after the owned range"
    own = slice(lo_, hi_)
    assert (want["finalised_at"][own] < steps).all() and got["status"][2] == 0
    assert np.array_equal(got["finalised_at"][own], want["finalised_at"][own] + a)
    assert np.array_equal(got["k_kept"][own], want["k_kept"][own])
    assert np.array_equal(got["drew_uniform"][own], want["drew_uniform"][own])
    assert np.allclose(got["probs"][own], want["regime_probs"][own, 1:], rtol=1e-9, atol=1e-13)
    assert np.allclose(got["logz"][own], want["logz"][own] - want["logz"][lo_ - 1], rtol=0, atol=1e-9)
    assert abs(got["seg_inc"][0] - (want["logz"][hi_ - 1] - want["logz"][lo_ - 1])) < 1e-9
    # nothing outside the owned range is written
    assert np.isnan(got["probs"][:lo_]).all() and np.isnan(got["probs"][hi_:]).all()
    assert (got["finalised_at"][:lo_] == -1).all() and (got["finalised_at"][hi_:] == -1).all()
    assert (got["logz"][:lo_] == 0).all() and (got["logz"][hi_:] == 0).all()


def test_filter_segment_right_halo_forced_is_counted(emu, oracle, default_model):
    """If the right halo ends while owned sites are still pending they are emitted by force and counted in status[2]."""
    g = golden("sg_sparse_s1.npz")
    lo = g["ref_strict_logobs"]
    pend = g["ref_strict_n_pending"]
    t = int(np.argmax(pend[100:400] >= 2)) + 100          # a step with >= 2 pending sites
    r = emu.sg_filter(g["vartheta"], g["theta"], lo[:t + 1], uniforms=g["uniforms"][:t + 1], own=(0, t + 1), last_segment=False)
    assert r["status"][2] == pend[t] and r["status"][3] == t + 1
    r = emu.sg_filter(g["vartheta"], g["theta"], lo[:t + 1], uniforms=g["uniforms"][:t + 1])
    assert r["status"][2] == 0


@pytest.mark.parametrize("mu,sigma,omega,u,n_particles", [
    ((0.9, 0.1, 0.5, 0.5), (0.05, 0.05, 0.1, 0.2886751), (0.99, 0.97, 0.95, 0.9), 5, 250),
    ((0.85, 0.15), (0.08, 0.08), (0.98, 0.98), 2, 40),
    ((0.8, 0.2, 0.5), (0.1, 0.1, 0.2), (0.97, 0.97, 0.9), 4, 64),
])
def test_filter_other_numbers_of_regimes_and_minimum_durations(emu, oracle, mu, sigma, omega, u, n_particles):
    """The recursion kernel is instantiated for R = 2 .. 6 (hyg_api.cu dispatches on R); every instantiation against the oracle."""
    from hygeia_b200 import model, philox, synthetic
    R = len(mu)
    vartheta, dim = model.get_known_parameters(mu, sigma, u=u)
    theta = model.default_theta(R, omega=omega)
    T = 1200
    rng = np.random.default_rng(40 + R)
    reg = synthetic.simulate_regimes(T, rng) % R
    nt, nm = synthetic.simulate_counts(reg, 2, rng, mu=mu, sigma=sigma)
    un = philox.uniforms_by_site(5, 0, T)
    want = oracle.run(vartheta, theta, un, nt, nm, None, tie_order="canonical", n_particles=n_particles)
    al, be = model.beta_parameters(mu, sigma)
    got = emu.sg_filter(vartheta, theta, oracle.emission(al, be, nt, nm), uniforms=un, n_particles=n_particles)
    for k in ("drew_uniform", "n_pending", "n_curr", "finalised_at", "support_hash"):
        assert np.array_equal(got[k], want[k]), k
    assert (got["k_kept"] != want["k_kept"]).sum() <= 1
    assert np.allclose(got["logz"], want["logz"], rtol=1e-10, atol=0)
    assert np.allclose(got["probs"], want["regime_probs"][:, 1:], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("mu,sigma,omega,u", [((0.9, 0.1, 0.5, 0.5), (0.05, 0.05, 0.1, 0.2886751), (0.99, 0.97, 0.95, 0.9), 3),
                                              ((0.85, 0.15, 0.5), (0.08, 0.08, 0.2), (0.98, 0.98, 0.9), 2)])
def test_parameter_estimation_other_numbers_of_regimes(emu, oracle, mu, sigma, omega, u):
    """K3 (score recursion + ADAM + table rebuild on the device) is instantiated per R as well: theta traces against the oracle."""
    from hygeia_b200 import model, philox, synthetic
    R = len(mu)
    vartheta, dim = model.get_known_parameters(mu, sigma, u=u)
    theta0 = model.default_theta(R, omega=omega) + 0.3 * np.random.default_rng(1).standard_normal(dim)
    T = 900
    rng = np.random.default_rng(50 + R)
    reg = synthetic.simulate_regimes(T, rng) % R
    nt, nm = synthetic.simulate_counts(reg, 2, rng, mu=mu, sigma=sigma)
    un = philox.uniforms_by_site(5, 0, T)
    want = oracle.run(vartheta, theta0, un, nt, nm, None, tie_order="canonical", param_est=True, n_steps_without_update=50)
    al, be = model.beta_parameters(mu, sigma)
    got = emu.sg_filter(vartheta, theta0, oracle.emission(al, be, nt, nm), uniforms=un, param_est=True, n_steps_without_update=50)
    assert np.abs(want["theta_trace"][-1] - theta0).max() > 1e-2          # theta moved
    assert np.allclose(got["theta_trace"], want["theta_trace"], rtol=0, atol=1e-9)
    assert np.allclose(got["logz"], want["logz"], rtol=1e-10, atol=0)
