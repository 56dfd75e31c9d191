"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle and the committed golden
vectors generated from the compiled reference.  Bars (BASELINE.json north_star): emission bit-exact for tabulated counts;
forward log-likelihoods and posterior regime probabilities within 1e-6 relative (fp64); regime calls identical with the
per-site uniforms injected."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # north_star tolerance for fp64


@pytest.fixture(scope="module")
def sess(built):
    from hygeia_b200.single_group import Session
    s = Session(0)
    yield s
    s.close()


def run_chain(sess, vartheta, theta, n_total, n_meth, positions=None, uniforms=None, seed=0, chain_id=0, taps=True, **kw):
    from hygeia_b200.single_group import make_run_args
    S, T = n_total.shape
    R = int(vartheta[1])
    npart = kw.get("n_particles_max", 250)
    sess.clear()
    sess.set_vartheta(vartheta)
    sess.set_theta(theta, T)
    ds = sess.add_dataset(n_total, n_meth)
    out = dict(regime_probs=np.full((T, 1 + R), np.nan), logz=np.zeros(T))
    if taps:
        out.update(k_kept=np.zeros(T, np.int32), drew_uniform=np.zeros(T, np.uint8), n_pending=np.zeros(T, np.int32),
                   n_curr=np.zeros(T, np.int32), finalised_at=np.full(T, -1, np.int32), ancestors=np.full((T, npart - R), -1, np.int16))
    spec = dict(dataset=ds, seed=seed, chain_id=chain_id, uniforms=uniforms, positions=positions, **out)
    sess.set_chains([spec])
    sess.emission()
    sess.filter(make_run_args(**kw))
    out["status"] = sess.download()[0]
    out["logobs"] = sess.get_logobs(ds, T)
    return out


CASES = ["sg_default_s4.npz", "sg_sparse_s1.npz", "sg_dense_s16.npz", "sg_few_particles.npz"]


@pytest.mark.parametrize("case", CASES)
def test_golden_reference_parity(sess, case):
    g = golden(case)
    r = run_chain(sess, g["vartheta"], g["theta"], g["n_total"], g["n_meth"], g["positions"], uniforms=g["uniforms"],
                  n_particles_max=int(g["n_particles"]), epsilon=float(g["epsilon"]))
    assert np.array_equal(r["logobs"], g["ref_strict_logobs"])                       # K1: bit-exact
    assert np.allclose(r["logz"], g["ref_strict_logz"], rtol=RTOL, atol=0)           # forward log-likelihoods
    assert np.max(np.abs(r["logz"] - g["ref_strict_logz"]) / np.abs(g["ref_strict_logz"])) < 1e-10
    want = g["ref_strict_regime_probs"]
    assert np.array_equal(r["regime_probs"][:, 0], want[:, 0])                       # genomic positions
    assert np.allclose(r["regime_probs"][:, 1:], want[:, 1:], rtol=RTOL, atol=1e-12)  # posteriors
    assert np.array_equal(r["regime_probs"][:, 1:].argmax(1), want[:, 1:].argmax(1))  # change-point / regime calls
    assert np.array_equal(r["drew_uniform"], g["ref_strict_drew_uniform"])
    assert np.array_equal(r["n_pending"], g["ref_strict_n_pending"])
    assert np.array_equal(r["finalised_at"], g["ref_strict_finalised_at"])
    # and against the reference built with its own flags (-O3 -ffast-math)
    assert np.allclose(r["regime_probs"][:, 1:], g["ref_fast_regime_probs"][:, 1:], rtol=RTOL, atol=1e-12)
    assert r["status"][0] == 0


def test_step_level_decisions_vs_oracle(sess, oracle, default_model):
    from hygeia_b200 import philox, synthetic
    T, S = 6000, 4
    ch = synthetic.make_chain(T, S, seed=555)
    u = philox.uniforms_by_site(21, 3, T)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u, ch["n_total"], ch["n_meth"], ch["positions"], want_ancestors=True)
    # device-side Philox (no injected array): same draws as the host copy
    got = run_chain(sess, default_model["vartheta"], default_model["theta"], ch["n_total"], ch["n_meth"], ch["positions"], seed=21, chain_id=3)
    assert np.array_equal(got["k_kept"], want["k_kept"])
    assert np.array_equal(got["n_curr"], want["n_curr"])
    assert np.array_equal(got["drew_uniform"], want["drew_uniform"])
    assert np.array_equal(got["finalised_at"], want["finalised_at"])
    same = (got["ancestors"] == want["ancestors"]).all(axis=1).mean()
    assert same > 0.95   # the rest are permutations among exactly tied (zero / equal) weights
    assert np.allclose(got["logz"], want["logz"], rtol=1e-10)
    assert np.allclose(got["regime_probs"], want["regime_probs"], rtol=RTOL, atol=1e-12)
    assert np.array_equal(got["regime_probs"][:, 1:].argmax(1), want["regime_probs"][:, 1:].argmax(1))
    acc = (got["regime_probs"][:, 1:].argmax(1) == ch["regimes"]).mean()
    assert acc > 0.99


def test_operator_mirror_signature(sess, oracle, default_model):
    """hygeia_b200.single_group.run_online_combined_inference == runOnlineCombinedInferenceCpp, argument for argument."""
    from hygeia_b200 import philox, synthetic
    from hygeia_b200.single_group import run_online_combined_inference
    T, S = 2500, 3
    ch = synthetic.make_chain(T, S, seed=8)
    out = run_online_combined_inference(default_model["vartheta"], default_model["theta"], ch["positions"], ch["n_total"], ch["n_meth"],
                                        250, 1, 2, True, 0.01, False, False, True, 200, 0.1, 0.01, False, 11, return_logz=True)
    want = oracle.run(default_model["vartheta"], default_model["theta"], philox.uniforms_by_site(11, 0, T), ch["n_total"], ch["n_meth"], ch["positions"])
    assert out["regimeProbabilityEstimates"].shape == (T, 7)
    assert np.allclose(out["regimeProbabilityEstimates"], want["regime_probs"], rtol=RTOL, atol=1e-12)
    assert np.allclose(out["logZ"], want["logz"], rtol=1e-10)
    assert out["thetaEstimates"] is None and out["cpuTime"] > 0


@pytest.mark.parametrize("T,S", [(1, 1), (2, 3), (7, 2), (41, 1), (43, 5), (1001, 2)])
def test_edge_lengths(sess, oracle, default_model, T, S):
    from hygeia_b200 import philox, synthetic
    ch = synthetic.make_chain(max(T, 50), S, seed=T * 7 + S)
    nt, nm, pos = ch["n_total"][:, :T].copy(), ch["n_meth"][:, :T].copy(), ch["positions"][:T].copy()
    u = philox.uniforms_by_site(1, 0, T)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u, nt, nm, pos)
    got = run_chain(sess, default_model["vartheta"], default_model["theta"], nt, nm, pos, uniforms=u)
    assert np.array_equal(got["logobs"], oracle.emission(*default_model["alpha_beta"], nt, nm))
    assert np.allclose(got["logz"], want["logz"], rtol=1e-10)
    if T > 1:  # T = 1 is outside the reference's domain: its loop never reaches the final-step flag (OnlineCombinedInference.h:74-95)
        assert np.allclose(got["regime_probs"], want["regime_probs"], rtol=RTOL, atol=1e-12)
    else:
        assert np.isclose(got["regime_probs"][0, 1:].sum(), 1.0)
    assert np.array_equal(got["n_curr"], want["n_curr"])


def test_missing_data_and_impossible_counts(sess, oracle, default_model):
    """All-zero coverage contributes nothing (logBB(0,0)=0); x > n gives -inf for that site (misc.h:636-639)."""
    from hygeia_b200 import philox, synthetic
    T, S = 400, 2
    ch = synthetic.make_chain(T, S, seed=2)
    nt, nm = ch["n_total"].copy(), ch["n_meth"].copy()
    nt[:, 100:180] = 0; nm[:, 100:180] = 0         # a stretch without any reads
    nt[0, 300] = 70; nm[0, 300] = 35               # deep coverage, still tabulated
    nt[1, 310] = 3000; nm[1, 310] = 1500           # beyond the table: device lgamma
    u = philox.uniforms_by_site(3, 0, T)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u, nt, nm)
    got = run_chain(sess, default_model["vartheta"], default_model["theta"], nt, nm, uniforms=u)
    lo = oracle.emission(*default_model["alpha_beta"], nt, nm)
    mask = np.ones(T, bool); mask[310] = False
    assert np.array_equal(got["logobs"][mask], lo[mask])
    assert np.allclose(got["logobs"][310], lo[310], rtol=1e-13)
    assert np.abs(got["logobs"][100:180]).max() < 1e-13
    assert np.allclose(got["logz"], want["logz"], rtol=1e-9)
    assert np.allclose(got["regime_probs"], want["regime_probs"], rtol=1e-5, atol=1e-10)   # one site went through device lgamma


def test_batch_of_chains_equals_single_runs(sess, oracle, default_model):
    """Several data sets x seeds in ONE launch (the whole-genome shape) give what separate runs give."""
    from hygeia_b200 import philox, synthetic
    from hygeia_b200.single_group import make_run_args
    lens = [1500, 700, 2300, 90]
    S = 3
    chains = [synthetic.make_chain(T, S, seed=40 + i) for i, T in enumerate(lens)]
    sess.clear()
    sess.set_vartheta(default_model["vartheta"])
    sess.set_theta(default_model["theta"], max(lens))
    specs, outs = [], []
    for i, ch in enumerate(chains):
        ds = sess.add_dataset(ch["n_total"], ch["n_meth"])
        for seed in (0, 1, 2):
            o = dict(regime_probs=np.full((lens[i], 7), np.nan), logz=np.zeros(lens[i]))
            outs.append((i, seed, o))
            specs.append(dict(dataset=ds, seed=seed, chain_id=i, positions=ch["positions"], **o))
    sess.set_chains(specs)
    sess.emission()
    sess.filter(make_run_args())
    st = sess.download()
    assert all(s[0] == 0 for s in st)
    for i, seed, o in outs:
        ch = chains[i]
        want = oracle.run(default_model["vartheta"], default_model["theta"], philox.uniforms_by_site(seed, i, lens[i]),
                          ch["n_total"], ch["n_meth"], ch["positions"])
        assert np.allclose(o["logz"], want["logz"], rtol=1e-10)
        assert np.allclose(o["regime_probs"], want["regime_probs"], rtol=RTOL, atol=1e-12)
    t = sess.timings()
    assert t["emission_launches"] == 1 and t["filter_launches"] >= 1 and t["ms_filter"] > 0


def test_full_size_properties(sess, default_model):
    """BASELINE config-2 shaped chunk (S = 32, long chain): size-independent properties of the outputs."""
    from hygeia_b200 import synthetic
    T, S = 400_000, 32
    rng = np.random.default_rng(9)
    regimes = synthetic.simulate_regimes(T, rng)
    nt, nm = synthetic.simulate_counts(regimes, S, rng)
    a = run_chain(sess, default_model["vartheta"], default_model["theta"], nt, nm, seed=5, taps=False)
    p = a["regime_probs"][:, 1:]
    assert not np.isnan(p).any() and p.min() >= 0.0
    assert np.allclose(p.sum(1), 1.0, atol=1e-9)                       # all R variances tested together => rows sum to 1
    assert np.all(np.diff(a["logz"]) < 50.0) and np.isfinite(a["logz"]).all()
    assert (p.argmax(1) == regimes).mean() > 0.995
    # linearity of the emission in the samples: logObs(S samples) = logObs(first half) + logObs(second half)
    h1 = run_chain(sess, default_model["vartheta"], default_model["theta"], nt[:16], nm[:16], seed=5, taps=False)
    h2 = run_chain(sess, default_model["vartheta"], default_model["theta"], nt[16:], nm[16:], seed=5, taps=False)
    assert np.allclose(a["logobs"], h1["logobs"] + h2["logobs"], rtol=1e-12, atol=1e-9)
    # determinism: same seed, same bits
    b = run_chain(sess, default_model["vartheta"], default_model["theta"], nt, nm, seed=5, taps=False)
    assert np.array_equal(a["regime_probs"], b["regime_probs"]) and np.array_equal(a["logz"], b["logz"])
    # a different seed changes draws but not the calls
    c = run_chain(sess, default_model["vartheta"], default_model["theta"], nt, nm, seed=6, taps=False)
    assert (c["regime_probs"][:, 1:].argmax(1) == p.argmax(1)).mean() > 0.9999


def test_error_behaviour(sess, default_model):
    from hygeia_b200.single_group import HygeiaError, make_run_args
    from hygeia_b200 import synthetic
    ch = synthetic.make_chain(100, 1, seed=1)
    sess.clear()
    sess.set_vartheta(default_model["vartheta"])
    sess.set_theta(default_model["theta"], 100)
    ds = sess.add_dataset(ch["n_total"], ch["n_meth"])
    sess.set_chains([dict(dataset=ds, logz=np.zeros(100))])
    sess.emission()
    with pytest.raises(HygeiaError):
        sess.filter(make_run_args(smc_resample_type=1))      # only optimal finite-state resampling exists
    with pytest.raises(HygeiaError):
        sess.filter(make_run_args(n_particles_max=300))
    with pytest.raises(HygeiaError):
        sess.set_theta(default_model["theta"][:10], 100)
    v = default_model["vartheta"].copy(); v[14] = 0.0          # is_kappa_fixed = FALSE
    with pytest.raises(HygeiaError):
        sess.set_vartheta(v[:15])


def test_parameter_estimation_parity(sess, oracle, default_model):
    """K3 on the device (--estimate_parameters): theta trace, log Z_t and posteriors against the oracle.  Device libm differs
    from glibc by ulps in the table rebuild, hence 1e-6 (north_star tolerance) instead of bit-level agreement."""
    from hygeia_b200 import model, philox, synthetic
    from hygeia_b200.single_group import make_run_args
    T, S = 3000, 3
    ch = synthetic.make_chain(T, S, seed=4242)
    u = philox.uniforms_by_site(9, 0, T)
    theta0 = model.default_theta() + 0.2 * np.random.default_rng(11).standard_normal(36)
    want = oracle.run(default_model["vartheta"], theta0, u, ch["n_total"], ch["n_meth"], ch["positions"], param_est=True)
    sess.clear()
    sess.set_vartheta(default_model["vartheta"])
    sess.set_theta(theta0, T)
    ds = sess.add_dataset(ch["n_total"], ch["n_meth"])
    out = dict(regime_probs=np.full((T, 7), np.nan), logz=np.zeros(T), theta_trace=np.zeros((T, 36)), k_kept=np.zeros(T, np.int32))
    sess.set_chains([dict(dataset=ds, seed=9, chain_id=0, positions=ch["positions"], **out)])
    sess.emission()
    sess.filter(make_run_args(use_online_parameter_estimation=True))
    sess.download()
    assert np.abs(want["theta_trace"][-1] - theta0).max() > 0.05                     # 15 ADAM steps moved theta
    assert np.allclose(out["theta_trace"], want["theta_trace"], rtol=RTOL, atol=1e-9)
    assert np.array_equal(out["theta_trace"][0], theta0)
    assert np.allclose(out["logz"], want["logz"], rtol=RTOL)
    assert np.allclose(out["regime_probs"], want["regime_probs"], rtol=1e-5, atol=1e-9)
    assert (out["k_kept"] == want["k_kept"]).mean() > 0.999
    # the operator mirror returns thetaEstimates exactly like runOnlineCombinedInferenceCpp
    from hygeia_b200.single_group import run_online_combined_inference
    r = run_online_combined_inference(default_model["vartheta"], theta0, ch["positions"], ch["n_total"], ch["n_meth"],
                                      use_online_parameter_estimation=True, rng_seed=9)
    assert r["thetaEstimates"].shape == (T, 36)
    assert np.allclose(r["thetaEstimates"], want["theta_trace"], rtol=RTOL, atol=1e-9)


def test_informative_data_underflow_regime(sess, oracle, default_model):
    """32 samples per site: regime classes underflow in the linear domain (site ~2865 of this chain); filtering and
    parameter estimation must stay finite and on the oracle (regression: 0 * inf in the score recursion, which also
    filled the lag set and slowed the chain down 50x)."""
    from hygeia_b200 import model, philox, synthetic
    from hygeia_b200.single_group import make_run_args
    T, S = 5000, 32
    ch = synthetic.make_chain(T, S, seed=7)
    u = philox.uniforms_by_site(1, 0, T)
    theta0 = model.default_theta()
    for pe in (False, True):
        want = oracle.run(default_model["vartheta"], theta0, u, ch["n_total"], ch["n_meth"], ch["positions"], param_est=pe)
        sess.clear()
        sess.set_vartheta(default_model["vartheta"])
        sess.set_theta(theta0, T)
        ds = sess.add_dataset(ch["n_total"], ch["n_meth"])
        out = dict(regime_probs=np.full((T, 7), np.nan), logz=np.zeros(T), k_kept=np.zeros(T, np.int32))
        if pe:
            out["theta_trace"] = np.zeros((T, 36))
        sess.set_chains([dict(dataset=ds, seed=1, chain_id=0, positions=ch["positions"], **out)])
        sess.emission()
        sess.filter(make_run_args(use_online_parameter_estimation=pe))
        status = sess.download()
        assert status[0][0] == 0                                   # no forced emissions
        assert np.isfinite(out["logz"]).all() and np.isfinite(out["regime_probs"]).all()
        assert np.allclose(out["logz"], want["logz"], rtol=RTOL)
        assert np.allclose(out["regime_probs"], want["regime_probs"], rtol=1e-5, atol=1e-9)
        assert (out["k_kept"] == want["k_kept"]).mean() > 0.999
        if pe:
            assert np.isfinite(out["theta_trace"]).all()
            # ADAM amplifies rounding noise in score components that are mathematically zero (see the emulation test)
            assert np.abs(out["theta_trace"] - want["theta_trace"]).max() < 2e-6


# ---- segmented execution (throughput mode, hyg_sg_set_segmentation) ------------------------------------------------
def _run_two_modes(sess, default_model, chains, segment_sites, halo_left, halo_right, seeds=(3,)):
    """chains: list of synthetic chains.  Returns (whole, segmented): lists of output dicts, one per (chain, seed)."""
    from hygeia_b200.single_group import make_run_args
    res = []
    for seg in (0, segment_sites):
        sess.clear()
        sess.set_segmentation(seg, halo_left, halo_right)
        sess.set_vartheta(default_model["vartheta"])
        sess.set_theta(default_model["theta"], max(c["n_total"].shape[1] for c in chains))
        specs, outs = [], []
        for ci, ch in enumerate(chains):
            ds = sess.add_dataset(ch["n_total"], ch["n_meth"])
            T = ch["n_total"].shape[1]
            for sd in seeds:
                o = dict(regime_probs=np.full((T, 7), np.nan), logz=np.full(T, np.nan), k_kept=np.full(T, -9, np.int32),
                         finalised_at=np.full(T, -1, np.int32), n_curr=np.zeros(T, np.int32))
                outs.append(o)
                specs.append(dict(dataset=ds, seed=sd, chain_id=ci, positions=ch["positions"], **o))
        sess.set_chains(specs)
        sess.emission()
        sess.filter(make_run_args())
        st = sess.download()
        for o, s_ in zip(outs, st):
            o["status"] = s_
        res.append((outs, sess.filter_units(), sess.timings()["ms_filter"]))
    sess.set_segmentation(0)
    return res


@pytest.mark.parametrize("S,lam,T", [(4, 30.0, 60000), (1, 10.0, 40000), (32, 30.0, 30000)])
def test_segmented_matches_whole_chain(sess, oracle, default_model, S, lam, T):
    """Throughput mode against the sequential whole-chain run (itself pinned to the reference by the tests above):
    posteriors and log Z within the north_star tolerance (observed: 1e-10), regime calls identical, every site written
    exactly once, no site forced at a segment end."""
    from hygeia_b200 import synthetic
    chains = [synthetic.make_chain(T, S, seed=90 + S, lam=lam), synthetic.make_chain(T // 3 + 17, S, seed=91 + S, lam=lam)]
    (whole, nu_w, _), (seg, nu_s, _) = _run_two_modes(sess, default_model, chains, segment_sites=8000, halo_left=3000, halo_right=3000,
                                                      seeds=(3, 4))
    assert nu_w == 4 and nu_s == 2 * (-(-T // 8000) + -(-(T // 3 + 17) // 8000))
    for w, s_ in zip(whole, seg):
        Tn = w["logz"].shape[0]
        assert np.isfinite(s_["regime_probs"]).all() and np.isfinite(s_["logz"]).all()
        assert np.array_equal(s_["regime_probs"][:, 0], w["regime_probs"][:, 0])
        assert np.abs(s_["regime_probs"][:, 1:] - w["regime_probs"][:, 1:]).max() < 1e-8
        assert np.array_equal(s_["regime_probs"][:, 1:].argmax(1), w["regime_probs"][:, 1:].argmax(1))
        assert np.allclose(s_["logz"], w["logz"], rtol=1e-9, atol=0)
        assert np.array_equal(s_["finalised_at"], w["finalised_at"])
        assert (s_["k_kept"][8000:] == w["k_kept"][8000:]).mean() > 0.999
        assert s_["status"][0] == 0 and s_["status"][2] == 0
        assert Tn <= w["status"][3] == Tn and Tn < s_["status"][3] < 1.5 * Tn


def test_segmented_single_segment_is_the_whole_chain(sess, default_model):
    """segment_sites >= T: one segment, no halo -> bit-identical to the default mode."""
    from hygeia_b200 import synthetic
    ch = synthetic.make_chain(5000, 2, seed=12)
    (whole, _, _), (seg, nu, _) = _run_two_modes(sess, default_model, [ch], segment_sites=5000, halo_left=100, halo_right=100)
    assert nu == 1
    for k in ("regime_probs", "logz", "k_kept", "finalised_at"):
        assert np.array_equal(whole[0][k], seg[0][k])


def test_segmented_vs_oracle_and_short_halo_detected(sess, oracle, default_model):
    """Against the CPU oracle directly, and: a deliberately short right halo on sparse data must be REPORTED (status[2])."""
    from hygeia_b200 import philox, synthetic
    T = 20000
    ch = synthetic.make_chain(T, 1, seed=5, lam=10.0)
    u = philox.uniforms_by_site(3, 0, T)
    want = oracle.run(default_model["vartheta"], default_model["theta"], u, ch["n_total"], ch["n_meth"], ch["positions"])
    (_, _, _), (seg, _, _) = _run_two_modes(sess, default_model, [ch], segment_sites=2500, halo_left=2500, halo_right=2500)
    assert np.abs(seg[0]["regime_probs"] - want["regime_probs"]).max() < 1e-8
    assert np.allclose(seg[0]["logz"], want["logz"], rtol=1e-9)
    assert seg[0]["status"][2] == 0
    (_, _, _), (short, _, _) = _run_two_modes(sess, default_model, [ch], segment_sites=2500, halo_left=2500, halo_right=2)
    fa = want["finalised_at"]   # a segment [t0, t1) with halo_right = 2 ends at step t1 + 1: later finalisations are forced there
    expected = sum(int((fa[t1 - 2500:t1] > t1 + 1).sum()) for t1 in range(2500, T, 2500))
    assert expected > 10 and short[0]["status"][2] == expected


def test_pinned_outputs_are_written_by_the_kernel(sess, default_model):
    """regime_probs in pinned host memory: K2 streams the rows into it (no staging, no D2H); results identical to the staged
    path (pageable buffers), in whole-chain and in segmented execution."""
    import torch
    from hygeia_b200 import synthetic
    from hygeia_b200.single_group import make_run_args
    T = 30000
    ch = synthetic.make_chain(T, 4, seed=21)
    res = {}
    for seg in (0, 6000):
        for pinned in (False, True):
            sess.clear()
            sess.set_segmentation(seg, 3000, 3000)
            sess.set_vartheta(default_model["vartheta"]); sess.set_theta(default_model["theta"], T)
            ds = sess.add_dataset(ch["n_total"], ch["n_meth"])
            if pinned:
                tp = torch.full((T, 7), float("nan"), dtype=torch.float64).pin_memory()
                probs, ptr = tp.numpy(), tp.data_ptr()
            else:
                probs = np.full((T, 7), np.nan); ptr = probs
            logz = np.zeros(T)
            sess.set_chains([dict(dataset=ds, seed=2, chain_id=0, positions=ch["positions"], regime_probs=ptr, logz=logz)])
            sess.emission(); sess.filter(make_run_args()); sess.download()
            assert np.isfinite(probs).all()
            res[(seg, pinned)] = (probs.copy(), logz.copy())
    sess.set_segmentation(0)
    for seg in (0, 6000):
        assert np.array_equal(res[(seg, True)][0], res[(seg, False)][0])
        assert np.array_equal(res[(seg, True)][1], res[(seg, False)][1])
    assert np.array_equal(res[(0, True)][0][:, 0], ch["positions"].astype(np.float64))
