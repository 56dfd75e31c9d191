"""GPU parity of the two-group path (K1 emission tables + K4 filter + K5 backward simulation) against oracle/tg_oracle.py,
through the C ABI (hyg_sg_add_dataset / hyg_sg_emission / hyg_tg_set_model / hyg_tg_run)."""
import numpy as np
import pytest

from _oracle import Oracle
from _tg_case import make_case, tg_oracle

pytestmark = pytest.mark.gpu


def _run_gpu(c, M, B, seed, chain, use_oracle_tables=True, want_taps=True, n_copies=1, hazard="exact", sort_preselect=(0, 0), sort_scratch_from=0):
    from hygeia_b200.two_group import TwoGroupSession
    s = TwoGroupSession(0)
    try:
        s.set_emission_model(c["mu"], c["sigma"], c["u"])
        for _ in range(n_copies):
            s.add_dataset(c["nt_c"], c["nm_c"])
            s.add_dataset(c["nt_k"], c["nm_k"])
        s.emission()
        m = c["model"]
        kw = dict(rho_control=m.rho_c, rho_case=m.rho_k) if use_oracle_tables else {}
        s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], c["u"], M, B, t_max=c["T"], hazard=hazard, sort_preselect=sort_preselect, sort_scratch_from=sort_scratch_from, **kw)
        specs = [dict(control_dataset=2 * i, case_dataset=2 * i + 1, T=c["T"], seed=seed, chain_id=chain + i) for i in range(n_copies)]
        return s.run(specs, want_taps=want_taps)
    finally:
        s.close()


def _oracle(c, M, B, seed, chain):
    o = Oracle()
    lo_c = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"])
    lo_k = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    return tg_oracle.run(c["model"], lo_c, lo_k, M=M, n_backward=B, seed=seed, chain=chain)


def _compare(g, r, frac=0.98):
    # fp64 throughout; device exp/log differ from the host's by an ulp, so weights agree to ~1e-13 and the discrete decisions
    # (K, ancestors, backward draws) agree except where a comparison is decided by that last ulp
    assert abs(g["log_normalizing_constant"] - r["log_norm"]) <= 1e-9 * abs(r["log_norm"])
    taps = g["taps"]
    assert (taps[:, 0] == r["taps"]["n_particles"]).all()
    assert (taps[:, 1] == r["taps"]["K"]).mean() >= frac
    assert (taps[:, 2] == r["taps"]["n_finite"]).mean() >= frac
    tr = g["trajectories"]
    assert (tr[:, :, 0] == r["traj_m"]).mean() >= frac
    assert (tr[:, :, 1:3] == r["traj_control"]).mean() >= frac
    assert (tr[:, :, 3:5] == r["traj_case"]).mean() >= frac


def test_two_group_parity_default():
    c = make_case(400, 3)
    g = _run_gpu(c, 50, 25, seed=1, chain=0)[0]
    r = _oracle(c, 50, 25, seed=1, chain=0)
    _compare(g, r)
    # and it recovers the simulated truth
    from hygeia_b200.two_group import summarise
    split, reg = summarise(g["trajectories"], 6)
    assert (reg[:, :6].argmax(1) == c["regimes"]).mean() > 0.9
    assert ((split > 0.5) == (c["regimes"] != c["reg_case"])).mean() > 0.9


@pytest.mark.parametrize("T,M,B,R", [(1, 50, 25, 6), (2, 50, 25, 6), (40, 7, 5, 6), (60, 20, 32, 4), (150, 64, 25, 5)])
def test_two_group_parity_edge_cases(T, M, B, R):
    c = make_case(T, 2, seed=11, R=R)
    g = _run_gpu(c, M, B, seed=3, chain=2)[0]
    r = _oracle(c, M, B, seed=3, chain=2)
    _compare(g, r, frac=0.95)


def test_two_group_device_hazard_tables():
    # tables built by the library itself (no injected hazards) give the same answer
    c = make_case(200, 2, seed=4)
    g = _run_gpu(c, 50, 25, seed=9, chain=1, use_oracle_tables=False, hazard="exact")[0]
    r = _oracle(c, 50, 25, seed=9, chain=1)
    _compare(g, r, frac=0.95)


def test_two_group_reference_mode_hazard_tables():
    # the default of `hygeia infer`: the hazard as the reference's fp32 code evaluates it, fixed value 0.1 from d = 94 (case group)
    c = make_case(260, 2, seed=4)
    m = c["model"]
    m.rho_c = tg_oracle.reference_hazard_table(c["omega_control"], np.full(6, 2.0), c["u"], m.d_max)
    m.rho_k = tg_oracle.reference_hazard_table(c["omega_case"], np.full(6, 2.0), c["u"], m.d_max)
    g = _run_gpu(c, 50, 25, seed=9, chain=1, use_oracle_tables=False, hazard="reference")[0]
    r = _oracle(c, 50, 25, seed=9, chain=1)
    _compare(g, r, frac=0.95)
    e = _run_gpu(c, 50, 25, seed=9, chain=1, use_oracle_tables=False, hazard="exact")[0]
    assert e["log_normalizing_constant"] != g["log_normalizing_constant"]     # the two modes are different models (2.8e-7 apart here)


@pytest.mark.parametrize("tag,hazard", [("e2e_short", "exact"), ("e2e_long", "reference")])
def test_two_group_against_the_reference_run(tag, hazard):
    """tests/golden/tg_reference.npz e2e_*: the reference's own filter_and_smoother_algorithm.run (unmodified, on oracle/shim_tf,
    this repo's Philox draws injected; tests/golden/make_golden_tg.py) -- the CUDA path samples the same trajectories."""
    from conftest import golden
    gd = golden("tg_reference.npz")
    c = make_case(int(gd[f"{tag}_T"]), int(gd[f"{tag}_S"]), seed=int(gd[f"{tag}_data_seed"]))
    if hazard == "reference":
        # the restatement's table: the library's own builder agrees with it to one fp32 ulp (tests/test_two_group_reference.py), but the
        # reference's hazard is fp32 noise around the truth and a last-bit difference can flip a resampling decision 100 sites on
        m = c["model"]
        m.rho_c = tg_oracle.reference_hazard_table(c["omega_control"], np.full(6, 2.0), c["u"], m.d_max)
        m.rho_k = tg_oracle.reference_hazard_table(c["omega_case"], np.full(6, 2.0), c["u"], m.d_max)
    g = _run_gpu(c, 50, 25, seed=int(gd[f"{tag}_seed"]), chain=int(gd[f"{tag}_chain"]), use_oracle_tables=(hazard == "reference"), hazard=hazard)[0]
    ln = float(gd[f"{tag}_log_norm"])
    assert abs(g["log_normalizing_constant"] - ln) <= 2e-6 * abs(ln)                       # the reference accumulates in fp32
    tr = g["trajectories"]
    assert (tr[:, :, 0] == gd[f"{tag}_traj_merged"]).mean() >= 0.99
    assert (tr[:, :, 1:3] == gd[f"{tag}_traj_control"]).mean() >= 0.99
    assert (tr[:, :, 3:5] == gd[f"{tag}_traj_case"]).mean() >= 0.99
    assert g["taps"][-1, 2] == int(gd[f"{tag}_n_final_finite"])


def test_two_group_sort_preselection_never_changes_the_result():
    # one-sample data + the smallest allowed sorted prefix: second and third sort attempts occur, the outcome stays the same
    c = make_case(400, 1, seed=3)
    base = _run_gpu(c, 50, 25, seed=1, chain=0, sort_preselect=(100000, 100000))[0]      # everything sorted at once
    assert set(np.unique(base["taps"][2:, 3]).tolist()) <= {0, 1}
    seen = set()
    for pre in ((0, 0), (1, 1), (1, 300)):
        g = _run_gpu(c, 50, 25, seed=1, chain=0, sort_preselect=pre)[0]
        assert g["log_normalizing_constant"] == base["log_normalizing_constant"]
        assert np.array_equal(g["trajectories"], base["trajectories"]) and np.array_equal(g["taps"][:, :3], base["taps"][:, :3])
        seen |= set(np.unique(g["taps"][:, 3]).tolist())
    assert {1, 2, 3} <= seen
    _compare(base, _oracle(c, 50, 25, seed=1, chain=0), frac=0.95)
    # the sort in the global scratch area (what more than 2048 selected particles would take), forced for every site
    for pre in ((0, 0), (100000, 100000)):
        g = _run_gpu(c, 50, 25, seed=1, chain=0, sort_preselect=pre, sort_scratch_from=1)[0]
        assert g["log_normalizing_constant"] == base["log_normalizing_constant"]
        assert np.array_equal(g["trajectories"], base["trajectories"]) and np.array_equal(g["taps"][:, :3], base["taps"][:, :3])


def test_two_group_trajectories_into_pinned_host_memory():
    """Caller-owned trajectory arrays: pageable ones are filled by a copy after the launch, page-locked ones are written by the
    backward pass itself (no staging copy).  Same numbers either way."""
    import torch
    from hygeia_b200.two_group import TwoGroupSession
    c = make_case(300, 2, seed=6)
    s = TwoGroupSession(0)
    try:
        s.set_emission_model(c["mu"], c["sigma"], c["u"])
        s.add_dataset(c["nt_c"], c["nm_c"]); s.add_dataset(c["nt_k"], c["nm_k"])
        s.emission()
        s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], c["u"], 50, 25, t_max=c["T"])
        spec = dict(control_dataset=0, case_dataset=1, T=c["T"], seed=3, chain_id=0)
        a = s.run([dict(spec)])[0]
        pinned = torch.full((c["T"], 25, 5), -7, dtype=torch.int32).pin_memory()
        b = s.run([dict(spec, trajectories=pinned.numpy())])[0]
        pageable = np.full((c["T"], 25, 5), -7, dtype=np.int32)
        d = s.run([dict(spec, trajectories=pageable)])[0]
    finally:
        s.close()
    assert b["trajectories"] is not None and np.array_equal(pinned.numpy(), a["trajectories"]) and np.array_equal(pageable, a["trajectories"])
    assert a["log_normalizing_constant"] == b["log_normalizing_constant"] == d["log_normalizing_constant"]
    assert (a["trajectories"][:, :, 0] >= 0).all()


def test_two_group_at_the_shape_of_baseline_config_3():
    """50 + 50 samples (emission differences of hundreds of nats between regimes) and a window long enough for sojourns beyond the
    sojourns where the reference-mode hazard of the control regimes turns into the constant 0.1 (197 ... 808)."""
    c = make_case(1500, 50, seed=13)
    m = c["model"]
    m.rho_c = tg_oracle.reference_hazard_table(c["omega_control"], np.full(6, 2.0), c["u"], m.d_max)
    m.rho_k = tg_oracle.reference_hazard_table(c["omega_case"], np.full(6, 2.0), c["u"], m.d_max)
    g = _run_gpu(c, 50, 25, seed=2, chain=5, use_oracle_tables=False, hazard="reference")[0]
    r = _oracle(c, 50, 25, seed=2, chain=5)
    _compare(g, r, frac=0.95)
    assert g["trajectories"][:, :, 1].max() > 200       # control sojourns past the first switch-over


def test_two_group_many_chains_are_independent():
    # 3 chains (same data, different chain ids) in one launch: each equals its own oracle run
    c = make_case(120, 2, seed=8)
    gs = _run_gpu(c, 50, 25, seed=5, chain=10, n_copies=3)
    for i, g in enumerate(gs):
        r = _oracle(c, 50, 25, seed=5, chain=10 + i)
        _compare(g, r, frac=0.95)
    assert gs[0]["log_normalizing_constant"] != gs[1]["log_normalizing_constant"] or \
        not np.array_equal(gs[0]["trajectories"], gs[1]["trajectories"])


def test_infer_mirror_outputs():
    from hygeia_b200 import two_group
    c = make_case(300, 3, seed=2)
    out = two_group.infer(c["nt_c"].T, c["nm_c"].T, c["nt_k"].T, c["nm_k"].T, c["theta"], seed=7)
    T = c["T"]
    assert out["backward_particles_merged_state"].shape == (T, 25) and out["backward_particles_merged_state"].dtype == np.int16
    assert out["backward_particles_control_state"].shape == (T, 25, 2)
    assert out["backward_particles_case_state"].shape == (T, 25, 2)
    assert out["split_probs"].shape == (T,) and out["regime_probs"].shape == (T, 12)
    assert list(out["log_normalizing_constants_optimal"].keys()) == [2400]
    np.testing.assert_allclose(out["regime_probs"][:, :6].sum(1), 1.0, atol=1e-6)
    assert (out["regime_probs"][:, :6].argmax(1) == c["regimes"]).mean() > 0.9
    # same call again: counter-based draws make it reproducible
    out2 = two_group.infer(c["nt_c"].T, c["nm_c"].T, c["nt_k"].T, c["nm_k"].T, c["theta"], seed=7)
    assert np.array_equal(out["backward_particles_control_state"], out2["backward_particles_control_state"])
    assert out["log_normalizing_constants_optimal"] == out2["log_normalizing_constants_optimal"]


def test_two_group_argument_errors():
    from hygeia_b200.two_group import TwoGroupSession, HygeiaError
    c = make_case(50, 2)
    s = TwoGroupSession(0)
    try:
        s.set_emission_model(c["mu"], c["sigma"], 3)
        with pytest.raises(HygeiaError):
            s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], 3, num_resampled=65)
        with pytest.raises(HygeiaError):
            s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], 3, num_backward=33)
        with pytest.raises(HygeiaError):
            s.set_two_group_model(c["logP"], c["omega_control"], np.full(6, 1.5), 3)
        s.set_two_group_model(c["logP"], c["omega_control"], c["omega_case"], 3)
        s.add_dataset(c["nt_c"], c["nm_c"])
        s.add_dataset(c["nt_k"][:, :40], c["nm_k"][:, :40])
        s.emission()
        with pytest.raises(HygeiaError):     # control and case differ in length
            s.run([dict(control_dataset=0, case_dataset=1, T=50)])
        with pytest.raises(HygeiaError):     # data set index out of range
            s.run([dict(control_dataset=0, case_dataset=5, T=50)])
    finally:
        s.close()
