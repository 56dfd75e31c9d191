"""Host-side mirror of the reference's two-group (case/control) inference call, on top of the C ABI.

What is mirrored (/root/reference/src/two_group):

* ``control_group_parameters``  -> get_estimated_control_group_param, run_inference_two_groups.py:76-89
* ``segment_index``             -> the --batch / --segment_size / --buffer_size windowing, run_inference_two_groups.py:194-219
* ``infer``                     -> run_non_marginal_pf_optimal + the test functions and the arrays written to disk,
                                   run_inference_two_groups.py:261-322 (= hygeia/filter_and_smoother_algorithm.py::run with
                                   optimal_resampling=True, multinomial_resampling=False)

The heavy lifting (emission tables, particle filter, backward simulation) happens in the CUDA library; nothing here
computes on the CPU apart from the O(T x B) tabulation of the returned trajectories into split / regime frequencies.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .model import DEFAULT_MU, DEFAULT_SIGMA, beta_parameters
from .single_group import HygeiaError, Session, _ptr


def control_group_parameters(theta, n_regimes):
    """theta of the single-group run (R(R-1) transition logits, then R sojourn logits) -> (log P, omega_control).

    log P: rows are exp(theta) normalised over the off-diagonal entries (run_inference_two_groups.py:80-87).
    omega_control: the reference passes sigmoid(sigmoid(theta_omega)) to the model (:145-150) which then builds the negative
    binomial with probs = logit(.) (case_control_regime_model.py:117-119), i.e. probs = sigmoid(theta_omega); the fp32
    round trip is not reproduced."""
    theta = np.asarray(theta, dtype=np.float64)
    R = int(n_regimes)
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
    omega = 1.0 / (1.0 + np.exp(-theta[-R:]))
    return logp, omega


HAZARD_MODES = {"reference": 0, "exact": 1}   # HYG_TG_HAZARD_REFERENCE / HYG_TG_HAZARD_EXACT


def segment_index(batch, segment_size, buffer_size, n_sites):
    """(index, return_index) of run_inference_two_groups.py:194-219: the window that is filtered and the part of it kept."""
    if batch * segment_size > n_sites:
        raise HygeiaError("Batch index is too large for the chromosome")
    lo = max(0, batch * segment_size - buffer_size)
    hi = min((batch + 1) * segment_size + buffer_size, n_sites)
    n = hi - lo
    if batch == 0:
        ret = np.arange(0, min(n, segment_size))
    else:
        ret = np.arange(buffer_size, min(n, buffer_size + segment_size))
    return np.arange(lo, hi), ret


class TwoGroupSession(Session):
    """Batched form: many (segment, seed) chains in one launch."""

    def set_emission_model(self, mu=DEFAULT_MU, sigma=DEFAULT_SIGMA, minimum_duration=3):
        alpha, beta = beta_parameters(mu, sigma)
        R = len(alpha)
        self.R = R
        a = np.ascontiguousarray(alpha); b = np.ascontiguousarray(beta); k = np.full(R, 2.0)
        self._check(self.lib.hyg_sg_set_model(self.ctx, R, int(minimum_duration), _ptr(a), _ptr(b), 1, _ptr(k)), "hyg_sg_set_model")

    def set_two_group_model(self, log_p_control, omega_control, omega_case, minimum_duration=3, num_resampled=50, num_backward=25,
                            merge_prob=0.1, split_prob=0.01, kappa_control=None, kappa_case=None, rho_control=None, rho_case=None,
                            t_max=4096, hazard="reference", sort_preselect=(0, 0), sort_scratch_from=0):
        """hazard: "reference" (default) = the hazard as the reference's fp32 TensorFlow code evaluates it, with the fixed value
        0.1 from the sojourn where its fp32 cdf rounds to 1 (d = 94 for omega = 0.8; case_control_regime_model.py:111-168);
        "exact" = the negative-binomial hazard in fp64.  Ignored when rho_control / rho_case tables are supplied."""
        if hazard not in HAZARD_MODES:
            raise HygeiaError(f"hazard must be one of {sorted(HAZARD_MODES)}")
        R = self.R
        m = _lib.HygTgModel()
        keep = []

        def arr(x, n):
            if x is None:
                return None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)) if np.ndim(x) <= 1 else np.asarray(x, dtype=np.float64))
            keep.append(a)
            return a.ctypes.data

        m.R, m.minimum_duration, m.num_resampled, m.num_backward = R, int(minimum_duration), int(num_resampled), int(num_backward)
        lp = np.ascontiguousarray(log_p_control, dtype=np.float64); keep.append(lp)
        m.log_p_control = lp.ctypes.data
        m.omega_control = arr(omega_control, R)
        m.omega_case = arr(omega_case, R)
        m.kappa_control = arr(kappa_control, R)
        m.kappa_case = arr(kappa_case, R)
        m.merge_prob, m.split_prob = float(merge_prob), float(split_prob)
        m.hazard_mode = HAZARD_MODES[hazard]
        m.sort_preselect[0], m.sort_preselect[1] = int(sort_preselect[0]), int(sort_preselect[1])   # tuning / test hook, see the header
        m.sort_scratch_from = int(sort_scratch_from)                                                # test hook, see the header
        if rho_control is not None:
            rc = np.ascontiguousarray(rho_control, dtype=np.float64); rk = np.ascontiguousarray(rho_case, dtype=np.float64)
            keep += [rc, rk]
            m.rho_control, m.rho_case, m.d_max = rc.ctypes.data, rk.ctypes.data, rc.shape[1] - 1
        self._check(self.lib.hyg_tg_set_model(self.ctx, C.byref(m), int(t_max)), "hyg_tg_set_model")
        self.B = int(num_backward)

    def run(self, specs, want_taps=False):
        """specs: list of dicts(control_dataset, case_dataset, T, seed, chain_id).  Returns a list of dicts with
        trajectories (T x B x 5 int32: merged, d_control, r_control, d_case, r_case), log_normalizing_constant, taps (T x 4:
        particles proposed, K, finite-weight particles, sort attempts)."""
        n = len(specs)
        arr = (_lib.HygTgChain * n)()
        outs = []
        for i, s in enumerate(specs):
            T = int(s["T"])
            traj = s.get("trajectories")     # optional caller-owned (T, B, 5) int32 array, reused between runs
            if traj is None:
                traj = np.zeros((T, self.B, 5), dtype=np.int32)
            elif traj.shape != (T, self.B, 5) or traj.dtype != np.int32 or not traj.flags.c_contiguous:
                raise HygeiaError("trajectories must be a C-contiguous int32 array of shape (T, B, 5)")
            ln = np.zeros(1)
            taps = np.zeros((T, 4), dtype=np.int32) if want_taps else None
            arr[i].control_dataset, arr[i].case_dataset = s["control_dataset"], s["case_dataset"]
            arr[i].seed, arr[i].chain_id = int(s.get("seed", 0)), int(s.get("chain_id", i))
            arr[i].trajectories, arr[i].log_normalizing_constant, arr[i].taps = _ptr(traj), _ptr(ln), _ptr(taps)
            outs.append(dict(trajectories=traj, log_normalizing_constant=ln, taps=taps))
        ms = C.c_float(0)
        self._check(self.lib.hyg_tg_run(self.ctx, arr, n, C.byref(ms)), "hyg_tg_run")
        for o in outs:
            o["log_normalizing_constant"] = float(o["log_normalizing_constant"][0])
        self.ms_two_group = ms.value
        return outs


def hazard_table(omega, kappa, minimum_duration, d_max, hazard="exact"):
    """Host copy of the hazard table the kernels use: rho[r][d], d = 0..d_max (hyg_tg_hazard_table for "exact",
    hyg_tg_reference_hazard_table for "reference" -- see TwoGroupSession.set_two_group_model)."""
    lib = _lib.load()
    om = np.ascontiguousarray(omega, dtype=np.float64); ka = np.ascontiguousarray(kappa, dtype=np.float64)
    out = np.zeros((len(om), d_max + 1))
    fn = lib.hyg_tg_hazard_table if HAZARD_MODES[hazard] == 1 else lib.hyg_tg_reference_hazard_table
    rc = fn(_ptr(om), _ptr(ka), len(om), int(minimum_duration), int(d_max), _ptr(out))
    if rc < 0:
        raise HygeiaError(f"hazard table failed ({rc})")
    return out


def reference_hazard_table(omega, kappa, minimum_duration, d_max):
    return hazard_table(omega, kappa, minimum_duration, d_max, hazard="reference")


def summarise(trajectories, n_regimes):
    """The reference's test functions averaged over the backward trajectories (run_inference_two_groups.py:233-240,294-297):
    split_probs[t] = mean(merged == 0); regime_probs[t] = (control regime frequencies, case regime frequencies)."""
    tr = np.asarray(trajectories)
    split = (tr[:, :, 0] == 0).mean(axis=1).astype(np.float32)
    R = n_regimes
    reg = np.concatenate([(tr[:, :, 2, None] == np.arange(R)).mean(axis=1), (tr[:, :, 4, None] == np.arange(R)).mean(axis=1)], axis=1)
    return split, reg.astype(np.float32)


_sessions = {}


def infer(n_total_reads_control, n_methylated_reads_control, n_total_reads_case, n_methylated_reads_case, theta_control,
          mu=DEFAULT_MU, sigma=DEFAULT_SIGMA, minimum_duration=3, omega_case=0.8, merge_prob=0.1, split_prob=0.01,
          num_resampled_particles=50, num_samples_backward=25, seed=0, *, device=0, hazard="reference"):
    """One segment of one chromosome, as ``hygeia infer`` runs it (run_inference_two_groups.py:220-322).

    Count matrices are (n_sites, n_samples) as read from the reference's ``n_total_reads_{control,case}_<chrom>.txt.gz``
    files; ``theta_control`` is the single-group estimate (``theta_<chrom>.csv.gz``).  Returns the arrays the reference
    writes: backward particles (merged T x B, control / case T x B x 2 = (duration, regime), int16), split_probs, regime_probs
    and the log normalising constant keyed like the reference's dict by N = M (2R + R^2)."""
    if device not in _sessions:
        _sessions[device] = TwoGroupSession(device)
    s = _sessions[device]
    mu = [float(x) for x in mu]          # absl's DEFINE_list hands over strings
    sigma = [float(x) for x in sigma]
    R = len(mu)
    mats = []
    for a in (n_total_reads_control, n_methylated_reads_control, n_total_reads_case, n_methylated_reads_case):
        a = np.asarray(a)
        if a.ndim != 2:
            raise HygeiaError("count matrices must be (n_sites, n_samples)")
        if a.min(initial=0) < 0 or a.max(initial=0) > 65535:
            raise HygeiaError("read counts must be in 0..65535 (uint16 device layout)")
        mats.append(np.ascontiguousarray(a.T, dtype=np.uint16))   # [S][T], site fastest
    ntc, nmc, ntk, nmk = mats
    if (nmc > ntc).any() or (nmk > ntk).any():
        raise HygeiaError("methylated reads exceed total reads")   # the asserts at run_inference_two_groups.py:210-211
    T = ntc.shape[1]
    if ntk.shape[1] != T:
        raise HygeiaError("control and case matrices differ in the number of sites")
    logp, omega_control = control_group_parameters(theta_control, R)
    s.clear()
    s.set_emission_model(mu, sigma, minimum_duration)
    s.add_dataset(ntc, nmc)   # data set 0
    s.add_dataset(ntk, nmk)   # data set 1
    dc, dk = 0, 1
    s.emission()
    s.set_two_group_model(logp, omega_control, np.full(R, float(omega_case)), minimum_duration, num_resampled_particles,
                          num_samples_backward, merge_prob, split_prob, t_max=T, hazard=hazard)
    out = s.run([dict(control_dataset=dc, case_dataset=dk, T=T, seed=seed, chain_id=0)])[0]
    tr = out["trajectories"]
    split, reg = summarise(tr, R)
    N = num_resampled_particles * (2 * R + R * R)
    return dict(
        backward_particles_merged_state=tr[:, :, 0].astype(np.int16),
        backward_particles_control_state=tr[:, :, 1:3].astype(np.int16),
        backward_particles_case_state=tr[:, :, 3:5].astype(np.int16),
        split_probs=split, regime_probs=reg,
        log_normalizing_constants_optimal={N: out["log_normalizing_constant"]},
        device_ms=s.ms_two_group,
    )
