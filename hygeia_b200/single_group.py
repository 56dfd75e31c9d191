"""Host-side mirror of the reference's single-group operator, on top of the C ABI.

``run_online_combined_inference`` takes the same arguments, in the same order and with the same meaning, as the Rcpp
export ``runOnlineCombinedInferenceCpp`` (/root/reference/src/single_group/src/cpp/singleGroup.cpp:76-96; R call site
src/single_group/bin/estimate_parameters_and_regimes:303-322) and returns the same three things
(``regimeProbabilityEstimates``, ``thetaEstimates``, ``cpuTime``).  ``Session`` is the batched, device-resident form used
for whole-genome sweeps (many chains = chromosomes x seeds in one launch).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class HygeiaError(RuntimeError):
    pass


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def make_run_args(n_particles_max=250, smc_proposal_type=1, smc_resample_type=2, use_online_marginal_smoothing=True,
                  epsilon=0.01, use_online_parameter_estimation=False, normalise_gradients=False, use_adam=True,
                  n_steps_without_parameter_update=200, learning_rate_exponent=0.1, learning_rate_factor=0.01,
                  lag_capacity=1024, allow_forced_emission=False, resample_full_sort=False):
    a = _lib.HygRunArgs()
    a.n_particles_max = n_particles_max
    a.smc_proposal_type = smc_proposal_type
    a.smc_resample_type = smc_resample_type
    a.use_online_marginal_smoothing = int(bool(use_online_marginal_smoothing))
    a.epsilon = epsilon
    a.use_online_parameter_estimation = int(bool(use_online_parameter_estimation))
    a.normalise_gradients = int(bool(normalise_gradients))
    a.use_adam = int(bool(use_adam))
    a.n_steps_without_parameter_update = n_steps_without_parameter_update
    a.learning_rate_exponent = learning_rate_exponent
    a.learning_rate_factor = learning_rate_factor
    a.lag_capacity = lag_capacity
    a.allow_forced_emission = int(bool(allow_forced_emission))
    a.resample_full_sort = int(bool(resample_full_sort))
    return a


STATUS_WORDS = ("forced_emissions", "max_pending", "halo_forced", "steps", "full_sorts", "tie_decided_sites", "double_draws",
                "overlap_rows_over_tolerance")


class Session:
    """One device context: model -> data sets -> chains -> emission (K1) -> recursion (K2) -> download."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.ctx = self.lib.hyg_create(device)
        if not self.ctx:
            raise HygeiaError("hyg_create failed: " + self.lib.hyg_create_error().decode())
        self._keep = []       # host arrays that the C side holds pointers to
        self._chains = None
        self.R = None

    def close(self):
        if self.ctx:
            self.lib.hyg_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise HygeiaError(f"{what} failed ({rc}): {self.lib.hyg_last_error(self.ctx).decode()}")
        return rc

    @property
    def stream(self):
        return self.lib.hyg_stream(self.ctx)

    def set_vartheta(self, vartheta):
        v = np.ascontiguousarray(vartheta, dtype=np.float64)
        self.R = int(v[1])
        self._check(self.lib.hyg_sg_set_vartheta(self.ctx, _ptr(v), len(v)), "hyg_sg_set_vartheta")

    def set_theta(self, theta, t_max):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self.lib.hyg_sg_set_theta(self.ctx, _ptr(th), len(th), int(t_max)), "hyg_sg_set_theta")

    def get_tables(self, d_max):
        R = self.R
        P = np.zeros((R, R)); om = np.zeros(R); rho = np.zeros((R, d_max)); ex = np.zeros((R, d_max), np.uint8)
        self._check(self.lib.hyg_sg_get_tables(self.ctx, _ptr(P), _ptr(om), d_max, _ptr(rho), _ptr(ex)), "hyg_sg_get_tables")
        return dict(P=P, omega=om, rho=rho, exit=ex)

    def clear(self):
        self._check(self.lib.hyg_sg_clear(self.ctx), "hyg_sg_clear")
        self._keep = []
        self._chains = None

    def add_dataset(self, n_total, n_meth):
        """Host uint16 arrays [S][T] (site fastest); copied host -> device."""
        nt = np.ascontiguousarray(n_total, dtype=np.uint16)
        nm = np.ascontiguousarray(n_meth, dtype=np.uint16)
        S, T = nt.shape
        return self._check(self.lib.hyg_sg_add_dataset(self.ctx, T, S, _ptr(nt), _ptr(nm), 0, T), "hyg_sg_add_dataset")

    def add_dataset_ptr(self, T, S, n_total_ptr, n_meth_ptr, on_device, pitch):
        """Raw pointers (pinned host memory or device memory, e.g. torch tensors' data_ptr())."""
        return self._check(self.lib.hyg_sg_add_dataset(self.ctx, T, S, n_total_ptr, n_meth_ptr, int(on_device), pitch), "hyg_sg_add_dataset")

    def set_chains(self, specs):
        """specs: list of dicts with keys dataset, seed, chain_id and optional host arrays / raw pointers:
        uniforms, positions, regime_probs, logz, k_kept, drew_uniform, n_pending, n_curr, finalised_at, support_hash, tie_flags."""
        n = len(specs)
        arr = (_lib.HygChain * n)()
        for i, s in enumerate(specs):
            arr[i].dataset = s["dataset"]
            arr[i].seed = s.get("seed", 0)
            arr[i].chain_id = s.get("chain_id", i)
            for k in ("uniforms", "positions", "regime_probs", "logz", "theta_trace", "k_kept", "drew_uniform", "n_pending",
                      "n_curr", "finalised_at", "support_hash", "tie_flags"):
                v = s.get(k)
                setattr(arr[i], k, _ptr(v))
                if v is not None and not isinstance(v, int):
                    self._keep.append(v)
        self._chains = arr
        self._check(self.lib.hyg_sg_set_chains(self.ctx, arr, n), "hyg_sg_set_chains")

    SEGMENT_AUTO = 0xFFFFFFFFFFFFFFFF

    def set_segmentation(self, segment_sites=0, halo_left=5000, halo_right=5000):
        """Throughput mode: cut every chain into concurrent segments of <= segment_sites sites (0 = whole chains, the
        reference's sequential run).  See hyg_sg_set_segmentation in include/hygeia_b200.h."""
        self._check(self.lib.hyg_sg_set_segmentation(self.ctx, int(segment_sites), int(halo_left), int(halo_right)), "hyg_sg_set_segmentation")

    def set_zero_copy_outputs(self, enable=True):
        """Pinned ``regime_probs`` buffers are written by the kernel directly (no D2H stage); False forces staging."""
        self._check(self.lib.hyg_sg_set_zero_copy_outputs(self.ctx, int(bool(enable))), "hyg_sg_set_zero_copy_outputs")

    def filter_units(self, with_segment_sites=False):
        """Units of the last filter launch; with_segment_sites: (units, segment size used, persistent CTAs)."""
        n, seg, g = C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
        self._check(self.lib.hyg_sg_filter_units(self.ctx, C.byref(n), C.byref(seg), C.byref(g)), "hyg_sg_filter_units")
        return (n.value, seg.value, g.value) if with_segment_sites else n.value

    def emission(self):
        self._check(self.lib.hyg_sg_emission(self.ctx), "hyg_sg_emission")

    def filter(self, run_args=None):
        a = run_args or make_run_args()
        self._check(self.lib.hyg_sg_filter(self.ctx, C.byref(a)), "hyg_sg_filter")

    def download(self):
        """Copies the outputs back and returns the per-chain status words (include/hygeia_b200.h); raises HygeiaError with
        HYG_ERR_CAPACITY when a lag set overflowed (unless allow_forced_emission was set)."""
        self._check(self.lib.hyg_sg_download(self.ctx, self._chains, len(self._chains)), "hyg_sg_download")
        return [tuple(c.status) for c in self._chains]

    def overlap_max_abs(self):
        """Segmented execution: per chain, the largest difference between the rows a segment recomputed in its right halo and
        the rows the next segment wrote (the run-time check of the LEFT halo); 0.0 for whole chains."""
        return [float(c.overlap_max_abs) for c in self._chains]

    def device_outputs(self, chain):
        """(regime_probs, logz) device addresses of a staged chain (0 where absent): for consumers that stay on the device."""
        a, b = C.c_void_p(0), C.c_void_p(0)
        self._check(self.lib.hyg_sg_device_outputs(self.ctx, int(chain), C.byref(a), C.byref(b)), "hyg_sg_device_outputs")
        return (a.value or 0), (b.value or 0)

    def sync(self):
        self._check(self.lib.hyg_sync(self.ctx), "hyg_sync")

    def timings(self):
        a, b = C.c_float(0), C.c_float(0)
        la, lb = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.hyg_sg_timings(self.ctx, C.byref(a), C.byref(b), C.byref(la), C.byref(lb)), "hyg_sg_timings")
        return dict(ms_emission=a.value, ms_filter=b.value, emission_launches=la.value, filter_launches=lb.value)

    def get_logobs(self, dataset, T):
        out = np.empty((T, self.R), dtype=np.float64)
        self._check(self.lib.hyg_sg_get_logobs(self.ctx, dataset, _ptr(out)), "hyg_sg_get_logobs")
        return out


_default_session = {}


def _session(device=0):
    if device not in _default_session:
        _default_session[device] = Session(device)
    return _default_session[device]


def run_online_combined_inference(vartheta, theta_init, genomic_positions, n_total_reads, n_methylated_reads,
                                  n_particles_max=250, smc_proposal_type=1, smc_resample_type=2,
                                  use_online_marginal_smoothing=True, epsilon=0.01,
                                  use_online_parameter_estimation=False, normalise_gradients=False, use_adam=True,
                                  n_steps_without_parameter_update=200, learning_rate_exponent=0.1, learning_rate_factor=0.01,
                                  randomise_rng_seed=False, rng_seed=0, *, uniforms_by_site=None, device=0, return_logz=False,
                                  lag_capacity=1024):
    """Drop-in for ``runOnlineCombinedInferenceCpp`` (singleGroup.cpp:76-96).

    ``n_total_reads`` / ``n_methylated_reads`` are (n_samples, n_cpg_sites) matrices as in the R call; they are narrowed to
    uint16 with the site index fastest, which is the device layout.  Returns a dict with ``regimeProbabilityEstimates``
    (T x (1+R): genomic position, regime_1..R), ``thetaEstimates`` (T x D in parameter mode, else None) and ``cpuTime``.
    ``uniforms_by_site`` injects the resampling draws (one per site); otherwise they come from Philox keyed by ``rng_seed``.
    ``status`` in the result names the kernel's status words (forced emissions, sites where an exact tie of weights decided a
    particle's fate, ...: see include/hygeia_b200.h); a lag-set overflow raises HygeiaError instead of returning a different
    estimator silently.
    """
    s = _session(device)
    vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
    theta = np.ascontiguousarray(theta_init, dtype=np.float64)
    nt = np.asarray(n_total_reads)
    nm = np.asarray(n_methylated_reads)
    if nt.max(initial=0) > 65535 or nm.max(initial=0) > 65535:
        raise HygeiaError("read counts above 65535 are not representable in the uint16 device layout")
    nt = np.ascontiguousarray(nt, dtype=np.uint16)
    nm = np.ascontiguousarray(nm, dtype=np.uint16)
    S, T = nt.shape
    R = int(vartheta[1])
    pos = None if genomic_positions is None else np.ascontiguousarray(genomic_positions, dtype=np.uint32)
    if randomise_rng_seed:
        rng_seed = int(np.random.SeedSequence().entropy & 0x7FFFFFFFFFFFFFFF)
    args = make_run_args(n_particles_max, smc_proposal_type, smc_resample_type, use_online_marginal_smoothing, epsilon,
                         use_online_parameter_estimation, normalise_gradients, use_adam, n_steps_without_parameter_update,
                         learning_rate_exponent, learning_rate_factor, lag_capacity=lag_capacity)
    probs = np.full((T, 1 + R), np.nan) if use_online_marginal_smoothing else None
    trace = np.zeros((T, len(theta))) if use_online_parameter_estimation else None
    logz = np.zeros(T) if return_logz else None
    un = None if uniforms_by_site is None else np.ascontiguousarray(uniforms_by_site, dtype=np.float64)
    sec = C.c_double(0.0)
    status = np.zeros(8, np.int32)
    rc = s.lib.hyg_sg_run_online_combined_inference(s.ctx, _ptr(vartheta), len(vartheta), _ptr(theta), len(theta), T, S, _ptr(pos),
                                                    _ptr(nt), _ptr(nm), C.byref(args), int(rng_seed) & 0xFFFFFFFFFFFFFFFF, _ptr(un),
                                                    _ptr(probs), _ptr(trace), _ptr(logz), C.byref(sec), _ptr(status))
    s._check(rc, "hyg_sg_run_online_combined_inference")
    out = dict(regimeProbabilityEstimates=probs, thetaEstimates=trace, cpuTime=sec.value,
               status=dict(zip(STATUS_WORDS, (int(v) for v in status))))
    if return_logz:
        out["logZ"] = logz
    return out
