"""ctypes bindings for the CPU checkers under oracle/ (TEST INFRASTRUCTURE).

* ``Oracle``  -> oracle/libhyg_oracle.so  : our restatement (oracle/sg_oracle.cpp)
* ``Ref``     -> oracle/_ref/libhyg_ref*.so: the reference's own headers compiled in place
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _OracleArgs(C.Structure):
    _fields_ = [
        ("vartheta", C.c_void_p), ("n_vartheta", C.c_uint32), ("theta", C.c_void_p), ("dim_theta", C.c_uint32),
        ("T", C.c_uint64), ("S", C.c_uint32), ("positions", C.c_void_p), ("n_total", C.c_void_p), ("n_meth", C.c_void_p),
        ("logobs", C.c_void_p),
        ("n_particles_max", C.c_uint32), ("use_smoothing", C.c_int32), ("epsilon", C.c_double), ("use_param_est", C.c_int32),
        ("normalise_gradients", C.c_int32), ("use_adam", C.c_int32), ("n_steps_without_update", C.c_uint32),
        ("lr_exponent", C.c_double), ("lr_factor", C.c_double),
        ("uniforms_by_site", C.c_void_p),
        ("regime_probs", C.c_void_p), ("theta_trace", C.c_void_p), ("logz", C.c_void_p), ("n_curr", C.c_void_p),
        ("k_kept", C.c_void_p), ("finalised_at", C.c_void_p), ("drew_uniform", C.c_void_p), ("n_pending", C.c_void_p),
        ("ancestors", C.c_void_p), ("seconds", C.c_void_p), ("tie_pairs", C.c_void_p), ("weights_prev", C.c_void_p), ("d_prev", C.c_void_p),
        ("tie_order", C.c_int32), ("support_hash", C.c_void_p), ("tie_flags", C.c_void_p),
    ]


class _RefArgs(C.Structure):
    _fields_ = [
        ("vartheta", C.c_void_p), ("n_vartheta", C.c_uint32), ("theta", C.c_void_p), ("dim_theta", C.c_uint32),
        ("T", C.c_uint64), ("S", C.c_uint32), ("positions", C.c_void_p), ("n_total", C.c_void_p), ("n_meth", C.c_void_p),
        ("n_particles_max", C.c_uint32), ("smc_proposal_type", C.c_uint32), ("smc_resample_type", C.c_uint32),
        ("use_smoothing", C.c_int32), ("epsilon", C.c_double), ("use_param_est", C.c_int32),
        ("normalise_gradients", C.c_int32), ("use_adam", C.c_int32), ("n_steps_without_update", C.c_uint32),
        ("lr_exponent", C.c_double), ("lr_factor", C.c_double),
        ("uniforms_by_site", C.c_void_p), ("rng_seed", C.c_uint64), ("stepwise", C.c_int32),
        ("regime_probs", C.c_void_p), ("theta_trace", C.c_void_p), ("logz", C.c_void_p), ("n_curr", C.c_void_p),
        ("finalised_at", C.c_void_p), ("drew_uniform", C.c_void_p), ("n_pending", C.c_void_p), ("seconds", C.c_void_p),
        ("dump_at", C.c_int64), ("dump_logw", C.c_void_p), ("dump_W", C.c_void_p), ("dump_d", C.c_void_p), ("dump_r", C.c_void_p),
    ]


def sample_fastest(counts_st):
    """[S][T] (site fastest, the device layout) -> uint32 S x T with the sample index fastest (arma::umat column-major)."""
    return np.ascontiguousarray(np.asarray(counts_st).T, dtype=np.uint32)


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(ORACLE_DIR, "libhyg_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing -- run `make -C oracle oracle` (or __graft_entry__.build())")
        self.lib = C.CDLL(path)
        self.lib.hygo_log_beta_binomial.restype = C.c_double
        self.lib.hygo_log_beta_binomial.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        self.lib.hygo_sg_run.argtypes = [C.POINTER(_OracleArgs)]

    def log_beta_binomial(self, x, n, a, b):
        return self.lib.hygo_log_beta_binomial(x, n, a, b)

    def emission(self, alpha, beta, n_total_st, n_meth_st):
        nt, nm = sample_fastest(n_total_st), sample_fastest(n_meth_st)
        T, S = nt.shape
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        beta = np.ascontiguousarray(beta, dtype=np.float64)
        out = np.empty((T, len(alpha)), dtype=np.float64)
        self.lib.hygo_sg_emission(_p(alpha), _p(beta), C.c_uint32(len(alpha)), C.c_uint64(T), C.c_uint32(S), _p(nt), _p(nm), _p(out))
        return out

    def tables(self, vartheta, theta, d_max):
        vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        R = int(vartheta[1])
        rho = np.zeros((R, d_max)); ex = np.zeros((R, d_max), np.uint8); g = np.zeros((R, d_max))
        P = np.zeros((R, R)); om = np.zeros(R)
        rc = self.lib.hygo_sg_tables(_p(vartheta), C.c_uint32(len(vartheta)), _p(theta), C.c_uint32(len(theta)), C.c_uint32(d_max),
                                     _p(rho), _p(ex), _p(g), _p(P), _p(om))
        assert rc == 0, rc
        return dict(rho=rho, exit=ex, grad=g, P=P, omega=om)

    def run(self, vartheta, theta, uniforms, n_total_st=None, n_meth_st=None, positions=None, logobs=None,
            n_particles=250, smoothing=True, epsilon=0.01, param_est=False, normalise=False, adam=True,
            n_steps_without_update=200, lr_exponent=0.1, lr_factor=0.01, want_ancestors=False, want_weights=False,
            tie_order="reference"):
        """tie_order: "reference" = the reference's own (arrangement-dependent) order of exactly equal weights, identical to
        oracle/_ref; "canonical" = log-weight, then (regime, sojourn): the storage-order-independent rule of the CUDA path."""
        vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        R = int(vartheta[1]); D = len(theta)
        if logobs is not None:
            logobs = np.ascontiguousarray(logobs, dtype=np.float64)
            T, S = logobs.shape[0], 0
            nt = nm = None
        else:
            nt, nm = sample_fastest(n_total_st), sample_fastest(n_meth_st)
            T, S = nt.shape
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert uniforms.shape[0] >= T
        pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.uint32)
        out = dict(
            regime_probs=np.full((T, 1 + R), np.nan) if smoothing else None,
            theta_trace=np.zeros((T, D)) if param_est else None,
            logz=np.zeros(T), n_curr=np.zeros(T, np.int32), k_kept=np.zeros(T, np.int32),
            finalised_at=np.full(T, -1, np.int32), drew_uniform=np.zeros(T, np.uint8), n_pending=np.zeros(T, np.int32),
            ancestors=np.full((T, n_particles - R), -1, np.int16) if want_ancestors else None,
            tie_pairs=np.zeros(T, np.int32),
            weights_prev=np.full((T, n_particles), np.nan) if want_weights else None,
            d_prev=np.zeros((T, n_particles), np.int32) if want_weights else None,
            support_hash=np.zeros(T, np.uint64), tie_flags=np.zeros(T, np.uint8),
        )
        sec = C.c_double(0.0)
        a = _OracleArgs(_p(vartheta), len(vartheta), _p(theta), D, T, S, _p(pos), _p(nt), _p(nm), _p(logobs),
                        n_particles, int(smoothing), epsilon, int(param_est), int(normalise), int(adam), n_steps_without_update,
                        lr_exponent, lr_factor, _p(uniforms),
                        _p(out["regime_probs"]), _p(out["theta_trace"]), _p(out["logz"]), _p(out["n_curr"]), _p(out["k_kept"]),
                        _p(out["finalised_at"]), _p(out["drew_uniform"]), _p(out["n_pending"]), _p(out["ancestors"]),
                        C.cast(C.pointer(sec), C.c_void_p), _p(out["tie_pairs"]), _p(out["weights_prev"]), _p(out["d_prev"]),
                        {"reference": 0, "canonical": 1}[tie_order], _p(out["support_hash"]), _p(out["tie_flags"]))
        rc = self.lib.hygo_sg_run(C.byref(a))
        assert rc == 0, f"hygo_sg_run failed: {rc}"
        out["seconds"] = sec.value
        return out


class Ref:
    """The reference's own C++ (oracle/_ref).  variant: '' (reference flags, -ffast-math) or '_strict' (-O2)."""

    def __init__(self, variant=""):
        path = os.path.join(ORACLE_DIR, "_ref", f"libhyg_ref{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.hygref_log_beta_binomial.restype = C.c_double
        self.lib.hygref_log_beta_binomial.argtypes = [C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        self.lib.hygref_sg_run.argtypes = [C.POINTER(_RefArgs)]

    @staticmethod
    def available(variant=""):
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", f"libhyg_ref{variant}.so"))

    def log_beta_binomial(self, x, n, a, b):
        return self.lib.hygref_log_beta_binomial(x, n, a, b)

    def emission(self, vartheta, n_total_st, n_meth_st):
        vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
        nt, nm = sample_fastest(n_total_st), sample_fastest(n_meth_st)
        T, S = nt.shape
        R = int(vartheta[1])
        out = np.empty((T, R), dtype=np.float64)
        self.lib.hygref_sg_emission(_p(vartheta), C.c_uint32(len(vartheta)), C.c_uint64(T), C.c_uint32(S), _p(nt), _p(nm), _p(out))
        return out

    def tables(self, vartheta, theta, d_max):
        vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        R = int(vartheta[1])
        rho = np.zeros((R, d_max)); ex = np.zeros((R, d_max), np.uint8); g = np.zeros((R, d_max))
        P = np.zeros((R, R)); om = np.zeros(R)
        self.lib.hygref_sg_tables(_p(vartheta), C.c_uint32(len(vartheta)), _p(theta), C.c_uint32(len(theta)), C.c_uint32(d_max),
                                  _p(rho), _p(ex), _p(g), _p(P), _p(om))
        return dict(rho=rho, exit=ex, grad=g, P=P, omega=om)

    def run(self, vartheta, theta, n_total_st, n_meth_st, positions=None, uniforms=None, rng_seed=0, stepwise=True,
            n_particles=250, smoothing=True, epsilon=0.01, param_est=False, normalise=False, adam=True,
            n_steps_without_update=200, lr_exponent=0.1, lr_factor=0.01, dump_at=-1):
        vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        R = int(vartheta[1]); D = len(theta)
        nt, nm = sample_fastest(n_total_st), sample_fastest(n_meth_st)
        T, S = nt.shape
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.uint32)
        out = dict(
            regime_probs=np.full((T, 1 + R), np.nan) if smoothing else None,
            theta_trace=np.zeros((T, D)) if param_est else None,
            logz=np.zeros(T), n_curr=np.zeros(T, np.int32), finalised_at=np.full(T, -1, np.int32),
            drew_uniform=np.zeros(T, np.uint8), n_pending=np.zeros(T, np.int32),
            dump_logw=np.full(n_particles, np.nan), dump_W=np.full(n_particles, np.nan),
            dump_d=np.zeros(n_particles, np.uint32), dump_r=np.zeros(n_particles, np.uint32),
        )
        sec = C.c_double(0.0)
        a = _RefArgs(_p(vartheta), len(vartheta), _p(theta), D, T, S, _p(pos), _p(nt), _p(nm),
                     n_particles, 1, 2, int(smoothing), epsilon, int(param_est), int(normalise), int(adam), n_steps_without_update,
                     lr_exponent, lr_factor, _p(uniforms), rng_seed, int(stepwise),
                     _p(out["regime_probs"]), _p(out["theta_trace"]), _p(out["logz"]), _p(out["n_curr"]),
                     _p(out["finalised_at"]), _p(out["drew_uniform"]), _p(out["n_pending"]),
                     C.cast(C.pointer(sec), C.c_void_p),
                     dump_at, _p(out["dump_logw"]), _p(out["dump_W"]), _p(out["dump_d"]), _p(out["dump_r"]))
        rc = self.lib.hygref_sg_run(C.byref(a))
        assert rc == 0
        out["seconds"] = sec.value
        return out
