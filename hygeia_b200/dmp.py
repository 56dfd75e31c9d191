"""Host-side mirror of the reference's DMP-calling reductions, on top of the C ABI (K6 + device sort/scan).

Same names, argument order and return values as /root/reference/src/two_group/multiple_testing.py
(``FDR_procedure`` :3-12, ``weighted_FDR_procedure`` :13-22); ``site_statistics`` computes what
``aggregate_results.py:125-147,181`` and ``get_dmps.py:63-76,119-126`` compute from the aggregated trajectory matrices.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .single_group import HygeiaError, _session


def _ctx(device):
    return _session(device)


def site_statistics(merge_states, control_regimes, case_regimes, n_regimes=6, test_regime_combinations=False, device=0):
    """merge_states / control_regimes / case_regimes: integer arrays [T][P] (site x particle, all seeds concatenated).

    Returns dict(split_probs[T], null_stats[T], control_freqs[T][R], case_freqs[T][R], pair_stats[T][R][R] or None, ms_device).
    """
    s = _ctx(device)
    m = np.ascontiguousarray(merge_states, dtype=np.int8)
    c = np.ascontiguousarray(control_regimes, dtype=np.int8)
    k = np.ascontiguousarray(case_regimes, dtype=np.int8)
    if m.ndim != 2 or m.shape != c.shape or m.shape != k.shape:
        raise HygeiaError("merge_states, control_regimes and case_regimes must be [T][P] arrays of one shape")
    T, P = m.shape
    R = int(n_regimes)
    out = dict(split_probs=np.empty(T), null_stats=np.empty(T), control_freqs=np.empty((T, R)), case_freqs=np.empty((T, R)),
               pair_stats=np.empty((T, R, R)) if test_regime_combinations else None)
    ms = C.c_float(0.0)
    rc = s.lib.hyg_tg_site_statistics(s.ctx, T, P, R, m.ctypes.data, c.ctypes.data, k.ctypes.data, 0, out["split_probs"].ctypes.data,
                                      out["null_stats"].ctypes.data, out["control_freqs"].ctypes.data, out["case_freqs"].ctypes.data,
                                      None if out["pair_stats"] is None else out["pair_stats"].ctypes.data, C.byref(ms))
    s._check(rc, "hyg_tg_site_statistics")
    out["ms_device"] = ms.value
    return out


def FDR_procedure(test_statistics, fdr_threshold, device=0):  # noqa: N802 -- the reference's name
    """(k, Qk, threshold) as multiple_testing.FDR_procedure."""
    s = _ctx(device)
    t = np.ascontiguousarray(test_statistics, dtype=np.float64).ravel()
    k = C.c_uint64(0); q = C.c_double(0.0); th = C.c_double(0.0)
    rc = s.lib.hyg_fdr_procedure(s.ctx, t.shape[0], t.ctypes.data, float(fdr_threshold), C.byref(k), C.byref(q), C.byref(th))
    s._check(rc, "hyg_fdr_procedure")
    return int(k.value), float(q.value), float(th.value)


def weighted_FDR_procedure(test_statistics, fdr_threshold, weights_false_positives, weights_false_negatives, device=0):  # noqa: N802
    """(ranking_indices[:s], Nsums[s-1]) as multiple_testing.weighted_FDR_procedure."""
    s = _ctx(device)
    t = np.ascontiguousarray(test_statistics, dtype=np.float64).ravel()
    wfp = np.ascontiguousarray(weights_false_positives, dtype=np.float64).ravel()
    wfn = np.ascontiguousarray(weights_false_negatives, dtype=np.float64).ravel()
    if wfp.shape != t.shape or wfn.shape != t.shape:
        raise HygeiaError("weights must have one entry per test statistic")
    idx = np.empty(t.shape[0], dtype=np.uint64)
    n = C.c_uint64(0); nk = C.c_double(0.0)
    rc = s.lib.hyg_weighted_fdr_procedure(s.ctx, t.shape[0], t.ctypes.data, float(fdr_threshold), wfp.ctypes.data, wfn.ctypes.data,
                                          C.byref(n), idx.ctypes.data, C.byref(nk))
    s._check(rc, "hyg_weighted_fdr_procedure")
    return idx[:n.value].astype(np.int64), float(nk.value)
