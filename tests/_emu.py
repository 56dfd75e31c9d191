"""ctypes bindings for tests/emu/libhyg_emu.so: the repo's DEVICE code compiled for the CPU under an emulation of the
CUDA execution model (TEST INFRASTRUCTURE -- lets the GPU-less box check kernel logic against the oracle)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
LIB = os.path.join(EMU_DIR, "libhyg_emu.so")


def build():
    src = [os.path.join(EMU_DIR, "emu_kernels.cpp"), os.path.join(HERE, "..", "hygeia_b200", "csrc", "hyg_tables.cpp")]
    deps = src + [os.path.join(EMU_DIR, "cuda_emu.h")] + [
        os.path.join(HERE, "..", "hygeia_b200", "csrc", f) for f in
        ("sg_filter.cuh", "sg_param.cuh", "hyg_tg.cuh", "sg_emission.cuh", "hyg_common.cuh", "hyg_dev_structs.h", "hyg_tables.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-I", EMU_DIR] + src + ["-o", LIB])
    return LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emu:
    def __init__(self):
        self.lib = C.CDLL(build())

    def sg_filter(self, vartheta, theta, logobs, uniforms=None, seed=0, chain_id=0, n_particles=250, smoothing=True,
                  epsilon=0.01, lcap=64, force_full_sort=False, param_est=False, normalise=False, adam=True,
                  n_steps_without_update=200, lr_exponent=0.1, lr_factor=0.01, t_off=0, own=None, last_segment=True):
        """`logobs` (and `uniforms`) are the LOCAL slice of a segment; own = (own_lo, own_hi) local indices (default: all)."""
        vartheta = np.ascontiguousarray(vartheta, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        logobs = np.ascontiguousarray(logobs, dtype=np.float64)
        T, R = logobs.shape
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = dict(probs=np.full((T, R + 1), np.nan), logz=np.zeros(T), k_kept=np.zeros(T, np.int32), drew_uniform=np.zeros(T, np.uint8),
                   n_pending=np.zeros(T, np.int32), n_curr=np.zeros(T, np.int32), finalised_at=np.full(T, -1, np.int32),
                   support_hash=np.zeros(T, np.uint64), tie_flags=np.zeros(T, np.uint8), status=np.zeros(8, np.int32),
                   seg_inc=np.zeros(1),
                   theta_trace=np.zeros((T, len(theta))) if param_est else None)
        rc = self.lib.hygemu_sg_filter(_p(vartheta), C.c_uint32(len(vartheta)), _p(theta), C.c_uint32(len(theta)), C.c_uint32(n_particles),
                                       C.c_uint64(T), _p(logobs), _p(uniforms), C.c_uint64(seed), C.c_uint32(chain_id),
                                       C.c_int(int(smoothing)), C.c_double(epsilon), C.c_int(lcap),
                                       C.c_int(int(param_est)), C.c_int(int(normalise)), C.c_int(int(adam)), C.c_uint32(n_steps_without_update),
                                       C.c_double(lr_exponent), C.c_double(lr_factor), _p(out["theta_trace"]),
                                       _p(out["probs"]), _p(out["logz"]), _p(out["k_kept"]), _p(out["drew_uniform"]), _p(out["n_pending"]),
                                       _p(out["n_curr"]), _p(out["finalised_at"]), _p(out["support_hash"]), _p(out["tie_flags"]), _p(out["status"]),
                                       C.c_uint64(t_off), C.c_uint64(own[0] if own else 0), C.c_uint64(own[1] if own else T),
                                       C.c_int(int(last_segment)), _p(out["seg_inc"]), C.c_int(int(force_full_sort)))
        assert rc == 0, rc
        out["probs"] = out["probs"][:, 1:]
        return out

    def sg_emission(self, alpha, beta, n_total_st, n_meth_st, nmax_table=255, nmax_smem=96, grid=3, block=64):
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        beta = np.ascontiguousarray(beta, dtype=np.float64)
        S, T = n_total_st.shape
        pitch = (T + 7) // 8 * 8
        nt = np.zeros((S, pitch), np.uint16); nt[:, :T] = n_total_st
        nm = np.zeros((S, pitch), np.uint16); nm[:, :T] = n_meth_st
        out = np.full((T, len(alpha)), np.nan)
        rc = self.lib.hygemu_sg_emission(_p(alpha), _p(beta), C.c_int(len(alpha)), C.c_uint64(T), C.c_uint32(S), C.c_uint64(pitch),
                                         _p(nt), _p(nm), C.c_int(nmax_table), C.c_int(nmax_smem), C.c_int(grid), C.c_int(block), _p(out))
        assert rc == 0, rc
        return out


def tg_run_emu(emu, model, lo_c, lo_k, M=50, B=25, seed=0, chain=0, preselect=(0, 0), scratch_from=0):
    """Two-group filter + backward simulation under emulation; `model` is an oracle/tg_oracle.TwoGroupModel."""
    lo_c = np.ascontiguousarray(lo_c, dtype=np.float64); lo_k = np.ascontiguousarray(lo_k, dtype=np.float64)
    T, R = lo_c.shape
    logP = np.ascontiguousarray(model.logP); logPm = np.ascontiguousarray(model.logPm)
    rho_c = np.ascontiguousarray(model.rho_c); rho_k = np.ascontiguousarray(model.rho_k)
    traj = np.zeros((T, B, 5), np.int32); ln = np.zeros(1); taps = np.zeros((T, 4), np.int32)
    rc = emu.lib.hygemu_tg_run(C.c_int(R), C.c_int(model.u), C.c_int(M), C.c_int(B), _p(logP), _p(logPm), _p(rho_c), _p(rho_k),
                               C.c_uint32(model.d_max), C.c_uint64(T), _p(lo_c), _p(lo_k), C.c_uint64(seed), C.c_uint32(chain),
                               _p(traj), _p(ln), _p(taps), C.c_int(preselect[0]), C.c_int(preselect[1]), C.c_int(scratch_from))
    assert rc == 0
    return dict(traj=traj, log_norm=float(ln[0]), taps=taps)
