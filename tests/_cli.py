"""Helpers for the tests of the `hygeia` command-line front end: run the binary, write inputs in the reference's formats."""
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "hygeia")


def run(*args, check=True, timeout=600):
    r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    if check and r.returncode != 0:
        raise AssertionError(f"hygeia {' '.join(map(str, args))} failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
    return r


def write_preprocess_style(path, a):
    """np.savetxt(fmt='%s', delimiter=',') of a float array, no header (preprocess_bed.py:464-467): counts look like 12.0"""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    with gzip.open(path, "wt") as f:
        for row in a:
            f.write(",".join(repr(float(v)) for v in row) + "\n")


def write_simulate_style(path, a, names):
    """readr::write_csv with a header line (input_output_functions.R:34-48)"""
    a = np.asarray(a)
    if a.ndim == 1:
        a = a[:, None]
    with gzip.open(path, "wt") as f:
        f.write(",".join(names) + "\n")
        for row in a:
            f.write(",".join(str(int(v)) for v in row) + "\n")


def read_csv(path):
    """(header, rows of raw strings) of a comma-separated file, gz or not"""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines[0].split(","), [ln.split(",") for ln in lines[1:]]


def make_infer_tree(root, chrom="21", n_batches=2, n_seeds=3, sites_per_batch=240, n_samples=1, R=6, B=25, M=50, seed=0):
    """A results tree as `hygeia infer` leaves it (run_inference_two_groups.py:246-255,304-322): per batch directory
    chrom_{chrom}_{batch}/ the comma-separated np.savetxt echoes of the window's counts and positions and, per seed, the int16
    backward trajectories (np.savez_compressed).  Trajectories are piecewise constant, with stretches where case and control
    differ, so that the DMP procedures have something to find.  Returns what was written, for checks."""
    rng = np.random.default_rng(seed)
    N = M * (2 * R + R * R)
    pos0 = 10_000
    out = []
    for b in range(n_batches):
        d = os.path.join(root, f"chrom_{chrom}_{b}")
        os.makedirs(d, exist_ok=True)
        T = sites_per_batch
        pos = pos0 + np.cumsum(rng.integers(2, 900, size=T))
        pos0 = int(pos[-1])
        nt_c = rng.poisson(20, size=(T, n_samples)); nm_c = rng.binomial(nt_c, 0.7)
        nt_k = rng.poisson(20, size=(T, n_samples)); nm_k = rng.binomial(nt_k, 0.4)
        np.savetxt(os.path.join(d, "observations_control.csv.gz"), nm_c.astype(np.int16), delimiter=",")
        np.savetxt(os.path.join(d, "observations_case.csv.gz"), nm_k.astype(np.int16), delimiter=",")
        np.savetxt(os.path.join(d, "n_total_reads_control.csv.gz"), nt_c.astype(np.int16), delimiter=",")
        np.savetxt(os.path.join(d, "n_total_reads_case.csv.gz"), nt_k.astype(np.int16), delimiter=",")
        np.savetxt(os.path.join(d, "positions.csv.gz"), pos, delimiter=",")
        # a common "truth" per batch, each trajectory a noisy copy of it
        base_c = np.repeat(rng.integers(0, R, size=T // 20 + 1), 20)[:T]
        base_k = base_c.copy()
        for s0 in rng.integers(0, T - 30, size=3):
            base_k[s0:s0 + 25] = (base_c[s0:s0 + 25] + 1 + rng.integers(0, R - 1)) % R
        seeds = []
        for sd in range(n_seeds):
            flip = rng.random((T, B)) < 0.04
            rc = np.where(flip, rng.integers(0, R, size=(T, B)), base_c[:, None])
            rk = np.where(rng.random((T, B)) < 0.04, rng.integers(0, R, size=(T, B)), base_k[:, None])
            merged = (rc == rk).astype(np.int16)
            dc = rng.integers(1, 400, size=(T, B)); dk = np.where(merged == 1, dc, rng.integers(1, 400, size=(T, B)))
            control = np.stack([dc, rc], -1).astype(np.int16); case = np.stack([dk, rk], -1).astype(np.int16)
            np.savez_compressed(os.path.join(d, f"optimal_backward_particles_merged_state_{N}_{sd}"), merged)
            np.savez_compressed(os.path.join(d, f"optimal_backward_particles_control_state_{N}_{sd}"), control)
            np.savez_compressed(os.path.join(d, f"optimal_backward_particles_case_state_{N}_{sd}"), case)
            seeds.append((merged, control, case))
        out.append(dict(positions=pos, seeds=seeds))
    return out
