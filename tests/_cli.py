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
