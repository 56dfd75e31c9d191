"""The `hygeia` front end without a GPU: version sub-command, argument errors, and the file formats of the CLI contract
(R's format(scientific = FALSE), readr-style doubles, np.savetxt / np.savez_compressed, the header-consuming reader)."""
import gzip

import numpy as np
import pytest

from _cli import BIN, run, write_preprocess_style


def _fmt(*xs):
    out = run("_selftest", "format", *xs).stdout.split("\n")
    return [ln[1:-1] for ln in out if ln]


def test_version_subcommands(monkeypatch):
    for flag in ("--version", "-v", "version"):
        r = run(flag)
        assert r.stdout.startswith("hygeia version ")
    monkeypatch.setenv("HYGEIA_VERSION", "9.9.9")
    # what the Nextflow modules do: hygeia --version | sed 's/hygeia version //g'   (3_estimate_regimes.nf:48-51)
    assert run("--version").stdout.strip().replace("hygeia version ", "") == "9.9.9"


def test_unknown_command_and_flags():
    assert run("frobnicate", check=False).returncode == 2          # hygeia.docker:56-62
    assert run(check=False).returncode == 1                        # no arguments: help, exit 1 (hygeia.docker:22-26)
    r = run("estimate_parameters_and_regimes", "--no_such_flag", "1", check=False)
    assert r.returncode == 1 and "unknown flag" in r.stderr
    r = run("infer", "--multinomial", "--data_dir", "/nonexistent", check=False)
    assert r.returncode == 1 and "multinomial" in r.stderr


def test_r_format_known_answers():
    # values of format(x, scientific = FALSE) in R 4.x with options(digits = 7)
    assert _fmt(1, 10, 100) == ["  1", " 10", "100"]
    assert _fmt(0.5, 0.25) == ["0.50", "0.25"]
    assert _fmt(0.1234567891) == ["0.1234568"]
    assert _fmt(10000, 0.001) == ["10000.000", "    0.001"]
    assert _fmt(1e-10) == ["0.0000000001"]
    assert _fmt(123456789, 1234567.891) == ["123456789", "  1234568"]
    assert _fmt(0.00001234, 123) == ["  0.00001234", "123.00000000"]
    assert _fmt(0, 1, 0.9999999999) == ["0", "1", "1"]
    assert _fmt(10023, 10187, 9) == ["10023", "10187", "    9"]
    # a probability column: 7 significant digits of the smallest entry decide the decimals of the whole column
    assert _fmt(0.9999994, 0.0000006123456789) == ["0.9999994000000", "0.0000006123457"]


def test_double_writers():
    out = run("_selftest", "readr", "0.5", "-1.63828305", "123456.789", "1.5e-7", "100", "0", "-5.304002150000001").stdout.split()
    assert out == ["0.5", "-1.63828305", "123456.789", "1.5e-7", "100", "0", "-5.304002150000001"]
    for s in out:
        float(s)
    out = run("_selftest", "pyrepr", "-1234.5678", "1e-5", "1.5e16", "12.0", "0.0001", "-3462.7766193165317").stdout.split()
    assert out == [repr(float(x)) for x in ("-1234.5678", "1e-5", "1.5e16", "12.0", "0.0001", "-3462.7766193165317")]


def test_reader_consumes_first_line_as_header(tmp_path):
    # SURVEY C-1: preprocess writes no header, readr::read_csv eats the first site; the two-group reader keeps it
    p = tmp_path / "n_total_reads.txt.gz"
    a = np.arange(12, dtype=float).reshape(4, 3)
    write_preprocess_style(p, a)
    rows, cols, nhead, total = run("_selftest", "read", p, 1).stdout.split()
    assert (int(rows), int(cols), int(nhead)) == (3, 3, 3) and float(total) == a[1:].sum()
    rows, cols, nhead, total = run("_selftest", "read", p, 0).stdout.split()
    assert (int(rows), int(cols), int(nhead)) == (4, 3, 0) and float(total) == a.sum()
    # plain (not gzip) files are read too
    q = tmp_path / "plain.csv"
    q.write_text("sample_1,sample_2\n1,2\n3,4\n")
    rows, cols, nhead, total = run("_selftest", "read", q, 1).stdout.split()
    assert (int(rows), int(cols), float(total)) == (2, 2, 10.0)


def test_numpy_writers(tmp_path):
    base = tmp_path / "st"
    run("_selftest", "npz", base)
    a = np.load(str(base) + "_i2.npz")
    assert a.files == ["arr_0"] and a["arr_0"].dtype == np.int16 and a["arr_0"].tolist() == [[1, -2], [3, 4], [5, 32767]]
    b = np.load(str(base) + "_f4.npz")["arr_0"]
    assert b.dtype == np.float32 and b.tolist() == [0.25, 0.5, 1.0]
    txt = gzip.open(str(base) + "_txt.csv.gz", "rt").read()
    assert txt == "1.200000000000000000e+01,3.000000000000000000e+00\n1.000000000000000000e+04,0.000000000000000000e+00\n"


def test_no_cpu_fallback(tmp_path):
    """Without a usable GPU the compute sub-commands fail loudly (exit 1), they do not fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for name, a in (("pos", np.arange(10, 40)), ("nt", np.full((30, 2), 20)), ("nm", np.full((30, 2), 10))):
        write_preprocess_style(tmp_path / f"{name}.txt.gz", a)
    r = run("estimate_parameters_and_regimes", "--genomic_positions_csv_file", tmp_path / "pos.txt.gz", "--n_total_reads_csv_file", tmp_path / "nt.txt.gz",
            "--n_methylated_reads_csv_file", tmp_path / "nm.txt.gz", "--regime_probabilities_csv_file", tmp_path / "out" / "regimes.csv.gz",
            "--estimate_regime_probabilities", check=False)
    assert r.returncode == 1 and "GPU" in r.stderr
    assert (tmp_path / "out").is_dir()              # parents of the outputs are created before anything else (B1''')
    assert not (tmp_path / "out" / "regimes.csv.gz").exists()


@pytest.mark.parametrize("n_sites,segment,want", [(2503, 1000, 3), (2000, 1000, 3), (999, 1000, 1)])
def test_get_chrom_segments(tmp_path, n_sites, segment, want):
    """get_chrom_segments.py:27-43 (run here against the reference script itself while developing): 1 + n // segment_size rows,
    also when n is an exact multiple -- `infer` then exits 0 for the empty last batch; parents of the output are created."""
    write_preprocess_style(tmp_path / "positions_9.txt.gz", np.arange(1, n_sites + 1) * 7.0)
    out = tmp_path / "deep" / "er" / "segments.csv"
    r = run("get_chrom_segments", "--input_file", tmp_path / "positions_9.txt.gz", "--chromosome", "chr9", "--segment_size", segment, "--output_csv", out)
    assert "Segment information saved to" in r.stdout
    assert out.read_text() == "chrom,segment_index\n" + "".join(f"chr9,{i}\n" for i in range(want))


def test_npz_and_matrix_readers_take_what_numpy_and_pandas_write(tmp_path):
    """`hygeia aggregate` reads the trajectories np.savez_compressed wrote (zip64 local headers, deflate) and `hygeia get_dmps` the
    tab-separated matrices pandas wrote: the readers against files written by NumPy / pandas themselves."""
    import pandas as pd
    rng = np.random.default_rng(3)
    a = rng.integers(-300, 300, size=(40, 25, 2)).astype(np.int16)
    np.savez_compressed(tmp_path / "c.npz", a)
    np.savez(tmp_path / "u.npz", a)
    for name in ("c.npz", "u.npz"):
        descr, shape, total = run("_selftest", "loadnpz", tmp_path / name).stdout.split()
        assert (descr, shape, float(total)) == ("<i2", "40x25x2", float(a.sum()))
    f = rng.random(17).astype(np.float32)
    np.savez_compressed(tmp_path / "f.npz", f)
    descr, shape, total = run("_selftest", "loadnpz", tmp_path / "f.npz").stdout.split()
    assert (descr, shape) == ("<f4", "17") and abs(float(total) - float(f.astype(np.float64).sum())) < 1e-6
    assert run("_selftest", "loadnpz", tmp_path / "missing.npz", check=False).returncode == 1
    m = pd.DataFrame(rng.integers(0, 6, size=(30, 12)).astype(np.int8))
    pos = pd.Series(np.cumsum(rng.integers(1, 900, size=30)), name="pos")
    m.set_index(pos).to_csv(tmp_path / "m.csv.gz", sep="\t", compression="gzip")
    rows, cols, name, si, sv = run("_selftest", "readmatrix", tmp_path / "m.csv.gz").stdout.split()
    assert (int(rows), int(cols), name, int(si), int(sv)) == (30, 12, "pos", int(pos.sum()), int(m.to_numpy().sum()))
