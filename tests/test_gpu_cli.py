"""The `hygeia` front end end to end on the GPU: files in the reference's formats in, files in the reference's formats out,
and the numbers in them equal to what the operator mirror returns for the same inputs."""
import gzip
import os

import numpy as np
import pytest

from _cli import read_csv, run, write_preprocess_style, write_simulate_style

pytestmark = pytest.mark.gpu

MU = "0.95,0.05,0.80,0.20,0.50,0.50"
SIGMA = "0.05,0.05,0.1,0.1,0.1,0.2886751"


def _write_single_group_inputs(d, ch):
    write_preprocess_style(d / "positions_1.txt.gz", ch["positions"])
    write_preprocess_style(d / "n_total_reads_1.txt.gz", ch["n_total"].T)          # site per row
    write_preprocess_style(d / "n_methylated_reads_1.txt.gz", ch["n_meth"].T)


def test_estimate_regimes_cli_matches_operator(tmp_path):
    from hygeia_b200 import model, synthetic
    from hygeia_b200.single_group import run_online_combined_inference
    T, S = 1500, 3
    ch = synthetic.make_chain(T, S, seed=31)
    _write_single_group_inputs(tmp_path, ch)
    # p / kappa / omega as ESTIMATE_PARAMETERS leaves them for ESTIMATE_REGIMES (3_estimate_regimes.nf:35-46)
    rng = np.random.default_rng(2)
    p = rng.random((6, 6)); np.fill_diagonal(p, 0.0); p /= p.sum(1, keepdims=True)
    omega = np.array(model.DEFAULT_OMEGA)
    with open(tmp_path / "p.csv", "w") as f:
        f.write(",".join(f"regime_{i}" for i in range(1, 7)) + "\n")
        for row in p:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    (tmp_path / "kappa.csv").write_text("kappa\n" + "\n".join(["2"] * 6) + "\n")
    (tmp_path / "omega.csv").write_text("omega\n" + "\n".join(repr(float(v)) for v in omega) + "\n")
    out = tmp_path / "nextflow_output" / "case_regimes_1.csv.gz"       # the directory does not exist yet (B1''')
    run("estimate_parameters_and_regimes", "--mu", MU, "--sigma", SIGMA, "--u", 3, "--p_input_csv_file", tmp_path / "p.csv",
        "--kappa_input_csv_file", tmp_path / "kappa.csv", "--omega_input_csv_file", tmp_path / "omega.csv",
        "--n_methylated_reads_csv_file", tmp_path / "n_methylated_reads_1.txt.gz", "--genomic_positions_csv_file", tmp_path / "positions_1.txt.gz",
        "--n_total_reads_csv_file", tmp_path / "n_total_reads_1.txt.gz", "--regime_probabilities_csv_file", out,
        "--estimate_regime_probabilities", "--randomise_rng_seed", "FALSE", "--rng_seed", 5)
    header, rows = read_csv(out)
    assert header == ["genomic_position"] + [f"regime_{i}" for i in range(1, 7)]
    assert len(rows) == T - 1                                           # the first site is eaten as the header (SURVEY C-1)
    for c in range(7):                                                  # format(): one width per column, right-justified
        assert len({len(r[c]) for r in rows}) == 1
    got = np.array([[float(x) for x in r] for r in rows])
    # the same call through the operator mirror; the column-major quirk of the theta conversion applies (SURVEY C-3)
    vartheta, _ = model.get_known_parameters(u=3)
    theta = model.convert_model_parameters_to_theta(p, omega)
    want = run_online_combined_inference(vartheta, theta, ch["positions"][1:], ch["n_total"][:, 1:], ch["n_meth"][:, 1:], rng_seed=5)
    w = want["regimeProbabilityEstimates"]
    assert np.array_equal(got[:, 0], w[:, 0])
    # 7 significant digits of the smallest entry in each column survive the formatting
    assert np.allclose(got[:, 1:], w[:, 1:], rtol=1e-6, atol=1e-12)
    assert (got[:, 1:].argmax(1) == ch["regimes"][1:]).mean() > 0.9

    # make_bed_file on that output (bin/make_bed_file)
    bed = tmp_path / "beds" / "case_1.bed"
    run("make_bed_file", "--chr", 1, "--regimes_file", out, "--output_file", bed)
    lines = bed.read_text().strip().split("\n")
    assert len(lines) == T - 1
    f = lines[10].split("\t")
    gp = int(got[10, 0])
    assert f[0] == "1" and int(f[1]) == gp - 1 and int(f[2]) == gp + 1 and f[5] == "." and int(f[6]) == gp - 1 and int(f[7]) == gp + 1
    assert f[3] == f"regime_{got[10, 1:].argmax() + 1}" and abs(float(f[4]) - got[10, 1:].max()) < 1e-12
    assert f[8] == ["248,118,109", "183,159,0", "0,186,56", "0,191,196", "97,156,255", "245,100,227"][got[10, 1:].argmax()]


def test_estimate_parameters_cli_outputs(tmp_path):
    from hygeia_b200 import synthetic
    T, S = 1201, 3
    ch = synthetic.make_chain(T, S, seed=32)
    # files with a header, as `hygeia simulate_data` writes them: nothing is dropped
    write_simulate_style(tmp_path / "pos.csv.gz", ch["positions"], ["genomic_positions"])
    write_simulate_style(tmp_path / "nt.csv.gz", ch["n_total"].T, [f"sample_{i + 1}" for i in range(S)])
    write_simulate_style(tmp_path / "nm.csv.gz", ch["n_meth"].T, [f"sample_{i + 1}" for i in range(S)])
    o = tmp_path / "nextflow_output"
    run("estimate_parameters_and_regimes", "--mu", MU, "--sigma", SIGMA, "--u", 3,
        "--n_methylated_reads_csv_file", tmp_path / "nm.csv.gz", "--genomic_positions_csv_file", tmp_path / "pos.csv.gz",
        "--n_total_reads_csv_file", tmp_path / "nt.csv.gz", "--regime_probabilities_csv_file", o / "regimes_1.csv.gz",
        "--theta_trace_csv_file", o / "theta_trace_1.csv.gz", "--p_csv_file", o / "p_1.csv.gz", "--kappa_csv_file", o / "kappa_1.csv.gz",
        "--omega_csv_file", o / "omega_1.csv.gz", "--theta_file", o / "theta_1.csv.gz",
        "--estimate_regime_probabilities", "--estimate_parameters", "--randomise_rng_seed", "FALSE", "--rng_seed", 3)
    h, rows = read_csv(o / "theta_trace_1.csv.gz")
    assert h == [f"theta_{i}" for i in range(1, 37)] and len(rows) == T
    trace = np.array([[float(x) for x in r] for r in rows])
    assert np.isfinite(trace).all() and np.abs(trace[-1] - trace[0]).max() > 1e-3          # six ADAM steps happened
    h, rows = read_csv(o / "p_1.csv.gz")
    p = np.array([[float(x) for x in r] for r in rows])
    assert h == [f"regime_{i}" for i in range(1, 7)] and p.shape == (6, 6)
    assert np.allclose(p.sum(1), 1.0) and np.all(np.diag(p) == 0.0)
    h, rows = read_csv(o / "omega_1.csv.gz")
    assert h == ["omega"] and len(rows) == 6
    assert np.allclose([float(r[0]) for r in rows], 1 / (1 + np.exp(-trace[-1, 30:])))
    h, rows = read_csv(o / "kappa_1.csv.gz")
    assert h == ["kappa"] and [float(r[0]) for r in rows] == [2.0] * 6
    h, rows = read_csv(o / "theta_1.csv.gz")
    assert h == ["data"] and np.array_equal([float(r[0]) for r in rows], trace[-1])         # shortest round-trip digits
    h, rows = read_csv(o / "regimes_1.csv.gz")
    assert len(rows) == T


def test_infer_cli_matches_mirror(tmp_path):
    from _tg_case import make_case
    from hygeia_b200 import two_group
    c = make_case(260, 2, seed=9)
    d = tmp_path / "preprocessed_data"; d.mkdir()
    sg = tmp_path / "single_group_estimation"; sg.mkdir()
    pos = 10000 + np.cumsum(np.random.default_rng(1).integers(2, 200, size=c["T"]))
    write_preprocess_style(d / "positions_21.txt.gz", pos)
    write_preprocess_style(d / "n_total_reads_control_21.txt.gz", c["nt_c"].T)
    write_preprocess_style(d / "n_methylated_reads_control_21.txt.gz", c["nm_c"].T)
    write_preprocess_style(d / "n_total_reads_case_21.txt.gz", c["nt_k"].T)
    write_preprocess_style(d / "n_methylated_reads_case_21.txt.gz", c["nm_k"].T)
    with gzip.open(sg / "theta_21.csv.gz", "wt") as f:
        f.write("data\n" + "\n".join(repr(float(v)) for v in c["theta"]) + "\n")
    res = tmp_path / "chrom_21_1_4"
    # batch 1 of segments of 100 sites with a 20-site halo: window [80, 220), kept [100, 200)
    run("infer", "--mu", MU, "--sigma", SIGMA, "--chrom", 21, "--single_group_dir", sg, "--data_dir", d, "--results_dir", res,
        "--seed", 4, "--batch", 1, "--segment_size", 100, "--buffer_size", 20)
    out = res / "chrom_21_1"
    names = sorted(os.listdir(out))
    assert names == sorted(["flags4.txt", "observations_control.csv.gz", "observations_case.csv.gz", "n_total_reads_control.csv.gz",
                            "n_total_reads_case.csv.gz", "positions.csv.gz", "optimal_backward_particles_merged_state_2400_4.npz",
                            "optimal_backward_particles_control_state_2400_4.npz", "optimal_backward_particles_case_state_2400_4.npz",
                            "optimal_split_probs_2400_4.npz", "optimal_regime_probs_2400_4.npz", "log_normalizing_constants_optimal_4.txt",
                            "optimal_time_4.txt", "optimal_time_backward_4.txt"])
    idx, ret = two_group.segment_index(1, 100, 20, c["T"])
    assert (idx[0], idx[-1], ret[0], ret[-1]) == (80, 219, 20, 119)
    merged = np.load(out / "optimal_backward_particles_merged_state_2400_4.npz")["arr_0"]
    control = np.load(out / "optimal_backward_particles_control_state_2400_4.npz")["arr_0"]
    case = np.load(out / "optimal_backward_particles_case_state_2400_4.npz")["arr_0"]
    split = np.load(out / "optimal_split_probs_2400_4.npz")["arr_0"]
    reg = np.load(out / "optimal_regime_probs_2400_4.npz")["arr_0"]
    assert merged.shape == (100, 25) and merged.dtype == np.int16 and control.shape == (100, 25, 2) and case.shape == (100, 25, 2)
    assert split.shape == (140,) and split.dtype == np.float32 and reg.shape == (140, 12)      # halo kept, as in the reference
    # the echo files: np.savetxt default format of the int16 window
    echoed = np.loadtxt(out / "n_total_reads_case.csv.gz", delimiter=",")
    assert np.array_equal(echoed, c["nt_k"].T[100:200])
    assert gzip.open(out / "positions.csv.gz", "rt").readline().strip() == "%.18e" % pos[100]
    # the same segment through the host mirror (chain id = batch index in the CLI)
    sl = slice(80, 220)
    s = two_group.TwoGroupSession(0)
    s.set_emission_model(c["mu"], c["sigma"], 3)
    s.add_dataset(c["nt_c"][:, sl], c["nm_c"][:, sl]); s.add_dataset(c["nt_k"][:, sl], c["nm_k"][:, sl])
    s.emission()
    logp, om = two_group.control_group_parameters(c["theta"], 6)
    s.set_two_group_model(logp, om, np.full(6, 0.8), 3, 50, 25, t_max=140)
    want = s.run([dict(control_dataset=0, case_dataset=1, T=140, seed=4, chain_id=1)])[0]
    s.close()
    tr = want["trajectories"]
    assert np.array_equal(merged, tr[20:120, :, 0]) and np.array_equal(control, tr[20:120, :, 1:3]) and np.array_equal(case, tr[20:120, :, 3:5])
    sp, rg = two_group.summarise(tr, 6)
    assert np.array_equal(split, sp) and np.allclose(reg, rg)
    txt = (out / "log_normalizing_constants_optimal_4.txt").read_text()
    assert txt.startswith("{2400: ") and float(txt.strip()[7:-1]) == want["log_normalizing_constant"]
    assert (out / "optimal_time_backward_4.txt").read_text() == "{}\n"
    # a batch index beyond the chromosome exits 0 without results (run_inference_two_groups.py:195-197)
    r = run("infer", "--mu", MU, "--sigma", SIGMA, "--chrom", 21, "--single_group_dir", sg, "--data_dir", d, "--results_dir", tmp_path / "x",
            "--seed", 4, "--batch", 7, "--segment_size", 100, "--buffer_size", 20)
    assert "Batch index is too large" in r.stdout


def test_estimate_regimes_cli_segmented_extension(tmp_path):
    """--segment_sites (an extension; default 0 = the reference's sequential run): same regimes file to the printed digits."""
    from hygeia_b200 import synthetic
    T, S = 24000, 2
    ch = synthetic.make_chain(T, S, seed=33)
    _write_single_group_inputs(tmp_path, ch)
    outs = []
    for k, extra in enumerate(([], ["--segment_sites", 5000, "--segment_halo", 2500], ["--segment_sites", "auto"])):
        out = tmp_path / f"regimes_{k}.csv.gz"
        run("estimate_parameters_and_regimes", "--mu", MU, "--sigma", SIGMA, "--u", 3,
            "--n_methylated_reads_csv_file", tmp_path / "n_methylated_reads_1.txt.gz", "--genomic_positions_csv_file", tmp_path / "positions_1.txt.gz",
            "--n_total_reads_csv_file", tmp_path / "n_total_reads_1.txt.gz", "--regime_probabilities_csv_file", out,
            "--estimate_regime_probabilities", "--randomise_rng_seed", "FALSE", "--rng_seed", 5, *extra)
        _, rows = read_csv(out)
        outs.append(np.array([[float(x) for x in r] for r in rows]))
    assert outs[0].shape == (T - 1, 7)
    for o in outs[1:]:
        assert np.array_equal(o[:, 0], outs[0][:, 0])
        assert np.abs(o[:, 1:] - outs[0][:, 1:]).max() < 1e-8
        assert np.array_equal(o[:, 1:].argmax(1), outs[0][:, 1:].argmax(1))


def _slurp(d):
    import gzip
    out = {}
    for name in sorted(os.listdir(d)):
        p = os.path.join(d, name)
        out[name] = gzip.open(p, "rb").read() if name.endswith(".gz") else open(p, "rb").read()
    return out


def test_aggregate_and_get_dmps_write_the_reference_files(tmp_path):
    """tests/golden/frontends.npz holds every file the reference's aggregate_results.py and get_dmps.py (run unmodified,
    tests/golden/make_golden_frontends.py) write for a small results tree; `hygeia aggregate` / `hygeia get_dmps` rebuild the
    tree from the same seed and must write the same bytes (gzip members compared after decompression)."""
    from conftest import golden
    from _cli import make_infer_tree
    g = golden("frontends.npz")
    tree, agg, dmp = tmp_path / "tree", tmp_path / "agg", tmp_path / "dmp"
    make_infer_tree(str(tree), chrom="21", n_batches=2, n_seeds=3, sites_per_batch=240, n_samples=1, seed=0)
    r = run("aggregate", "--results_dir", tree, "--output_dir", agg, "--seeds", 3, "--chrom", 21, "--num_batches", 30, "--compute_freqs")
    assert "Successfully processed 2 batches" in r.stdout
    got = _slurp(agg)
    want = {k.split("/", 1)[1]: bytes(v) for k, v in g.items() if k.startswith("aggregate/")}
    assert sorted(got) == sorted(want)
    for name in want:
        assert got[name] == want[name], name
    run("get_dmps", "--results_dir", agg, "--output_dir", dmp, "--chrom", 21, "--test_regime_combinations",
        "--fdr_thresholds", 0.01, "--fdr_thresholds", 0.05, "--fdr_thresholds", 0.2)
    got = _slurp(dmp)
    want = {k.split("/", 1)[1]: bytes(v) for k, v in g.items() if k.startswith("get_dmps/")}
    assert sorted(got) == sorted(want)
    differing = [name for name in want if got[name] != want[name]]
    # weighted_FDR_procedure ranks with NumPy's unstable argsort: where ranking values tie, WHICH of the tied sites make the cut is
    # not defined by the reference; the unweighted files and the headline weighted files must be identical
    assert not [n for n in differing if not n.startswith("weighted_dmp_") or n.count("_") == 2], differing
    for name in differing:
        a, b = got[name].split(b"\n"), want[name].split(b"\n")
        assert len(a) == len(b) and a[0] == b[0], name      # same number of selected sites


def test_aggregate_keeps_every_sample_column_and_reports_missing_inputs(tmp_path):
    from _cli import make_infer_tree
    tree, agg = tmp_path / "tree", tmp_path / "agg"
    make_infer_tree(str(tree), chrom="7", n_batches=1, n_seeds=2, sites_per_batch=50, n_samples=3, seed=4)
    run("aggregate", "--results_dir", tree, "--output_dir", agg, "--seeds", 2, "--chrom", 7)
    txt = _slurp(agg)["n_total_reads_case_chrom_7.csv.gz"].decode().split("\n")
    assert txt[0] == "pos\t0\t1\t2" and len(txt) == 52
    r = run("aggregate", "--results_dir", tree, "--output_dir", agg, "--seeds", 3, "--chrom", 7, check=False)   # seed 2 was never run
    assert r.returncode != 0 and "optimal_backward_particles_merged_state_2400_2.npz" in r.stderr
    r = run("aggregate", "--results_dir", tree, "--output_dir", agg, "--chrom", "8", check=False)               # no such chromosome
    assert r.returncode == 1 and "No data was processed" in r.stdout


def test_estimate_regimes_cli_with_four_regimes(tmp_path):
    """--mu / --sigma of length 4 and --u 5: the front end, the tables and the kernels are generic in R and u."""
    from hygeia_b200 import model, synthetic
    from hygeia_b200.single_group import run_online_combined_inference
    mu, sigma, omega = (0.9, 0.1, 0.5, 0.5), (0.05, 0.05, 0.1, 0.2886751), (0.99, 0.97, 0.95, 0.9)
    R, T, S = 4, 1200, 2
    rng = np.random.default_rng(8)
    reg = synthetic.simulate_regimes(T, rng) % R
    nt, nm = synthetic.simulate_counts(reg, S, rng, mu=mu, sigma=sigma)
    ch = dict(positions=synthetic.simulate_positions(T, rng), n_total=nt, n_meth=nm)
    _write_single_group_inputs(tmp_path, ch)
    p = rng.random((R, R)); np.fill_diagonal(p, 0.0); p /= p.sum(1, keepdims=True)
    with open(tmp_path / "p.csv", "w") as f:
        f.write(",".join(f"regime_{i}" for i in range(1, R + 1)) + "\n")
        for row in p:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    (tmp_path / "kappa.csv").write_text("kappa\n" + "\n".join(["2"] * R) + "\n")
    (tmp_path / "omega.csv").write_text("omega\n" + "\n".join(repr(float(v)) for v in omega) + "\n")
    out = tmp_path / "o" / "regimes.csv.gz"
    run("estimate_parameters_and_regimes", "--mu", ",".join(map(str, mu)), "--sigma", ",".join(map(str, sigma)), "--u", 5,
        "--p_input_csv_file", tmp_path / "p.csv", "--kappa_input_csv_file", tmp_path / "kappa.csv", "--omega_input_csv_file", tmp_path / "omega.csv",
        "--n_methylated_reads_csv_file", tmp_path / "n_methylated_reads_1.txt.gz", "--genomic_positions_csv_file", tmp_path / "positions_1.txt.gz",
        "--n_total_reads_csv_file", tmp_path / "n_total_reads_1.txt.gz", "--regime_probabilities_csv_file", out,
        "--estimate_regime_probabilities", "--randomise_rng_seed", "FALSE", "--rng_seed", 5)
    header, rows = read_csv(out)
    assert header == ["genomic_position"] + [f"regime_{i}" for i in range(1, R + 1)] and len(rows) == T - 1
    got = np.array([[float(x) for x in r] for r in rows])
    vartheta, _ = model.get_known_parameters(mu, sigma, u=5)
    theta = model.convert_model_parameters_to_theta(p, np.array(omega))
    want = run_online_combined_inference(vartheta, theta, ch["positions"][1:], nt[:, 1:], nm[:, 1:], rng_seed=5)["regimeProbabilityEstimates"]
    assert np.array_equal(got[:, 0], want[:, 0]) and np.allclose(got[:, 1:], want[:, 1:], rtol=1e-6, atol=1e-12)
    assert (got[:, 1:].argmax(1) == reg[1:]).mean() > 0.9
