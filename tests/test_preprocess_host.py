"""`hygeia preprocess` (host-side ETL, no GPU) against an independent pandas restatement of preprocess_bed.py and hand-made cases.
Parity with the script itself is unpinned (polars is not installable here): DESIGN.md section 7."""
import gzip
import os
import sys

import numpy as np

from _cli import run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import preprocess_oracle as po  # noqa: E402

HEADER = "\t".join(po.BED_COLUMNS) + "\n"


def _bed_rows(rng, chrom, sites, p_both=0.6, p_plus=0.2, extra_cols=False):
    rows = []
    for s in sites:                      # s = 0-based start of the C of the CpG on the + strand
        kind = rng.random()
        cov_p, cov_n = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        pm_p, pm_n = round(float(rng.random() * 100), 1), round(float(rng.random() * 100), 1)

        def row(start, strand, cov, pm, gt="CG"):
            r = [chrom, start, start + 1, ".", 0, strand, start, start + 1, "0,0,0", cov, pm, gt, "CG", 30]
            if extra_cols:
                r += ["extra", 1]
            return "\t".join(str(x) for x in r) + "\n"
        if kind < p_both:
            rows += [row(s, "+", cov_p, pm_p), row(s + 1, "-", cov_n, pm_n)]
        elif kind < p_both + p_plus:
            rows.append(row(s, "+", cov_p, pm_p))
        else:
            rows.append(row(s + 1, "-", cov_n, pm_n))
    rows.append("\t".join(str(x) for x in ["chrOther", 5, 6, ".", 0, "+", 5, 6, "0,0,0", 9, 50.0, "CG", "CG", 30] + (["e", 1] if extra_cols else [])) + "\n")
    rows.append("\t".join(str(x) for x in [chrom, 7, 8, ".", 0, "+", 7, 8, "0,0,0", 9, 50.0, "CHH", "CHH", 30] + (["e", 1] if extra_cols else [])) + "\n")
    return rows


def _setup(tmp_path, seed, n_control, n_case, chrom="chr21", extra_cols=False, all_covered=False):
    rng = np.random.default_rng(seed)
    listed = np.sort(rng.choice(np.arange(100, 20000, 2), size=400, replace=False))
    cpg = tmp_path / "cpg.tsv"
    with open(cpg, "w") as f:
        f.write("seqID\tstart\tend\n")
        for s in listed:
            f.write(f"{chrom}\t{s + 1}\t{s + 2}\n")          # 1-based start in the CpG list
        f.write("chrOther\t11\t12\n")
    paths = {}
    for g, n in (("control", n_control), ("case", n_case)):
        paths[g] = []
        for k in range(n):
            sites = listed if all_covered else np.sort(np.concatenate([rng.choice(listed, size=330, replace=False), [50, 60000]]))   # two unlisted
            p = tmp_path / f"{g}_{k}.bed"
            with open(p, "w") as f:
                f.write(HEADER)
                f.writelines(_bed_rows(rng, chrom, sites, extra_cols=extra_cols) if not all_covered else
                             [f"{chrom}\t{s}\t{s + 1}\t.\t0\t+\t{s}\t{s + 1}\t0,0,0\t{10 + (s % 7)}\t{(s % 11) * 10.0}\tCG\tCG\t30\n" for s in sites])
            paths[g].append(str(p))
    return str(cpg), paths, chrom


def _read(path):
    return gzip.open(path, "rt").read()


def _check(out, chrom, want):
    fl = want["any_null"]
    assert _read(out / f"positions_{chrom}.txt.gz") == po.savetxt_lines(want["positions"], False)
    assert _read(out / f"cpg_sites_merged_{chrom}.txt.gz") == f"{len(want['positions'])}\n"
    for g in ("control", "case"):
        for k in ("n_methylated_reads", "n_total_reads"):
            p = out / f"{k}_{g}_{chrom}.txt.gz"
            if f"{k}_{g}" in want:
                assert _read(p) == po.savetxt_lines(want[f"{k}_{g}"], fl), p
            else:
                assert not p.exists()


def test_preprocess_matches_the_restatement(tmp_path):
    cpg, paths, chrom = _setup(tmp_path, 1, 3, 2, extra_cols=True)
    out = tmp_path / "out"
    args = ["preprocess", "--cpg_file_path", cpg, "--output_path", out, "--chromosome", chrom]
    for p in paths["control"]:
        args += ["--control_data_path", p]
    for p in paths["case"]:
        args += ["--case_data_path", p]
    r = run(*args)
    assert "Successfully processed 400 CpG sites" in r.stdout
    want = po.preprocess(cpg, chrom, paths["control"], paths["case"])
    assert want["any_null"] and (want["n_total_reads_control"] >= want["n_methylated_reads_control"]).all()
    assert (want["n_total_reads_control"] > 0).mean() > 0.5
    _check(out, chrom, want)


def test_preprocess_one_group_integer_output_and_missing_file(tmp_path):
    # every listed site covered in every sample: the frame never holds a null and the script writes integers ("12", not "12.0")
    cpg, paths, chrom = _setup(tmp_path, 2, 0, 2, all_covered=True)
    out = tmp_path / "out"
    run("preprocess", "--cpg_file_path", cpg, "--output_path", out, "--chromosome", chrom, "--case_data_path", paths["case"][0],
        "--case_data_path", paths["case"][1], "--case_id_names", "a", "--case_id_names", "b")
    want = po.preprocess(cpg, chrom, [], paths["case"])
    assert not want["any_null"]
    _check(out, chrom, want)
    # a sample file that does not exist becomes a column of zeros (preprocess_bed.py:289-296)
    out2 = tmp_path / "out2"
    run("preprocess", "--cpg_file_path", cpg, "--output_path", out2, "--chromosome", chrom, "--case_data_path", paths["case"][0],
        "--case_data_path", str(tmp_path / "nope.bed"))
    want2 = po.preprocess(cpg, chrom, [], [paths["case"][0], str(tmp_path / "nope.bed")])
    assert want2["any_null"] and (want2["n_total_reads_case"][:, 1] == 0).all()
    _check(out2, chrom, want2)


def test_preprocess_known_answers_and_errors(tmp_path):
    cpg = tmp_path / "cpg.tsv"
    cpg.write_text("seqID\tstart\tend\n7\t101\t102\n7\t201\t202\n7\t301\t302\n8\t5\t6\n")
    bed = tmp_path / "s.bed"
    bed.write_text(HEADER +
                   "7\t100\t101\t.\t0\t+\t100\t101\t0,0,0\t2\t25.0\tCG\tCG\t30\n"      # + only: 2 reads, 25 % -> round(0.5) = 1 methylated, round(1.5) = 2 unmethylated
                   "7\t201\t202\t.\t0\t-\t201\t202\t0,0,0\t10\t50.0\tCG\tCG\t30\n"     # - only: position 201 - 1 = 200
                   "7\t300\t301\t.\t0\t+\t300\t301\t0,0,0\t0\t0.0\tCG\tCG\t30\n"       # both strands without coverage: dropped -> null -> 0
                   "7\t301\t302\t.\t0\t-\t301\t302\t0,0,0\t0\t0.0\tCG\tCG\t30\n")
    out = tmp_path / "o"
    run("preprocess", "--cpg_file_path", cpg, "--output_path", out, "--chromosome", "7", "--control_data_path", bed)
    assert _read(out / "positions_7.txt.gz") == "100\n200\n300\n"
    assert _read(out / "n_methylated_reads_control_7.txt.gz") == "1.0\n5.0\n0.0\n"
    assert _read(out / "n_total_reads_control_7.txt.gz") == "3.0\n10.0\n0.0\n"         # the script's rounding: 1 + 2 reads from a coverage of 2
    assert run("preprocess", "--cpg_file_path", cpg, "--output_path", out, "--chromosome", "9", "--control_data_path", bed, check=False).returncode == 1
    assert run("preprocess", "--cpg_file_path", cpg, "--output_path", out, check=False).returncode == 1
    assert run("preprocess", "--output_path", out, "--control_data_path", bed, check=False).returncode == 1
