"""Test infrastructure: an independent pandas restatement of src/two_group/preprocess_bed.py (polars 1.8.2), used only by tests/.

PARITY UNPINNED: polars cannot be installed in the build image, so neither this file nor `hygeia preprocess` has been compared with
the script itself; both restate its data flow as read -- preprocess_bed.py:124-181 (read_bed_file: skip one row, 14 named columns,
chromosome and ref_genotype == 'CG'), :183-263 (collapse_strands: full join of + and - rows on (+).end == (-).start, nulls -> 0,
coverage-weighted percentage, sites without coverage dropped), :265-358 (process_sample_data: round(coverage x pct / 100) with
Rust's f64::round, i.e. half away from zero; full join onto the CpG list WITHOUT key coalescing, so calls at unlisted positions end
up with a null position), :360-428 (nulls -> 0, null positions dropped), :430-470 (np.savetxt(fmt='%s'))."""
import numpy as np
import pandas as pd

BED_COLUMNS = ["chr", "start", "end", "name", "score", "strand", "thickStart", "thickEnd", "itemRgb", "coverage", "percent_methylated",
               "ref_genotype", "sample_genotype", "quality_score"]


def _round_half_away(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def collapse(path, chromosome):
    bed = pd.read_csv(path, sep="\t", skiprows=1, header=None, usecols=range(14), names=BED_COLUMNS, dtype={"chr": str})
    bed = bed[(bed["chr"] == chromosome) & (bed["ref_genotype"] == "CG")]
    pos = bed[bed["strand"] == "+"]
    neg = bed[bed["strand"] == "-"]
    m = pos.merge(neg, how="outer", left_on=["chr", "end"], right_on=["chr", "start"], suffixes=("", "_neg"))
    cp = m["coverage"].fillna(0).astype(float); cn = m["coverage_neg"].fillna(0).astype(float)
    pp = m["percent_methylated"].fillna(0).astype(float); pn = m["percent_methylated_neg"].fillna(0).astype(float)
    tot = cp + cn
    start = m["start"].where(m["start"].notna(), m["start_neg"] - 1)
    out = pd.DataFrame({"start": start, "total": tot, "avg": (cp * pp + cn * pn) / tot.where(tot > 0, np.nan)})
    out = out[out["total"] > 0].sort_values("start", kind="stable")
    out["start"] = out["start"].astype(np.int64)
    return out


def preprocess(cpg_file, chromosome, control_paths, case_paths):
    """Returns dict(positions, n_methylated_reads_control, n_total_reads_control, ..._case, any_null) as the script saves them."""
    cpg = pd.read_csv(cpg_file, sep="\t", dtype={"seqID": str})
    pos0 = np.sort((cpg[cpg["seqID"] == chromosome]["start"].to_numpy() - 1).astype(np.int64), kind="stable")
    res = dict(positions=pos0)
    any_null = False
    frames = {}
    for group, paths in (("control", control_paths), ("case", case_paths)):
        if not paths:
            continue
        meth = np.full((len(pos0), len(paths)), np.nan); unmeth = np.full((len(pos0), len(paths)), np.nan)
        for s, p in enumerate(paths):
            try:
                c = collapse(p, chromosome)
            except FileNotFoundError:
                continue
            j = pd.DataFrame({"Pos0": pos0}).merge(c, how="left", left_on="Pos0", right_on="start")
            meth[:, s] = _round_half_away(j["total"] * j["avg"] / 100.0)
            unmeth[:, s] = _round_half_away(j["total"] * (100.0 - j["avg"]) / 100.0)
        any_null = any_null or bool(np.isnan(meth).any())
        frames[group] = (np.nan_to_num(meth), np.nan_to_num(unmeth))
    for group, (meth, unmeth) in frames.items():
        res[f"n_methylated_reads_{group}"] = meth
        res[f"n_total_reads_{group}"] = meth + unmeth
    res["any_null"] = any_null
    return res


def savetxt_lines(a, as_float):
    """np.savetxt(fmt='%s', delimiter=',') of an integer-valued array: '12.0' when the frame went through float64, else '12'."""
    a = np.asarray(a)
    if a.ndim == 1:
        a = a[:, None]
    f = (lambda v: repr(float(v))) if as_float else (lambda v: str(int(v)))
    return "".join(",".join(f(v) for v in row) + "\n" for row in a)
