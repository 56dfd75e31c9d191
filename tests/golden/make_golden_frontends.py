"""Golden files for the `hygeia aggregate` / `hygeia get_dmps` front ends, produced by the REFERENCE'S OWN SCRIPTS.

    python tests/golden/make_golden_frontends.py        (build container only: needs /root/reference, pandas, absl)

src/two_group/aggregate_results.py and get_dmps.py are scripts that parse flags and do all their work at import time; they
are run here unmodified, as subprocesses, on a small results tree in the format `hygeia infer` writes
(tests/_cli.py::make_infer_tree: np.savetxt echoes + np.savez_compressed trajectories; one sample per group, which is all the
reference's `aggregate` can read).  Every file they write is stored, decompressed, in tests/golden/frontends.npz; the GPU
test rebuilds the same tree from the same seed and compares the CLI's files with these byte for byte."""
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/src/two_group"

import _cli  # noqa: E402

TREE = dict(chrom="21", n_batches=2, n_seeds=3, sites_per_batch=240, n_samples=1, seed=0)


def slurp(d):
    out = {}
    for name in sorted(os.listdir(d)):
        p = os.path.join(d, name)
        raw = gzip.open(p, "rb").read() if name.endswith(".gz") else open(p, "rb").read()
        out[name] = np.frombuffer(raw, dtype=np.uint8)
    return out


def main():
    with tempfile.TemporaryDirectory() as tmp:
        tree, agg, dmp = (os.path.join(tmp, x) for x in ("tree", "agg", "dmp"))
        _cli.make_infer_tree(tree, **TREE)
        subprocess.run([sys.executable, os.path.join(REF, "aggregate_results.py"), "--results_dir", tree, "--output_dir", agg, "--seeds", str(TREE["n_seeds"]),
                        "--chrom", TREE["chrom"], "--num_batches", "30", "--compute_freqs"], check=True, cwd=tmp, stdout=subprocess.DEVNULL)
        subprocess.run([sys.executable, os.path.join(REF, "get_dmps.py"), "--results_dir", agg, "--output_dir", dmp, "--chrom", TREE["chrom"],
                        "--test_regime_combinations", "--fdr_thresholds", "0.01", "--fdr_thresholds", "0.05", "--fdr_thresholds", "0.2"],
                       check=True, cwd=REF, stdout=subprocess.DEVNULL)
        files = {f"aggregate/{k}": v for k, v in slurp(agg).items()}
        files.update({f"get_dmps/{k}": v for k, v in slurp(dmp).items()})
    np.savez_compressed(os.path.join(HERE, "frontends.npz"), **files)
    n_rows = {k: int((v == 10).sum()) - 1 for k, v in files.items() if k.startswith("get_dmps/") and "_0." in k and k.count("_") <= 2}
    print(len(files), "files;", "rows selected:", n_rows)


if __name__ == "__main__":
    main()
