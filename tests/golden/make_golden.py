"""Generate the committed golden vectors from the REFERENCE ITSELF (run in the build container only).

The reference's own single-group C++ headers are compiled in place (oracle/Makefile -> oracle/_ref/libhyg_ref*.so) and
driven through oracle/ref_driver.cpp; this script records inputs and outputs of a few small chains so that the GPU box
(where /root/reference does not exist) can still pin the oracle and the CUDA path to reference behaviour.

    python tests/golden/make_golden.py

Writes tests/golden/sg_*.npz.  `ref_fast` = the reference's build flags (-O3 -ffast-math ...), `ref_strict` = -O2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hygeia_b200 import model, philox, synthetic  # noqa: E402
from _oracle import Ref  # noqa: E402

CASES = [
    # name, T, S, data seed, lambda, p_missing, n_particles, epsilon
    ("default_s4", 700, 4, 101, 30.0, 0.05, 250, 0.01),
    ("sparse_s1", 600, 1, 102, 8.0, 0.20, 250, 0.01),
    ("dense_s16", 400, 16, 103, 30.0, 0.05, 250, 0.01),
    ("few_particles", 500, 2, 104, 20.0, 0.05, 60, 0.001),
]


# Long one-sample chains (the production shape of the single-group workflow): exact ties between weights are
# systematic there and the reference breaks them by whatever order std::sort leaves, so these pin
# (i) the oracle's reference tie order bit for bit and (ii) the CUDA path up to the first site where a tie decided a fate.
LONG_CASES = [
    # name, T, S, data seed, lambda, p_missing
    ("long_s1_sparse", 12000, 1, 301, 8.0, 0.20),
    ("long_s1_l10", 12000, 1, 303, 10.0, 0.05),
    ("long_s1_l30", 12000, 1, 304, 30.0, 0.05),
    ("long_s2_sparse", 12000, 2, 305, 8.0, 0.20),
]


def main_long():
    from _oracle import Oracle
    strict = Ref("_strict")
    orc = Oracle()
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    for name, T, S, seed, lam, pmiss in LONG_CASES:
        ch = synthetic.make_chain(T, S, seed=seed, lam=lam, p_missing=pmiss)
        u = philox.uniforms_by_site(seed, 0, T)
        r = strict.run(vartheta, theta, ch["n_total"], ch["n_meth"], ch["positions"], uniforms=u, stepwise=True)
        o = orc.run(vartheta, theta, u, ch["n_total"], ch["n_meth"], ch["positions"], tie_order="reference")
        assert np.array_equal(o["logz"], r["logz"]) and np.array_equal(o["regime_probs"], r["regime_probs"], equal_nan=True)
        np.savez_compressed(os.path.join(HERE, f"sg_{name}.npz"), vartheta=vartheta, theta=theta, n_total=ch["n_total"], n_meth=ch["n_meth"],
                            positions=ch["positions"], philox_seed=seed, ref_strict_regime_probs=r["regime_probs"][:, 1:],
                            ref_strict_logz=r["logz"], ref_strict_drew_uniform=r["drew_uniform"], ref_strict_finalised_at=r["finalised_at"],
                            ref_tie_flags=o["tie_flags"])
        rel = np.nonzero(o["tie_flags"] & 2)[0]
        print(name, "logZ_T", repr(r["logz"][-1]), "sites where a tie decided a fate:", len(rel), "first", rel[:3])


def main():
    strict, fast = Ref("_strict"), Ref("")
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    alpha, beta = model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA)
    for name, T, S, seed, lam, pmiss, npart, eps in CASES:
        ch = synthetic.make_chain(T, S, seed=seed, lam=lam, p_missing=pmiss)
        u = philox.uniforms_by_site(seed, 0, T)
        out = {}
        for tag, lib in (("strict", strict), ("fast", fast)):
            r = lib.run(vartheta, theta, ch["n_total"], ch["n_meth"], ch["positions"], uniforms=u, stepwise=True,
                        n_particles=npart, epsilon=eps)
            r0 = lib.run(vartheta, theta, ch["n_total"], ch["n_meth"], ch["positions"], uniforms=u, stepwise=False,
                         n_particles=npart, epsilon=eps)
            assert np.array_equal(r["regime_probs"], r0["regime_probs"]), "restated outer loop differs from the reference's run()"
            out[f"ref_{tag}_regime_probs"] = r["regime_probs"]
            out[f"ref_{tag}_logz"] = r["logz"]
            out[f"ref_{tag}_drew_uniform"] = r["drew_uniform"]
            out[f"ref_{tag}_n_pending"] = r["n_pending"]
            out[f"ref_{tag}_finalised_at"] = r["finalised_at"]
            out[f"ref_{tag}_logobs"] = lib.emission(vartheta, ch["n_total"], ch["n_meth"])
        np.savez_compressed(os.path.join(HERE, f"sg_{name}.npz"), vartheta=vartheta, theta=theta, alpha=alpha, beta=beta,
                            n_total=ch["n_total"], n_meth=ch["n_meth"], positions=ch["positions"], regimes=ch["regimes"], uniforms=u,
                            n_particles=npart, epsilon=eps, **out)
        print(name, "logZ_T strict", repr(out["ref_strict_logz"][-1]), "fast", repr(out["ref_fast_logz"][-1]),
              "draws", int(out["ref_strict_drew_uniform"].sum()), "max pending", int(out["ref_strict_n_pending"].max()))
    # sojourn tables and P/omega for the default theta and a perturbed theta (known answers of SURVEY.md appendix D-3)
    rng = np.random.default_rng(5)
    theta2 = theta + 0.3 * rng.standard_normal(theta.shape)
    tabs = {}
    for tag, th in (("default", theta), ("perturbed", theta2)):
        t = strict.tables(vartheta, th, 4000)
        for k, v in t.items():
            tabs[f"{tag}_{k}"] = v
        tabs[f"{tag}_theta"] = th
    np.savez_compressed(os.path.join(HERE, "sg_tables.npz"), vartheta=vartheta, **tabs)
    print("tables written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "long":
        main_long()
    else:
        main()
        main_long()
