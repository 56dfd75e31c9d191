#!/usr/bin/env python
"""bench.py -- Hygeia single-group whole-genome seed sweep on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this framework (CUDA)
    python bench.py --impl reference --gpus 1 --steps K --warmup W # the reference's own CPU code (oracle/_ref), bounded sample
    torchrun ... bench.py --gpus N ...                             # one rank per GPU, weak scaling (2 seeds per GPU)

A "step" is one pass of the hot path over the whole synthetic genome: K1 (emission table of every chromosome) + K2 (the
recursion for every chromosome x seed chain).  Workload at N = 1 = BASELINE.json configs[1]: ~28 M CpG sites in 22
synthetic chromosomes, 32 samples sharing one regime path, 2 seeds.  N GPUs run seeds {2r, 2r+1} of the same genome on
rank r (configs[2] at N = 8), no data-path collective; per-chromosome log-evidences are all-reduced over NCCL at the end
of each step.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CpG-site x sample updates/sec per seed sweep"
UNIT = "site*sample*seed/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--total-sites", type=int, default=28_000_000)
    p.add_argument("--samples", type=int, default=32)
    p.add_argument("--seeds-per-gpu", type=int, default=2)
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--cpu-sites", type=int, default=1500, help="sites per chain of the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--emission-only", action="store_true", help="time K1 alone (kernel tuning aid; not a bench line)")
    p.add_argument("--segment-sites", type=int, default=-1,
                   help="K2 execution: -1 = segmented, size chosen by the library (default); 0 = whole chains (the reference's "
                        "sequential run); N = segments of <= N sites")
    p.add_argument("--halo", type=int, default=5000, help="left/right halo of a segment (sites)")
    p.add_argument("--no-e2e-pipeline", dest="e2e_pipeline", action="store_false",
                   help="end-to-end leg with ONE session (copies not overlapped with the recursion)")
    p.add_argument("--no-mode-check", action="store_true", help="skip the full-size comparison of segmented vs whole-chain results")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------------------------
# synthetic genome: regime paths on the host, counts on the device
# ------------------------------------------------------------------------------------------------------------------
def make_genome(total_sites, S, device, data_seed=20261018):
    import torch
    from hygeia_b200 import model, synthetic
    alpha, beta = model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA)
    lens = synthetic.chromosome_lengths(total_sites)
    g = torch.Generator(device=device)
    g.manual_seed(data_seed)
    al = torch.tensor(alpha, device=device, dtype=torch.float32)
    be = torch.tensor(beta, device=device, dtype=torch.float32)
    chroms = []
    for c, T in enumerate(lens):
        rng = np.random.default_rng(data_seed + c)
        regimes = synthetic.simulate_regimes(T, rng)
        pitch = (T + 7) // 8 * 8
        r = torch.from_numpy(regimes.astype(np.int64)).to(device)
        n = torch.poisson(torch.full((S, T), 30.0, device=device), generator=g)
        n = n * (torch.rand((S, T), device=device, generator=g) >= 0.05)
        a_t = al[r].expand(S, T).contiguous()
        b_t = be[r].expand(S, T).contiguous()
        ga = torch._standard_gamma(a_t, generator=g)
        gb = torch._standard_gamma(b_t, generator=g)
        pm = ga / (ga + gb)
        x = torch.binomial(n, pm.clamp(0, 1), generator=g)
        nt = torch.zeros((S, pitch), dtype=torch.uint16, device=device)
        nm = torch.zeros((S, pitch), dtype=torch.uint16, device=device)
        nt[:, :T] = n.to(torch.int32).to(torch.uint16)
        nm[:, :T] = x.to(torch.int32).to(torch.uint16)
        pos = torch.from_numpy(synthetic.simulate_positions(T, rng).astype(np.int64))
        chroms.append(dict(T=T, pitch=pitch, n_total=nt, n_meth=nm, positions=pos.to(torch.int32), regimes=regimes))
        del n, a_t, b_t, ga, gb, pm, x, r
    return chroms


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own C++ (oracle/_ref) on a bounded sample, one chain per core
# ------------------------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import Oracle, Ref
    vartheta, theta, nt, nm, u, kind = job
    t0 = time.perf_counter()
    if kind == "reference":
        Ref("").run(vartheta, theta, nt, nm, None, uniforms=u, stepwise=False)
    else:
        Oracle().run(vartheta, theta, u, nt, nm)
    return time.perf_counter() - t0


def cpu_baseline(chains_np, vartheta, theta, S, sites, max_procs=None):
    """chains_np: list of (n_total[S][>=sites], n_meth) numpy slices.  Returns dict for the JSON line."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import Ref
    from hygeia_b200 import philox
    kind = "reference" if Ref.available("") else "port"
    cores = os.cpu_count() or 1
    n = min(cores, len(chains_np)) if max_procs is None else min(max_procs, len(chains_np))
    jobs = []
    for i in range(n):
        nt, nm = chains_np[i]
        jobs.append((vartheta, theta, np.ascontiguousarray(nt[:, :sites]), np.ascontiguousarray(nm[:, :sites]),
                     philox.uniforms_by_site(0, i, sites), kind))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n) as pool:
        per = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    units = n * sites * S
    return {"value": units / max(per) if per else None, "unit": UNIT, "cores": n, "kind": kind,
            "sample": f"first {sites} sites of {n} chromosomes, S={S}, 1 seed each, one chain per core "
                      f"(oracle/_ref = reference headers, -O3 -ffast-math); slowest chain {max(per):.1f} s, pool wall {wall:.1f} s",
            "host_cores": cores}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hygeia_b200 import model, synthetic
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    S = args.samples
    cores = os.cpu_count() or 1
    n = min(cores, 22)
    sites = max(200, args.cpu_sites // 3)
    chains = []
    for i in range(n):
        rng = np.random.default_rng(20261018 + i)
        regimes = synthetic.simulate_regimes(sites, rng)
        chains.append(synthetic.simulate_counts(regimes, S, rng))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline(chains, vartheta, theta, S, sites)
    vals, secs = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        r = cpu_baseline(chains, vartheta, theta, S, sites)
        secs.append(time.perf_counter() - t0)
        vals.append(r["value"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"single_group whole-genome synthetic (~{args.total_sites // 1_000_000}M CpGs), {S} samples, "
                                   f"{args.seeds_per_gpu} seeds -- bounded sample: first {sites} sites of {n} chromosomes, 1 seed",
                       "samples": S},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from hygeia_b200 import model
    from hygeia_b200.single_group import Session, make_run_args

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S, n_seeds = args.samples, args.seeds_per_gpu
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    R = 6
    chroms = make_genome(args.total_sites, S, dev)
    torch.cuda.synchronize()
    total_T = sum(c["T"] for c in chroms)
    units_per_step = total_T * S * n_seeds
    seeds = [rank * n_seeds + k for k in range(n_seeds)]

    # pinned host mirrors (inputs for the end-to-end leg, outputs for both legs)
    for c in chroms:
        c["h_nt"] = torch.empty((S, c["T"]), dtype=torch.uint16).pin_memory()
        c["h_nm"] = torch.empty((S, c["T"]), dtype=torch.uint16).pin_memory()
        c["h_nt"].copy_(c["n_total"][:, :c["T"]]); c["h_nm"].copy_(c["n_meth"][:, :c["T"]])
        c["h_pos"] = c["positions"].pin_memory()
        c["out"] = [dict(probs=torch.empty((c["T"], 1 + R), dtype=torch.float64).pin_memory(),
                         logz=torch.empty(c["T"], dtype=torch.float64).pin_memory()) for _ in seeds]
    torch.cuda.synchronize()

    sess = Session(local)
    run_args = make_run_args()
    seg_request = Session.SEGMENT_AUTO if args.segment_sites < 0 else args.segment_sites
    sess.set_segmentation(seg_request, args.halo, args.halo)

    def stage(device_resident, ses=None, which=None):
        """Stage the chromosomes `which` (indices into chroms; default all) on session `ses` (default the main one)."""
        ses = ses or sess
        which = range(len(chroms)) if which is None else which
        ses.clear()
        ses.set_vartheta(vartheta)
        ses.set_theta(theta, max(c["T"] for c in chroms))
        specs = []
        for ci in which:
            c = chroms[ci]
            if device_resident:
                ds = ses.add_dataset_ptr(c["T"], S, c["n_total"].data_ptr(), c["n_meth"].data_ptr(), True, c["pitch"])
            else:
                ds = ses.add_dataset_ptr(c["T"], S, c["h_nt"].data_ptr(), c["h_nm"].data_ptr(), False, c["T"])
            for k, sd in enumerate(seeds):
                specs.append(dict(dataset=ds, seed=sd, chain_id=ci, positions=c["h_pos"].data_ptr(),
                                  regime_probs=c["out"][k]["probs"].data_ptr(), logz=c["out"][k]["logz"].data_ptr()))
        ses.set_chains(specs)
        return len(specs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    evid = torch.zeros(len(chroms), dtype=torch.float64, device=dev)

    def step_device():
        sess.emission()
        sess.filter(run_args)
        sess.sync()
        if world > 1:  # cross-shard reduction of the per-chromosome sufficient statistic (sum over seeds of log Z_T)
            dist.all_reduce(evid)
        return sess.timings()

    # ---- device-resident leg: inputs already in HBM ----
    n_chains = stage(True)
    if args.emission_only:
        ms = []
        for _ in range(3 + args.steps):
            sess.emission(); sess.sync(); ms.append(sess.timings()["ms_emission"])
        alg = total_T * S * 4 + total_T * R * 8
        best = min(ms[3:]); avg = float(np.mean(ms[3:]))
        print(json.dumps({"emission_only_ms": ms, "GBps_avg": alg / avg / 1e6, "GBps_best": alg / best / 1e6,
                          "frac_of_6538.9": alg / avg / 1e6 / 6538.9}))
        return
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    ev_ms, em_ms, f_ms = [], [], []
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tm = step_device()
            ev_ms.append(tm["ms_emission"] + tm["ms_filter"]); em_ms.append(tm["ms_emission"]); f_ms.append(tm["ms_filter"])
        barrier()
        wall = time.perf_counter() - t0
    launches = tm["emission_launches"] + tm["filter_launches"]
    n_units, seg_sites, workers = sess.filter_units(with_segment_sites=True)
    st = sess.download()   # also brings the per-chain status words back (outside the timed region)
    stepped = int(sum(x[3] for x in st))
    forced_halo = int(sum(x[2] for x in st)); forced_lag = int(sum(x[0] for x in st))
    dev_s = sum(ev_ms) / 1000.0
    t_all = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t_all[0]), float(t_all[1])
    value = units_per_step * world * args.steps / dev_s

    # ---- roofline: algorithmic bytes / CUDA-event time of each kernel (events are recorded by the library on its own stream) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (burst copy figure)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    longest = max(c["T"] for c in chroms)
    # K2, the dominant kernel (99.7 % of the step in profiles/r01_launches_4M_segmented.csv): per site and chain it reads the R emission
    # log-densities and writes the R posterior probabilities and log Z_t.  It is a sequential recursion (one CTA per chain, 250
    # particles), bound by per-site latency, not by HBM -- the fraction below says how far from HBM it is, the
    # us/site figure is the number to optimise.
    k2_alg = total_T * n_seeds * (R * 8 + R * 8 + 8)   # + the running log-evidence (owned rows only: halo steps are overhead)
    k2_ms = float(np.mean(f_ms))
    k2_ach = k2_alg / (k2_ms / 1000.0) / 1e9
    # DRAM bytes per site measured by ncu --set full (profiles/r01_k2_final_details.txt, r01_k1_final_details.txt: 4M-site
    # capture, 2 seeds; K2 332.4 MB for 8M owned site-chains -- the posterior rows go to pinned host memory, not to DRAM), scaled
    K2_DRAM_B_PER_SITE_CHAIN, K1_DRAM_B_PER_SITE = 41.5, 176.2
    roofline = {"bound": "hbm", "kernel": "sg_filter_kernel<6,0> (K2: particle filter + fixed-lag smoother), 1 persistent launch per step, "
                                          "one CTA per (chain, segment) unit",
                "share_of_step": k2_ms / (k2_ms + float(np.mean(em_ms))),
                "achieved": k2_ach, "peak": peak, "unit": "GB/s", "frac": k2_ach / peak, "peak_source": peak_source,
                "algorithmic_bytes_per_launch": k2_alg, "ms_per_launch": k2_ms,
                "traffic": K2_DRAM_B_PER_SITE_CHAIN * total_T * n_seeds,
                "traffic_source": "ncu dram__bytes_read+write of a 4M-site capture, per owned site-chain, scaled to this launch "
                                  "(posterior rows leave over PCIe into pinned host memory and are not DRAM traffic)",
                "latency_bound": {"chains": n_chains, "units": n_units, "segment_sites": seg_sites, "halo_sites": args.halo if seg_sites else 0,
                                  "resident_ctas": workers, "sms": 148, "longest_chain_sites": longest,
                                  "sites_stepped_incl_halos": stepped, "halo_overhead": stepped / float(total_T * n_seeds) - 1.0,
                                  "us_per_site_per_cta": 1000.0 * k2_ms * workers / max(stepped, 1),
                                  "sites_forced_at_segment_end": forced_halo, "sites_forced_lag_set_full": forced_lag,
                                  "note": "a strictly sequential recursion per unit: the bound is per-site latency x units in flight, "
                                          "HBM is idle; see DESIGN.md section 4 (K2) and 6"}}
    # K1, the HBM-streaming kernel north_star sets the roofline target for
    alg_bytes = total_T * S * 2 * 2 + total_T * R * 8
    ach = alg_bytes / (float(np.mean(em_ms)) / 1000.0) / 1e9
    roofline_emission = {"bound": "hbm", "kernel": "sg_emission_kernel<6> (K1), 1 persistent launch per step over all chromosomes",
                         "share_of_step": float(np.mean(em_ms)) / (k2_ms + float(np.mean(em_ms))),
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_source,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": float(np.mean(em_ms)),
                         "traffic": K1_DRAM_B_PER_SITE * total_T,
                         "traffic_source": "ncu dram__bytes_read+write of a 4M-site capture, per site, scaled to this launch",
                         "limiter": "shared-memory wavefronts of the fp64 table look-ups (L1 pipe 94 % busy, 59 % of wavefronts are bank conflicts)"}

    # ---- end-to-end leg: host (pinned) buffers in, host buffers out, through the public API ----
    e2e = None
    if not args.no_e2e:
        h2d = sum(2 * S * c["T"] * 2 + c["T"] * 4 * n_seeds for c in chroms)
        d2h = sum((c["T"] * (1 + R) * 8 + c["T"] * 8) * n_seeds for c in chroms)
        # Two contexts (two streams) take half of the chromosomes each, so that the host -> device copy of the second half and
        # the device -> host copy of the first half's log Z overlap the recursion of the other half.
        order = sorted(range(len(chroms)), key=lambda i: -chroms[i]["T"])
        halves = [order[0::2], order[1::2]] if args.e2e_pipeline else [list(range(len(chroms)))]
        sessions = [sess]
        if len(halves) > 1:
            s2 = Session(local)
            s2.set_segmentation(seg_request, args.halo, args.halo)
            sessions.append(s2)

        def step_e2e():
            for ses, which in zip(sessions, halves):
                stage(False, ses, which)
                ses.emission()
                ses.filter(run_args)
            for ses in sessions:
                ses.download()
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": units_per_step * world * args.e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1000.0 * e2e_s / args.e2e_steps, "steps": args.e2e_steps,
               "api": "hygeia_b200.single_group.Session: add_dataset(pinned host) -> set_chains -> emission -> filter -> download"
                      + (", two Sessions (contexts) with half of the chromosomes each so that copies overlap the other half's recursion; "
                         if len(halves) > 1 else "; ") +
                      "posterior rows are written by K2 straight into the pinned host buffers (counted in d2h_bytes_per_step), "
                      "log Z is staged in HBM and copied"}
        # sanity on the downloaded results of the last step
        p = chroms[-1]["out"][0]["probs"].numpy()
        acc = float((p[:, 1:].argmax(1) == chroms[-1]["regimes"]).mean())
        e2e["regime_call_accuracy_vs_simulated_truth_last_chromosome"] = acc
        e2e["non_finite_posterior_rows"] = int(sum(int((~np.isfinite(o["probs"].numpy())).any(1).sum()) for c in chroms for o in c["out"]))

    # ---- full-size parity of the two execution modes: segmented (timed above) vs the sequential whole-chain run ----
    mode_check = None
    if e2e is not None and seg_sites and not args.no_mode_check:
        seg_out = [[(o["probs"].numpy().copy(), o["logz"].numpy().copy()) for o in c["out"]] for c in chroms[-3:]]
        sess.set_segmentation(0, args.halo, args.halo)
        sub = chroms[-3:]   # the three shortest chromosomes, every seed: whole-chain runs of 0.4-0.75 M sites
        sess.clear(); sess.set_vartheta(vartheta); sess.set_theta(theta, max(c["T"] for c in chroms))
        specs = []
        for ci, c in enumerate(sub):
            ds = sess.add_dataset_ptr(c["T"], S, c["h_nt"].data_ptr(), c["h_nm"].data_ptr(), False, c["T"])
            for k, sd in enumerate(seeds):
                specs.append(dict(dataset=ds, seed=sd, chain_id=len(chroms) - 3 + ci, positions=c["h_pos"].data_ptr(),
                                  regime_probs=c["out"][k]["probs"].data_ptr(), logz=c["out"][k]["logz"].data_ptr()))
        sess.set_chains(specs); sess.emission(); sess.filter(run_args); sess.download()
        whole_ms = sess.timings()["ms_filter"]
        dp = dz = 0.0; calls = 0; rows = 0
        for c, so in zip(sub, seg_out):
            for o, (sp, sz) in zip(c["out"], so):
                wp, wz = o["probs"].numpy(), o["logz"].numpy()
                dp = max(dp, float(np.abs(sp[:, 1:] - wp[:, 1:]).max()))
                dz = max(dz, float((np.abs(sz - wz) / np.abs(wz)).max()))
                calls += int((sp[:, 1:].argmax(1) != wp[:, 1:].argmax(1)).sum()); rows += wp.shape[0]
        mode_check = {"compared": f"segmented vs whole-chain execution on the 3 shortest chromosomes x {n_seeds} seeds ({rows} site rows)",
                      "max_abs_posterior_diff": dp, "max_rel_logz_diff": dz, "differing_regime_calls": calls, "tolerance": 1e-6,
                      "whole_chain_us_per_site": 1000.0 * whole_ms / max(c["T"] for c in sub)}
        sess.set_segmentation(seg_request, args.halo, args.halo)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sl = [(c["h_nt"].numpy()[:, :args.cpu_sites].copy(), c["h_nm"].numpy()[:, :args.cpu_sites].copy()) for c in chroms]
        cpu = cpu_baseline(sl, vartheta, theta, S, args.cpu_sites)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1000.0 * dev_s / args.steps, "wall_ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"single_group whole-genome synthetic (~{args.total_sites // 1_000_000}M CpGs in 22 chromosomes), "
                                       f"{S} samples, {n_seeds} seeds per GPU ({n_seeds * world} seeds total), 250 particles, u=3",
                           "sites": total_T, "samples": S, "seeds": n_seeds * world, "chains_per_gpu": n_chains,
                           "k2_execution": (f"segmented: {n_units} units of <= {seg_sites} sites + {args.halo}-site halos, all resident CTAs busy"
                                            if seg_sites else "whole chains (sequential per chromosome x seed)"),
                           "l2": "inputs (3.6 GB counts + 1.3 GB emission table per GPU) exceed the 126 MB L2; no flush needed",
                           "parallelism": f"seeds sharded over {world} GPU(s), no data-path collective"},
                "gpu_launches": int(launches) * args.steps, "clocks": clk.summary(), "roofline": roofline, "roofline_emission": roofline_emission,
                "e2e": e2e, "mode_check": mode_check, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
