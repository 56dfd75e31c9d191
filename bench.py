#!/usr/bin/env python
"""bench.py -- Hygeia's single-group inference hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                  # this framework (CUDA): BASELINE configs[1]
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference's own CPU code (oracle/_ref), bounded sample
    torchrun ... bench.py --gpus N ...                              # one rank per GPU (configs[2] at N = 8)
    ... --scaling strong        fixed 16 seeds x 22 chromosomes = 352 chains, LPT-assigned over the ranks
    ... --config c1 | c5        BASELINE configs[0] (T = 600 k, S = 4, one chain pair) / configs[4] (S = 1000, 4 seeds, counts
                                sharded by chromosome)
    ... --config c4             BASELINE configs[3]: two-group (case/control) filter + backward simulation, 50 + 50 samples, the
                                genome cut into the reference's 100 000-site windows with 5 000-site halos, 1 seed per GPU

A "step" is one pass of the hot path over the synthetic genome: K1 (emission table of every chromosome the rank holds) + K2 (the
recursion for every chromosome x seed chain of the rank) + the multi-GPU exchange of the results (below).  Default workload at
N = 1 = BASELINE configs[1]: ~28 M CpG sites in 22 synthetic chromosomes, 32 samples sharing one regime path, 2 seeds.

Multi-GPU (chains never exchange data; what crosses NVLink are RESULTS).
  weak   (default): rank r runs seeds {2r, 2r+1} of the same genome (N = 8 is configs[2]: 16 seeds).
  strong          : 352 chains assigned longest-first (hygeia_b200.sharding.chains_for_rank(by="chain")); every rank computes the
                    emission tables of the chromosomes it owns a chain of.
  In both, inside the timed region: an NCCL all-gather of every chain's final log-evidence log Z_T, and an NCCL reduce (sum) to
  rank 0 of the posterior rows summed over the rank's seeds, per chromosome (T x 6 fp64) -- the seed-averaged regime
  posteriors, i.e. what the reference obtains by concatenating the per-seed files (src/two_group/aggregate_results.py:125-147).
  Rank 0 checks the gathered evidences against a recomputation (weak: the sum over seeds of its own chains must equal its share).

Prints ONE JSON line (rank 0).
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
DATA_SEED = 20261018


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--config", default="c2", choices=["c1", "c2", "c4", "c5"])
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    p.add_argument("--total-sites", type=int, default=28_000_000)
    p.add_argument("--samples", type=int, default=None)
    p.add_argument("--seeds-per-gpu", type=int, default=None, help="weak scaling: seeds per rank (default 2; c5: 4 in total)")
    p.add_argument("--strong-seeds", type=int, default=16)
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--cpu-sites", type=int, default=2500, help="sites per chain of the bounded CPU-baseline sample")
    p.add_argument("--ref-sites", type=int, default=5000, help="sites per chain and step of --impl reference")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--emission-only", action="store_true", help="time K1 alone (kernel tuning aid; not a bench line)")
    p.add_argument("--segment-sites", type=int, default=-1,
                   help="K2 execution: -1 = segmented, size chosen by the library (default); 0 = whole chains (the reference's "
                        "sequential run); N = segments of <= N sites")
    p.add_argument("--halo", type=int, default=5000, help="left/right halo of a segment (sites)")
    p.add_argument("--no-e2e-pipeline", dest="e2e_pipeline", action="store_false")
    p.add_argument("--no-mode-check", action="store_true", help="skip the comparison of segmented vs whole-chain results")
    p.add_argument("--no-whole-chain", action="store_true", help="skip the whole-chain (reference semantics) timing leg")
    p.add_argument("--staged-outputs", action="store_true", help="device-resident leg with the posteriors staged in HBM (as at N > 1) instead of streamed to pinned host memory")
    a = p.parse_args()
    if a.samples is None:
        a.samples = {"c1": 4, "c2": 32, "c4": 50, "c5": 1000}[a.config]
    if a.seeds_per_gpu is None:
        a.seeds_per_gpu = 1 if a.config == "c4" else 2
    if a.config == "c1":
        a.total_sites = 600_000 if a.total_sites == 28_000_000 else a.total_sites
    return a


def chromosome_lengths(args):
    from hygeia_b200 import synthetic
    if args.config == "c1":
        return [args.total_sites]
    return synthetic.chromosome_lengths(args.total_sites)


# ------------------------------------------------------------------------------------------------------------------
# synthetic genome: regime paths on the host, counts on the device; chromosome c is a pure function
# of (DATA_SEED, c), so every rank that needs it generates the same data
# ------------------------------------------------------------------------------------------------------------------
def make_chromosome(c, T, S, device, regimes=None, salt=0):
    import torch
    from hygeia_b200 import model, synthetic
    alpha, beta = model.beta_parameters(model.DEFAULT_MU, model.DEFAULT_SIGMA)
    g = torch.Generator(device=device)
    g.manual_seed(DATA_SEED * 131 + c + 100003 * salt)
    al = torch.tensor(alpha, device=device, dtype=torch.float32)
    be = torch.tensor(beta, device=device, dtype=torch.float32)
    rng = np.random.default_rng(DATA_SEED + c)
    if regimes is None:
        regimes = synthetic.simulate_regimes(T, rng)
    pitch = (T + 7) // 8 * 8
    r = torch.from_numpy(regimes.astype(np.int64)).to(device)
    nt = torch.zeros((S, pitch), dtype=torch.uint16, device=device)
    nm = torch.zeros((S, pitch), dtype=torch.uint16, device=device)
    blk = max(1, min(S, (1 << 28) // max(T, 1)))   # samples per generation block: bounds the fp32 temporaries
    for s0 in range(0, S, blk):
        s1 = min(S, s0 + blk)
        n = torch.poisson(torch.full((s1 - s0, T), 30.0, device=device), generator=g)
        n = n * (torch.rand((s1 - s0, T), device=device, generator=g) >= 0.05)
        ga = torch._standard_gamma(al[r].expand(s1 - s0, T).contiguous(), generator=g)
        gb = torch._standard_gamma(be[r].expand(s1 - s0, T).contiguous(), generator=g)
        x = torch.binomial(n, (ga / (ga + gb)).clamp(0, 1), generator=g)
        nt[s0:s1, :T] = n.to(torch.int32).to(torch.uint16)
        nm[s0:s1, :T] = x.to(torch.int32).to(torch.uint16)
        del n, ga, gb, x
    pos = torch.from_numpy(synthetic.simulate_positions(T, rng).astype(np.int64)).to(torch.int32)
    return dict(index=c, T=T, pitch=pitch, n_total=nt, n_meth=nm, positions=pos, regimes=regimes)


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
            "sample": f"EXTRAPOLATED from the first {sites} sites of {n} chromosomes, S={S}, 1 seed each, one chain per core "
                      f"(oracle/_ref = reference headers, -O3 -ffast-math); per-site cost is constant along a chain; "
                      f"slowest chain {max(per):.1f} s, pool wall {wall:.1f} s",
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
    lens = chromosome_lengths(args)
    n = min(cores, max(len(lens), 2 if args.config == "c1" else 1))
    sites = min(args.ref_sites, min(lens))
    if S >= 1000:
        sites = max(100, sites // 25)   # the reference's cost grows with S (0.27 ms + 0.28 ms x S per site)
    chains = []
    for i in range(n):
        rng = np.random.default_rng(DATA_SEED + i)
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
            "ms_per_step": 1000.0 * float(np.mean(secs)), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, 1) + f" -- EXTRAPOLATED from a bounded sample: first {sites} sites of {n} chromosomes, "
                                                            "1 seed, one chain per core (per-site cost is constant along a chain)",
                       "samples": S},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(args, world):
    if args.config == "c1":
        return f"single_group synthetic chr22-sized CpG set ({args.total_sites} sites), {args.samples} samples, 2 seeds (BASELINE configs[0])"
    if args.config == "c5":
        return (f"emission-heavy cohort: whole-genome synthetic (~{args.total_sites // 1_000_000}M CpGs in 22 chromosomes), {args.samples} samples, "
                f"4 seeds, counts sharded by chromosome over {world} GPU(s) (BASELINE configs[4])")
    if args.scaling == "strong":
        return (f"single_group whole-genome synthetic (~{args.total_sites // 1_000_000}M CpGs in 22 chromosomes), {args.samples} samples, "
                f"{args.strong_seeds} seeds = {22 * args.strong_seeds} chains LPT-sharded over {world} GPU(s) (BASELINE configs[2], strong scaling)")
    return (f"single_group whole-genome synthetic (~{args.total_sites // 1_000_000}M CpGs in 22 chromosomes), {args.samples} samples, "
            f"{args.seeds_per_gpu} seeds per GPU ({args.seeds_per_gpu * world} seeds total), 250 particles, u=3 "
            f"(BASELINE configs[{1 if world == 1 else 2}])")



# ------------------------------------------------------------------------------------------------------------------
# --config c4: the two-group path (K1 for both groups + K4/K5), BASELINE configs[3]
# ------------------------------------------------------------------------------------------------------------------
TG_SEGMENT, TG_BUFFER = 100_000, 5_000     # run_inference_two_groups.py:64-72


def tg_windows(T):
    """(lo, hi, own_lo, own_hi) of the reference's batches of a chromosome (run_inference_two_groups.py:194-219)."""
    out = []
    for b in range(1 + T // TG_SEGMENT):
        if b * TG_SEGMENT >= T:
            break
        lo, hi = max(0, b * TG_SEGMENT - TG_BUFFER), min((b + 1) * TG_SEGMENT + TG_BUFFER, T)
        out.append((lo, hi, b * TG_SEGMENT, min((b + 1) * TG_SEGMENT, T)))
    return out


def _tg_cpu_worker(job):
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from _oracle import Oracle
    from _tg_case import make_case
    import tg_oracle
    sites, S, seed = job
    c = make_case(sites, S, seed=seed)
    o = Oracle()
    lo_c = o.emission(c["alpha"], c["beta"], c["nt_c"], c["nm_c"]); lo_k = o.emission(c["alpha"], c["beta"], c["nt_k"], c["nm_k"])
    t0 = time.perf_counter()
    tg_oracle.run(c["model"], lo_c, lo_k, M=50, n_backward=25, seed=seed, chain=0)
    return time.perf_counter() - t0


def tg_cpu_baseline(S, sites, n_procs):
    """oracle/tg_oracle.py (NumPy restatement of the reference's TensorFlow path, pinned to it: DESIGN.md section 2), one chain per core."""
    import multiprocessing as mp
    n = max(1, min(os.cpu_count() or 1, n_procs))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(n) as pool:
        per = pool.map(_tg_cpu_worker, [(sites, S, 11 + i) for i in range(n)])
    wall = time.perf_counter() - t0
    return {"value": n * sites * 2 * S / max(per), "unit": UNIT, "cores": n, "kind": "port",
            "sample": f"EXTRAPOLATED from {n} chains of {sites} sites, {S} + {S} samples, 50 ancestors x 48 proposals, 25 backward trajectories "
                      f"(oracle/tg_oracle.py, NumPy fp64, one chain per core); slowest chain {max(per):.1f} s, pool wall {wall:.1f} s"}


def tg_workload(args, world):
    return (f"two_group whole-genome synthetic (~{args.total_sites // 1_000_000}M CpGs in 22 chromosomes), {args.samples} + {args.samples} samples, "
            f"u=3, 50 ancestors x 48 proposals, 25 backward trajectories, {TG_SEGMENT}-site windows with {TG_BUFFER}-site halos, "
            f"{args.seeds_per_gpu} seed(s) per GPU ({args.seeds_per_gpu * world} in total) (BASELINE configs[3])")


def run_two_group_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    S = args.samples
    sites = 600
    vals, secs = [], []
    for _ in range(max(0, min(args.warmup, 1)) + args.steps):
        t0 = time.perf_counter()
        r = tg_cpu_baseline(S, sites, os.cpu_count() or 1)
        secs.append(time.perf_counter() - t0); vals.append(r["value"])
    vals, secs = vals[-args.steps:], secs[-args.steps:]
    v = float(np.mean(vals))
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1000.0 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": {"workload": tg_workload(args, 1) + " -- " + r["sample"], "samples": S},
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_two_group(args):
    import torch
    import torch.distributed as dist
    from hygeia_b200 import model
    from hygeia_b200.two_group import TwoGroupSession, control_group_parameters
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=dev)
    S, R, B, M = args.samples, 6, 25, 50
    lens = chromosome_lengths(args)
    seeds = [rank * args.seeds_per_gpu + k for k in range(args.seeds_per_gpu)]
    # control and case share the regime path except on stretches of 300 sites every 20 000 (the case group moves two regimes on)
    chroms = {}
    for c, T in enumerate(lens):
        ctl = make_chromosome(c, T, S, dev)
        reg_k = ctl["regimes"].copy()
        for s0 in range(7000, T - 300, 20000):
            reg_k[s0:s0 + 300] = (reg_k[s0:s0 + 300] + 2) % R
        cse = make_chromosome(c, T, S, dev, regimes=reg_k, salt=1)
        chroms[c] = (ctl, cse)
    torch.cuda.synchronize()
    windows = [(c, w) for c, T in enumerate(lens) for w in tg_windows(T)]
    owned = sum(w[3] - w[2] for _, w in windows)
    stepped = sum(w[1] - w[0] for _, w in windows)
    units_all = owned * 2 * S * args.seeds_per_gpu * world
    theta = model.default_theta()
    logp, omega_control = control_group_parameters(theta, R)

    the_session = TwoGroupSession(local)

    def build(device_resident, host=None):
        s = the_session          # one context: its caching allocator keeps the 40 GB of per-sweep buffers between sweeps
        s.clear()
        s.set_emission_model(model.DEFAULT_MU, model.DEFAULT_SIGMA, 3)
        specs = []
        for wi, (c, (lo, hi, _, _)) in enumerate(windows):
            ctl, cse = chroms[c]
            ids = []
            for g_i, ch in enumerate((ctl, cse)):
                if device_resident:
                    ids.append(s.add_dataset_ptr(hi - lo, S, ch["n_total"].data_ptr() + 2 * lo, ch["n_meth"].data_ptr() + 2 * lo, True, ch["pitch"]))
                else:
                    h_nt, h_nm = host[(c, g_i)]
                    ids.append(s.add_dataset_ptr(hi - lo, S, h_nt.data_ptr() + 2 * lo, h_nm.data_ptr() + 2 * lo, False, ch["pitch"]))
            for sd in seeds:
                specs.append(dict(control_dataset=ids[0], case_dataset=ids[1], T=hi - lo, seed=sd, chain_id=wi))
        s.set_two_group_model(logp, omega_control, np.full(R, 0.8), 3, M, B, t_max=TG_SEGMENT + 2 * TG_BUFFER)
        return s, specs

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    sess, specs = build(True)
    # trajectories come back to the host in both legs (hyg_tg_run is synchronous): pinned host arrays, allocated once
    pinned = [torch.empty((sp["T"], B, 5), dtype=torch.int32).pin_memory() for sp in specs]
    bufs = [t.numpy() for t in pinned]
    for sp, b in zip(specs, bufs):
        sp["trajectories"] = b
    evid = torch.zeros(world * len(specs), dtype=torch.float64, device=dev)

    def step():
        sess.emission()
        out = sess.run(specs)
        if multi:   # the log-evidence of every window of every seed, gathered on all ranks (model comparison across seeds)
            mine = torch.tensor([o["log_normalizing_constant"] for o in out], dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(evid, mine)
            torch.cuda.synchronize()
        tm = sess.timings()
        return tm["ms_emission"], sess.ms_two_group, out

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    em, tg = [], []
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            a, b, out = step()
            em.append(a); tg.append(b)
        barrier()
        wall = time.perf_counter() - t0
    t_all = torch.tensor([wall, sum(em) + sum(tg)], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    wall, dev_ms = float(t_all[0]), float(t_all[1])
    # device-resident value: the two kernels' event time (hyg_tg_run also copies 15 GB of trajectories back; that copy is in e2e)
    value = units_all * args.steps / (dev_ms / 1000.0)
    split = float(np.mean([(o["trajectories"][:, :, 0] == 0).mean() for o in out[:8]]))
    ms_tg = float(np.mean(tg))
    site_chains = stepped * len(seeds)
    alg = site_chains * (2 * 8 * R + B * 5 * 4 + 2 * (24 + M * 32))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6538.9)) if isinstance(peaks, dict) else 6538.9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": tg_workload(args, world), "samples": 2 * S, "windows_per_seed": len(windows), "inputs": "22 GB of counts, larger than L2: no flush between steps"},
            "gpu_launches": 2 * args.steps, "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": "tg_kernel (K4/K5), one persistent launch per step over all windows x seeds of the rank", "achieved": alg / (ms_tg / 1000.0) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (ms_tg / 1000.0) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms_tg, "share_of_step": ms_tg / (ms_tg + float(np.mean(em))),
                         "latency_bound": {"site_chains_per_s": site_chains / (ms_tg / 1000.0), "us_per_site_per_chain_in_flight": 1000.0 * ms_tg * 296 / site_chains, "us_per_site_per_sm": 1000.0 * ms_tg * 148 / site_chains,
                                           "halo_overhead": stepped / owned - 1.0, "mean_split_fraction_first_windows": split,
                                           "note": "a sequential recursion per window: the bound is per-site latency x 296 chains in flight (two resident CTAs per SM) (DESIGN.md section 4, K4/K5)"}},
            "roofline_emission": {"bound": "hbm", "kernel": "sg_emission_kernel<6> (K1), both groups", "ms_per_launch": float(np.mean(em)),
                                  "achieved": stepped * (2 * 4 * S + 2 * 8 * R) / (float(np.mean(em)) / 1000.0) / 1e9, "peak": peak, "unit": "GB/s"},
            "wall_ms_per_step_incl_trajectory_download": 1000.0 * wall / args.steps}
    line["roofline_emission"]["frac"] = line["roofline_emission"]["achieved"] / peak
    # ---- end to end: counts from pinned host memory, trajectories back in host memory ----
    line["e2e"] = None
    if not args.no_e2e:
        host = {}
        for c, (ctl, cse) in chroms.items():
            for g_i, ch in enumerate((ctl, cse)):
                host[(c, g_i)] = (ch["n_total"].cpu().pin_memory(), ch["n_meth"].cpu().pin_memory())
        h2d = sum(2 * S * (w[1] - w[0]) * 2 for _, w in windows)
        d2h = sum(b.nbytes for b in bufs)
        secs = []
        for _ in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            s2, sp2 = build(False, host)
            for sp, b in zip(sp2, bufs):
                sp["trajectories"] = b
            s2.emission()
            o2 = s2.run(sp2)
            barrier()
            secs.append(time.perf_counter() - t0)
        e = float(np.mean(secs[1:]))
        same = bool(all(a["log_normalizing_constant"] == b["log_normalizing_constant"] for a, b in zip(out, o2)))
        line["e2e"] = {"value": units_all / e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1000.0 * e,
                       "log_evidences_equal_to_device_resident_leg": same,
                       "api": "hygeia_b200.two_group.TwoGroupSession: clear -> set_emission_model -> add_dataset(pinned host windows) -> emission -> set_two_group_model -> run (trajectories into pinned host arrays)",
                       "log_evidence_first_window": o2[0]["log_normalizing_constant"]}
    line["cpu_baseline"] = None
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = tg_cpu_baseline(S, 400, os.cpu_count() or 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()

class _DevArr:
    """Raw device memory -> torch tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.config == "c4":
        return run_two_group_reference(args) if args.impl == "reference" else run_two_group(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from hygeia_b200 import model, sharding
    from hygeia_b200.single_group import Session, make_run_args

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S = args.samples
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    R = 6
    lens = chromosome_lengths(args)
    n_chrom = len(lens)

    # ---- which (chromosome, seed) chains does this rank own? ----
    if args.config == "c5":
        total_seeds = 4
        bins = sharding.lpt_assign(lens, world)                      # counts sharded BY CHROMOSOME
        my_chains = [(c, sd) for c in sorted(bins[rank]) for sd in range(total_seeds)]
        scaling = "strong"
    elif args.scaling == "strong":
        total_seeds = args.strong_seeds
        my_chains = sharding.chains_for_rank(lens, total_seeds, rank, world, by="chain")
        scaling = "strong"
    else:
        total_seeds = args.seeds_per_gpu * world
        my_chains = [(c, rank * args.seeds_per_gpu + k) for c in range(n_chrom) for k in range(args.seeds_per_gpu)]
        scaling = "weak"
    my_chroms = sorted({c for c, _ in my_chains})
    chroms = {c: make_chromosome(c, lens[c], S, dev) for c in my_chroms}
    torch.cuda.synchronize()
    total_T = sum(lens)
    units_per_step_all = total_T * S * total_seeds                  # whole job, all ranks
    my_site_chains = sum(lens[c] for c, _ in my_chains)

    # pinned host mirrors (inputs for the end-to-end leg, outputs for both legs)
    big = S >= 256   # c5: no pinned mirror of the counts (112 GB); the e2e leg is skipped
    for c, ch in chroms.items():
        if not big:
            ch["h_nt"] = torch.empty((S, ch["T"]), dtype=torch.uint16).pin_memory()
            ch["h_nm"] = torch.empty((S, ch["T"]), dtype=torch.uint16).pin_memory()
            ch["h_nt"].copy_(ch["n_total"][:, :ch["T"]]); ch["h_nm"].copy_(ch["n_meth"][:, :ch["T"]])
        ch["h_pos"] = ch["positions"].pin_memory()
    outs = {}
    for c, sd in my_chains:
        T = lens[c]
        outs[(c, sd)] = dict(probs=torch.empty((T, 1 + R), dtype=torch.float64).pin_memory(), logz=torch.empty(T, dtype=torch.float64).pin_memory())
    torch.cuda.synchronize()

    sess = Session(local)
    run_args = make_run_args()
    seg_request = Session.SEGMENT_AUTO if args.segment_sites < 0 else args.segment_sites
    if args.config == "c1" and args.segment_sites < 0:
        seg_request = Session.SEGMENT_AUTO
    sess.set_segmentation(seg_request, args.halo, args.halo)
    multi = world > 1

    def stage(device_resident, ses=None, which=None, device_outputs=False):
        """Stage the chains `which` (default all of the rank's) on session `ses`; returns the staged (c, seed) list."""
        ses = ses or sess
        which = my_chains if which is None else which
        ses.clear()
        ses.set_zero_copy_outputs(not device_outputs)
        ses.set_vartheta(vartheta)
        ses.set_theta(theta, max(lens))
        ds_of = {}
        specs = []
        for c, sd in which:
            ch = chroms[c]
            if c not in ds_of:
                if device_resident:
                    ds_of[c] = ses.add_dataset_ptr(ch["T"], S, ch["n_total"].data_ptr(), ch["n_meth"].data_ptr(), True, ch["pitch"])
                else:
                    ds_of[c] = ses.add_dataset_ptr(ch["T"], S, ch["h_nt"].data_ptr(), ch["h_nm"].data_ptr(), False, ch["T"])
            o = outs[(c, sd)]
            specs.append(dict(dataset=ds_of[c], seed=sd, chain_id=c, positions=ch["h_pos"].data_ptr(),
                              regime_probs=o["probs"].data_ptr(), logz=o["logz"].data_ptr()))
        ses.set_chains(specs)
        return list(which)

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs already in HBM ----
    staged = stage(True, device_outputs=multi or args.staged_outputs)
    n_chains = len(staged)
    # multi-GPU exchange: device views of the staged outputs (valid until the next set_chains)
    evid_all = torch.zeros(world * max(1, n_chains), dtype=torch.float64, device=dev) if multi else None
    exch = None
    if multi:
        views = []
        for i, (c, sd) in enumerate(staged):
            pp, zp = sess.device_outputs(i)
            views.append((c, sd, torch.as_tensor(_DevArr(pp, (lens[c], 1 + R)), device=dev), torch.as_tensor(_DevArr(zp, (lens[c],)), device=dev)))
        # every rank must issue the same sequence of collectives: reduce per chromosome, in chromosome order, over ALL chromosomes
        psum = {c: torch.zeros((lens[c], R), dtype=torch.float64, device=dev) for c in range(n_chrom)} if args.config != "c5" else \
               {c: torch.zeros((lens[c], R), dtype=torch.float64, device=dev) for c in my_chroms}
        max_chains = torch.tensor([n_chains], device=dev)
        dist.all_reduce(max_chains, op=dist.ReduceOp.MAX)
        slots = int(max_chains.item())
        evid_mine = torch.zeros(slots, dtype=torch.float64, device=dev)
        evid_all = torch.zeros(world * slots, dtype=torch.float64, device=dev)
        exch = dict(views=views, psum=psum, slots=slots, evid_mine=evid_mine)

    coll_bytes = [0]

    def exchange():
        """NCCL: all-gather of log Z_T of every chain; reduce(sum) to rank 0 of the posteriors summed over this rank's seeds."""
        v = exch
        # c5: chromosomes are disjoint across ranks, their seed sums are complete where they are
        coll_bytes[0] = sharding.exchange_results(dist, v["views"], v["psum"], v["evid_mine"], evid_all, range(n_chrom) if args.config != "c5" else [])

    def step_device():
        sess.emission()
        sess.filter(run_args)
        sess.sync()
        if multi:
            exchange()
            torch.cuda.synchronize()   # the exchange belongs to this step: NCCL kernels left running would share the SMs with the next K2
        return sess.timings()

    if args.emission_only:
        ms = []
        for _ in range(3 + args.steps):
            sess.emission(); sess.sync(); ms.append(sess.timings()["ms_emission"])
        my_T = sum(lens[c] for c in my_chroms)
        alg = my_T * S * 4 + my_T * R * 8
        best = min(ms[3:]); avg = float(np.mean(ms[3:]))
        print(json.dumps({"emission_only_ms": ms, "GBps_avg": alg / avg / 1e6, "GBps_best": alg / best / 1e6,
                          "frac_of_6538.9": alg / avg / 1e6 / 6538.9, "S": S}))
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
    ovl = sess.overlap_max_abs()
    stepped = int(sum(x[3] for x in st))
    forced_halo = int(sum(x[2] for x in st)); forced_lag = int(sum(x[0] for x in st))
    exact_sorts = int(sum(x[4] for x in st)); tie_sites = int(sum(x[5] for x in st)); ovl_bad = int(sum(x[7] for x in st))
    dev_s = sum(ev_ms) / 1000.0
    t_all = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t_all[0]), float(t_all[1])
    # whole-job throughput: everything all ranks processed / the barrier-bracketed time of the K steps (max over ranks); at N = 1
    # that time is K1 + K2 device time plus launch gaps, at N > 1 it includes the NCCL exchange of the results
    value = units_per_step_all * args.steps / wall

    # the gathered evidences: rank 0 holds log Z_T of every chain of the job -- sanity: finite, and this rank's slice is its own
    gather_check = None
    if not multi:
        # one GPU: the same sum over this job's chains, from the host copies -- comparable with the gathered sum of a multi-GPU run of
        # the same job (--scaling strong: 352 chains whatever N is)
        lz = [float(outs[(c, sd)]["logz"][-1]) for c, sd in my_chains]
        gather_check = {"chains_gathered": len(lz), "all_finite": bool(np.isfinite(lz).all()), "collective_bytes_per_step": 0,
                        "sum_log_evidence": float(np.sum(lz))}
    if multi and rank == 0:
        ea = evid_all.cpu().numpy().reshape(world, -1)
        gather_check = {"chains_gathered": int((ea != 0).sum()), "all_finite": bool(np.isfinite(ea).all()),
                        "collective_bytes_per_step": int(coll_bytes[0]),
                        "sum_log_evidence": float(ea.sum())}

    # ---- roofline ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (burst copy figure)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    longest = max(lens[c] for c in my_chroms)
    my_T = sum(lens[c] for c in my_chroms)
    # K2, the dominant kernel at S <= 32: per site and chain it reads the R emission log-densities and writes the R posterior
    # probabilities and log Z_t.  It is a sequential recursion (one CTA per chain / segment, 250 particles), bound by per-site latency,
    # not by HBM -- the fraction below says how far from HBM it is, the us/site figure is the number to optimise.
    k2_alg = my_site_chains * (R * 8 + R * 8 + 8)
    k2_ms = float(np.mean(f_ms)); k1_ms = float(np.mean(em_ms))
    k2_ach = k2_alg / (k2_ms / 1000.0) / 1e9
    K2_DRAM_B_PER_SITE_CHAIN, K1_DRAM_B_PER_SITE_S32 = 41.5, 176.2
    k2 = {"bound": "hbm", "kernel": "sg_filter_kernel<6,0> (K2: particle filter + fixed-lag smoother), 1 persistent launch per step, "
                                    "one CTA per (chain, segment) unit",
          "share_of_step": k2_ms / (k2_ms + k1_ms),
          "achieved": k2_ach, "peak": peak, "unit": "GB/s", "frac": k2_ach / peak, "peak_source": peak_source,
          "algorithmic_bytes_per_launch": k2_alg, "ms_per_launch": k2_ms,
          "traffic": K2_DRAM_B_PER_SITE_CHAIN * my_site_chains,
          "traffic_source": "ncu dram__bytes_read+write of a 4M-site capture, per owned site-chain, scaled to this launch "
                            "(posterior rows leave over PCIe into pinned host memory and are not DRAM traffic)",
          "latency_bound": {"chains": n_chains, "units": n_units, "segment_sites": seg_sites, "halo_sites": args.halo if seg_sites else 0,
                            "resident_ctas": workers, "sms": 148, "longest_chain_sites": longest,
                            "sites_stepped_incl_halos": stepped, "halo_overhead": stepped / float(max(my_site_chains, 1)) - 1.0,
                            "us_per_site_per_cta": 1000.0 * k2_ms * min(workers, n_units) / max(stepped, 1),
                            "us_per_site_per_sm": 1000.0 * k2_ms * min(148, n_units) / max(stepped, 1),
                            "site_chains_per_s": my_site_chains / (k2_ms / 1000.0),
                            "sites_forced_at_segment_end": forced_halo, "sites_forced_lag_set_full": forced_lag,
                            "sites_sorted_on_full_words": exact_sorts, "sites_where_an_exact_tie_decided_a_fate": tie_sites,
                            "segment_overlap_max_abs_posterior_diff": float(max(ovl)) if ovl else 0.0,
                            "segment_overlap_rows_over_1e-6": ovl_bad,
                            "note": "a strictly sequential recursion per unit: the bound is per-site latency x units in flight, "
                                    "HBM is idle; see DESIGN.md section 4 (K2) and 6"}}
    alg_bytes = my_T * S * 2 * 2 + my_T * R * 8
    ach = alg_bytes / (k1_ms / 1000.0) / 1e9
    k1 = {"bound": "hbm", "kernel": "sg_emission_kernel<6> (K1), 1 persistent launch per step over all chromosomes of the rank",
          "share_of_step": k1_ms / (k2_ms + k1_ms),
          "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_source,
          "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": k1_ms,
          "traffic": (K1_DRAM_B_PER_SITE_S32 * my_T) if S == 32 else None,
          "traffic_source": "ncu dram__bytes_read+write of a 4M-site S=32 capture, per site, scaled to this launch" if S == 32 else None}
    # the dominant kernel goes under "roofline"; the other one beside it
    if k1_ms > k2_ms:
        roofline, other_key, other = k1, "roofline_filter", k2
    else:
        roofline, other_key, other = k2, "roofline_emission", k1

    # ---- whole-chain leg (the reference's own semantics: every chain one sequential run; no halo approximation) ----
    whole = None
    if seg_sites and not args.no_whole_chain and args.config != "c5":
        sess.set_segmentation(0, args.halo, args.halo)
        stage(True, device_outputs=False)
        sess.emission(); sess.filter(run_args); sess.sync()            # warm
        t0 = time.perf_counter()
        sess.emission(); sess.filter(run_args); sess.sync()
        w_wall = time.perf_counter() - t0
        tw = sess.timings()
        w_all = torch.tensor([w_wall], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(w_all, op=dist.ReduceOp.MAX)
        whole = {"value": units_per_step_all / float(w_all[0]), "unit": UNIT, "ms_per_step": 1000.0 * float(w_all[0]),
                 "ms_filter": tw["ms_filter"], "us_per_site_longest_chain": 1000.0 * tw["ms_filter"] / longest,
                 "note": "segment_sites = 0: each chromosome x seed chain is ONE sequential run, exactly OnlineCombinedInference::run; "
                         f"{n_chains} chains on {min(n_chains, 148)} of 148 SMs, bounded by the longest chromosome"}
        sess.set_segmentation(seg_request, args.halo, args.halo)

    # ---- end-to-end leg: host (pinned) buffers in, host buffers out, through the public API ----
    e2e = None
    if not args.no_e2e and not big:
        n_seeds_of = {}
        for c, sd in my_chains:
            n_seeds_of[c] = n_seeds_of.get(c, 0) + 1
        h2d = sum(2 * S * lens[c] * 2 for c in my_chroms) + sum(lens[c] * 4 for c, _ in my_chains)
        d2h = sum((lens[c] * (1 + R) * 8 + lens[c] * 8) for c, _ in my_chains)
        # Two contexts (two streams) take half of the chromosomes each, so that the host -> device copy of the second half and
        # the device -> host copy of the first half's log Z overlap the recursion of the other half.
        order = sorted(my_chroms, key=lambda c: -lens[c])
        halves = [order[0::2], order[1::2]] if (args.e2e_pipeline and len(order) > 1) else [order]
        sessions = [sess]
        if len(halves) > 1:
            s2 = Session(local)
            s2.set_segmentation(seg_request, args.halo, args.halo)
            sessions.append(s2)

        def step_e2e():
            for ses, hc in zip(sessions, halves):
                stage(False, ses, [(c, sd) for c, sd in my_chains if c in hc])
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
        if multi:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": units_per_step_all * args.e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1000.0 * e2e_s / args.e2e_steps, "steps": args.e2e_steps,
               "api": "hygeia_b200.single_group.Session: add_dataset(pinned host) -> set_chains -> emission -> filter -> download"
                      + (", two Sessions (contexts) with half of the chromosomes each so that copies overlap the other half's recursion; "
                         if len(halves) > 1 else "; ") +
                      "posterior rows are written by K2 straight into the pinned host buffers (counted in d2h_bytes_per_step), "
                      "log Z is staged in HBM and copied"}
        last = my_chains[-1]
        p = outs[last]["probs"].numpy()
        e2e["regime_call_accuracy_vs_simulated_truth_last_chain"] = float((p[:, 1:].argmax(1) == chroms[last[0]]["regimes"]).mean())
        e2e["non_finite_posterior_rows"] = int(sum(int((~np.isfinite(o["probs"].numpy())).any(1).sum()) for o in outs.values()))

    # ---- parity of the two execution modes: segmented (timed above) vs the sequential whole-chain run ----
    mode_check = None
    if e2e is not None and seg_sites and not args.no_mode_check:
        sub_c = sorted(my_chroms, key=lambda c: lens[c])[:3]
        sub = [(c, sd) for c, sd in my_chains if c in sub_c]
        seg_out = {k: (outs[k]["probs"].numpy().copy(), outs[k]["logz"].numpy().copy()) for k in sub}
        sess.set_segmentation(0, args.halo, args.halo)
        stage(False, sess, sub)
        sess.emission(); sess.filter(run_args); sess.download()
        dp = dz = 0.0; calls = 0; rows = 0
        for k in sub:
            sp, sz = seg_out[k]
            wp, wz = outs[k]["probs"].numpy(), outs[k]["logz"].numpy()
            dp = max(dp, float(np.abs(sp[:, 1:] - wp[:, 1:]).max()))
            dz = max(dz, float((np.abs(sz - wz) / np.abs(wz)).max()))
            calls += int((sp[:, 1:].argmax(1) != wp[:, 1:].argmax(1)).sum()); rows += wp.shape[0]
        mode_check = {"compared": f"segmented vs whole-chain execution on the {len(sub_c)} shortest chromosomes of the rank x its seeds ({rows} site rows)",
                      "max_abs_posterior_diff": dp, "max_rel_logz_diff": dz, "differing_regime_calls": calls, "tolerance": 1e-6}
        sess.set_segmentation(seg_request, args.halo, args.halo)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not big:
        cs = min(args.cpu_sites, min(lens))
        sl = [(chroms[c]["h_nt"].numpy()[:, :cs].copy(), chroms[c]["h_nm"].numpy()[:, :cs].copy()) for c in my_chroms]
        if args.config == "c1":
            sl = sl * 2
        cpu = cpu_baseline(sl, vartheta, theta, S, cs)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1000.0 * wall / args.steps, "device_ms_per_step": 1000.0 * dev_s / args.steps, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, world),
                           "sites": total_T, "samples": S, "seeds": total_seeds, "chains_on_rank0": n_chains,
                           "k2_execution": (f"segmented: {n_units} units of <= {seg_sites} sites + {args.halo}-site halos on rank 0, "
                                            f"left-halo overlap check on (max |dp| {float(max(ovl)) if ovl else 0.0:.2e})"
                                            if seg_sites else "whole chains (sequential per chromosome x seed)"),
                           "l2": "inputs (counts + emission table per GPU) exceed the 126 MB L2; no flush needed",
                           "parallelism": (f"chains sharded over {world} GPU(s) ({scaling}); NCCL all-gather of log Z_T + reduce of the "
                                           f"seed-summed posteriors inside the timed region" if multi else "1 GPU")},
                "site_chains_per_s": (total_T * total_seeds) * args.steps / wall,
                "whole_chain": whole, "whole_chain_value": whole["value"] if whole else None,
                "gpu_launches": int(launches) * args.steps, "clocks": clk.summary(), "roofline": roofline, other_key: other,
                "e2e": e2e, "mode_check": mode_check, "gather_check": gather_check, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
