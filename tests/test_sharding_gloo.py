"""N > 1 host path on CPU: world_size-2 gloo run of the chain partitioning + the cross-shard evidence reduction."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from hygeia_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lpt_balances():
    lens = [230, 220, 165, 150, 150, 145, 150, 125, 120, 135, 130, 125, 75, 85, 85, 105, 115, 70, 105, 75, 40, 60]
    for world in (1, 2, 4, 8):
        bins = sharding.lpt_assign(lens, world)
        assert sorted(i for b in bins for i in b) == list(range(len(lens)))
        loads = [sum(lens[i] for i in b) for b in bins]
        assert max(loads) <= sum(lens) / world + max(lens)
    assert sharding.lpt_assign(lens, 2) == sharding.lpt_assign(lens, 2)      # deterministic


def test_chains_for_rank_partitions():
    lens = [50, 40, 30]
    for by in ("seed", "chain"):
        seen = []
        for r in range(4):
            seen += sharding.chains_for_rank(lens, 8, r, 4, by=by)
        assert sorted(seen) == sorted((c, s) for c in range(3) for s in range(8))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from _oracle import Oracle
    from hygeia_b200 import model, philox, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [260, 180, 120]
    vartheta, _ = model.get_known_parameters()
    theta = model.default_theta()
    mine = sharding.chains_for_rank(lens, 4, rank, world, by="chain")
    evid = torch.zeros(len(lens), dtype=torch.float64)
    o = Oracle()
    for c, s in mine:
        ch = synthetic.make_chain(lens[c], 2, seed=100 + c)
        r = o.run(vartheta, theta, philox.uniforms_by_site(s, c, lens[c]), ch["n_total"], ch["n_meth"], smoothing=False)
        evid[c] += r["logz"][-1]
    dist.all_reduce(evid)           # the only cross-shard exchange of the path: a sum of sufficient statistics
    if rank == 0:
        q.put(evid.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    two = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q1 = ctx.Queue()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    p = ctx.Process(target=_worker, args=(0, 1, port, q1))
    p.start()
    one = q1.get(timeout=300)
    p.join(timeout=60)
    assert np.allclose(one, two, rtol=1e-12)
    assert np.all(one < 0)


def _posterior(c, sd, T, R=6):
    rng = np.random.default_rng(1000 * c + sd)
    p = rng.random((T, 1 + R))
    p[:, 0] = np.arange(T)
    return p, np.cumsum(-rng.random(T))


def _exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens, n_seeds, R = [50, 40, 30, 20], 4, 6
    mine = sharding.chains_for_rank(lens, n_seeds, rank, world, by="chain")
    views = []
    for c, sd in mine:
        p, z = _posterior(c, sd, lens[c])
        views.append((c, sd, torch.from_numpy(p), torch.from_numpy(z)))
    slots = -(-len(lens) * n_seeds // world) + 2                       # every rank uses the same number of slots
    psum = {c: torch.zeros((lens[c], R), dtype=torch.float64) for c in range(len(lens))}
    evid_mine = torch.zeros(slots, dtype=torch.float64)
    evid_all = torch.zeros(world * slots, dtype=torch.float64)
    nbytes = sharding.exchange_results(dist, views, psum, evid_mine, evid_all, range(len(lens)))
    if rank == 0:
        q.put((dict((c, psum[c].numpy().copy()) for c in psum), evid_all.numpy().copy(), nbytes))
    dist.destroy_process_group()


def test_result_exchange_on_two_ranks():
    """bench.py's multi-GPU exchange (sharding.exchange_results) on gloo: rank 0 ends with the posteriors summed over ALL seeds of
    every chromosome and every rank with the log-evidence of every chain of the job."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    psum, evid_all, nbytes = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lens, n_seeds = [50, 40, 30, 20], 4
    for c, T in enumerate(lens):
        want = sum(_posterior(c, sd, T)[0][:, 1:] for sd in range(n_seeds))
        assert np.allclose(psum[c], want, rtol=1e-13)
    want_evid = sorted(_posterior(c, sd, T)[1][-1] for c, T in enumerate(lens) for sd in range(n_seeds))
    got_evid = sorted(v for v in evid_all if v != 0.0)
    assert np.allclose(got_evid, want_evid, rtol=1e-15) and len(got_evid) == 16
    assert nbytes == evid_all.size * 8 + sum(T * 6 * 8 for T in lens)
