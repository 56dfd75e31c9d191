"""Chain partitioning across the GPUs of one box.

Chains (chromosome x seed for single-group, chromosome x batch x seed for two-group) never exchange data
(the reference fans them out as separate Nextflow processes, main.nf:21-30,47-75), so multi-GPU
is a partition problem: longest-processing-time-first bin packing, because a chain is strictly sequential and the
longest chromosome bounds the makespan.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence, Tuple


def lpt_assign(lengths: Sequence[int], n_bins: int) -> List[List[int]]:
    """Assign items (by index) to n_bins, longest first onto the least-loaded bin.  Deterministic."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    heap: List[Tuple[int, int]] = [(0, b) for b in range(n_bins)]
    heapq.heapify(heap)
    bins: List[List[int]] = [[] for _ in range(n_bins)]
    for i in order:
        load, b = heapq.heappop(heap)
        bins[b].append(i)
        heapq.heappush(heap, (load + lengths[i], b))
    return bins


def chains_for_rank(chrom_lengths: Sequence[int], n_seeds: int, rank: int, world: int, by: str = "seed"):
    """(chromosome, seed) pairs owned by `rank`.

    by="seed": seeds are dealt round-robin to ranks and every rank holds every chromosome (weak scaling, the counts of a
    chromosome are replicated -- BASELINE configs[2] at 16 seeds / 8 GPUs = 2 seeds per GPU);
    by="chain": LPT over all chromosome x seed chains (strong scaling; a chromosome's emission table is recomputed on every
    rank that owns one of its seeds)."""
    if by == "seed":
        return [(c, s) for s in range(n_seeds) if s % world == rank for c in range(len(chrom_lengths))]
    chains = [(c, s) for c in range(len(chrom_lengths)) for s in range(n_seeds)]
    bins = lpt_assign([chrom_lengths[c] for c, _ in chains], world)
    return [chains[i] for i in bins[rank]]
