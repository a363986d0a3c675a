"""Multi-GPU search: one process per GPU, records sharded by the top bits of the mixed flank key.

Every rule of the search is local to one (left,right) key, so after ONE exchange step the GPUs are
independent:

  0. all ranks agree on the partition plan (``kb_shard_plan``: same shard count, same total size);
  1. each rank ingests its own subset of the input files and runs K1 + partition level 0
     (``kb_shard_extract``): records grouped by level-0 digit, hence by owner shard (contiguous digit ranges);
  2. counts all-to-all, per-digit counts all-gather, then the records all-to-all
     (``torch.distributed.all_to_all_single`` over NCCL / NVLink; ``gloo`` in the CPU tests) straight between
     library-owned device buffers;
  3. each rank runs the remaining partition levels and the bucket hash on its shard (``kb_shard_search``); the
     (source rank, digit) pieces it received are the parents of level 1; survivor rows are gathered on rank 0
     (set semantics: no ordering step).

The reference has no counterpart (it is single-host multiprocessing, krisp_fasta.py:86-123); the file
fan-out mirrors ``sortedKmersParallel``: files are independent extraction units.
"""
import os
import sys
import time

import numpy as np

_DEBUG = bool(os.environ.get("KRISP_DEBUG"))


def _dbg(msg):
    if _DEBUG:
        print(f"[krisp_b200 rank {os.environ.get('RANK', '?')} {time.time():.3f}] {msg}", file=sys.stderr, flush=True)


def assign_files(n_files, world_size, sizes=None):
    """File index -> rank.  Greedy by size (largest first) when sizes are known, else round-robin."""
    if sizes is None:
        return [i % world_size for i in range(n_files)]
    load = [0] * world_size
    owner = [0] * n_files
    for i in sorted(range(n_files), key=lambda j: (-sizes[j], j)):
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += sizes[i]
    return owner


class _DeviceArray:
    """Raw device pointer -> something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def _wrap(ptr, n, device):
    import torch
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DeviceArray(ptr, n), device=device)


def first_digit(shard, n_shards, n_digits):
    """First level-0 digit owned by `shard` (the host restatement of shard_first_digit in csrc/kb_api.cu)."""
    return (shard * n_digits) // n_shards


def exchange(searcher, send_ptr, send_counts, digit_counts, device, group=None):
    """Counts all-to-all + per-digit counts all-gather + records all-to-all.  Returns (records received — they sit in
    the searcher's receive buffer, ordered by source rank and digit —, the piece counts [source][digit of my shard], info)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _dbg(f"exchange: send_counts={send_counts}")
    sc = torch.tensor(send_counts, dtype=torch.int64, device=device)
    rc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(rc, sc, group=group)
    dg = torch.tensor(digit_counts, dtype=torch.int64, device=device)
    alld = torch.empty(world * len(digit_counts), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(alld, dg, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    nd = len(digit_counts)
    lo, hi = first_digit(rank, world, nd), first_digit(rank + 1, world, nd)
    pieces = alld.view(world, nd)[:, lo:hi].reshape(-1).tolist()
    n_recv = sum(recv_counts)
    _dbg(f"exchange: recv_counts={recv_counts}")
    recv_ptr = searcher.shard_recv_buffer(n_recv)
    wrap = searcher.wrap_records if hasattr(searcher, "wrap_records") else _wrap
    send = wrap(send_ptr, sum(send_counts), device)
    recv = wrap(recv_ptr, n_recv, device)
    _dbg("exchange: records all_to_all")
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    _dbg("exchange: done")
    return n_recv, pieces, {"sent": int(sum(send_counts) - send_counts[rank]), "received": n_recv}


def sharded_search(searcher, device, have_outgroup=True, group=None, total_bases=None):
    """Steps 0-3 on the sequences this rank has added to `searcher`.  Returns this rank's SearchResult.
    `total_bases` = bases over all ranks (all-reduced from ``searcher.bases_added`` when not given)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if total_bases is None:
        t = torch.tensor([int(getattr(searcher, "bases_added", 0))], dtype=torch.int64, device=device)
        dist.all_reduce(t, group=group)
        total_bases = int(t.item())
    searcher.shard_plan(world, rank, total_bases)
    _dbg("shard_extract")
    send_ptr, counts, digits = searcher.shard_extract()
    prof = list(searcher.last_profile()) if hasattr(searcher, "last_profile") else []
    ev = None
    if device is not None and getattr(device, "type", "cpu") == "cuda":
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    n_recv, pieces, info = exchange(searcher, send_ptr, counts, digits, device, group)
    if ev:
        ev[1].record()
    _dbg(f"shard_search n={n_recv}")
    res = searcher.shard_search(n_recv, pieces, have_outgroup=have_outgroup)
    _dbg(f"shard_search done: {getattr(res, 'n_groups', '?')} groups")
    if ev:
        ev[1].synchronize()
        prof.append(("K4 all_to_all (NCCL)", ev[0].elapsed_time(ev[1])))
    if hasattr(res, "profile"):
        res.profile = prof + list(res.profile)
        res.exchange = info
    return res


def gather_rows(rows, group=None):
    """All ranks' rows on every rank, canonically sorted (row sets of different shards are disjoint)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, list(rows), group=group)
    out = []
    for p in parts:
        out.extend(p)
    return sorted(out)


def digit_of_key(mixed_key, flank_bits, bits0):
    """Level-0 digit of a mixed flank key — the host restatement of the device digit (csrc/kb_part.cuh): its top bits."""
    return np.asarray(mixed_key, dtype=np.uint64) >> np.uint64(flank_bits - bits0)
