"""Multi-GPU search: one process per GPU, records sharded by a hash of the flank key.

Every rule of the search is local to one (left,right) key, so after ONE exchange step the GPUs are
independent:

  1. each rank ingests its own subset of the input files and runs K1 + one partition pass
     (``kb_shard_extract``): records grouped by destination shard, counts per shard;
  2. counts all-to-all, then the records all-to-all (``torch.distributed.all_to_all_single`` over
     NCCL / NVLink; ``gloo`` in the CPU tests) straight between library-owned device buffers;
  3. each rank sorts and groups its shard (``kb_shard_search``); survivor rows are gathered on rank 0
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


def exchange(searcher, send_ptr, send_counts, device, group=None):
    """Counts all-to-all + records all-to-all.  Returns the number of records received (they sit in the
    searcher's receive buffer, ordered by source rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    _dbg(f"exchange: send_counts={send_counts}")
    sc = torch.tensor(send_counts, dtype=torch.int64, device=device)
    rc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    n_recv = sum(recv_counts)
    _dbg(f"exchange: recv_counts={recv_counts}")
    recv_ptr = searcher.shard_recv_buffer(n_recv)
    send = searcher.wrap_records(send_ptr, sum(send_counts), device) if hasattr(searcher, "wrap_records") \
        else _wrap(send_ptr, sum(send_counts), device)
    recv = searcher.wrap_records(recv_ptr, n_recv, device) if hasattr(searcher, "wrap_records") \
        else _wrap(recv_ptr, n_recv, device)
    _dbg("exchange: records all_to_all")
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    _dbg("exchange: done")
    return n_recv, {"sent": int(sum(send_counts) - send_counts[dist.get_rank(group)]), "received": n_recv}


def sharded_search(searcher, device, have_outgroup=True, group=None):
    """Steps 1-3 on the sequences this rank has added to `searcher`.  Returns this rank's SearchResult."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    _dbg("shard_extract")
    send_ptr, counts = searcher.shard_extract(world)
    prof = list(searcher.last_profile()) if hasattr(searcher, "last_profile") else []
    ev = None
    if device is not None and getattr(device, "type", "cpu") == "cuda":
        import torch
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    n_recv, info = exchange(searcher, send_ptr, counts, device, group)
    if ev:
        ev[1].record()
    _dbg(f"shard_search n={n_recv}")
    res = searcher.shard_search(n_recv, have_outgroup=have_outgroup)
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


def shard_of_key(mixed_key, n_shards):
    """Destination shard of a mixed flank key — the host restatement of kb_digit()'s shard mode
    (csrc/kb_sort.cuh): the low 16 bits scaled to [0, n_shards)."""
    low = np.asarray(mixed_key, dtype=np.uint64) & np.uint64(0xFFFF)
    return ((low << np.uint64(16)) * np.uint64(n_shards)) >> np.uint64(32)
