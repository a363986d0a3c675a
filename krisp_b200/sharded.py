"""Multi-GPU search: one process per GPU, records sharded by the top bits of the mixed flank key.

Every rule of the search is local to one (left,right) key, so after ONE exchange step the GPUs are
independent.  Three exchanges, all between library-owned device buffers:

``slab_search`` (default; one-word records):
  1. all ranks derive the same plan (``kb_shard_slab_plan``); each rank ingests its own files and runs K1 fused with partition
     level 0 (``kb_shard_slab_extract``): every level-0 digit's records sit in a fixed-capacity slab laid out as in the owner's
     receive buffer — no count exchange before the data;
  2. the slab fill levels are all-gathered on the device; level 1 of the rank's OWN slabs starts at once (``kb_shard_slab_own``);
  3. digit groups travel as bulk peer copies (copy engines over NVLink, CUDA IPC mappings) behind each other on a copy stream, a
     tiny all-reduce per group on a vote stream says "landed everywhere";
  4. level 1 + bucket hash per group on the owner (``kb_shard_slab_level``) while the next group is in flight;
  5. ``kb_shard_slab_finish`` + an all-reduce of the status: re-plan (divergent genomes), slab overflow (-> exact exchange) and a grown
     survivor table are decided by all ranks alike.
``direct_search`` (multi-word records; slab overflow): K1 with the level-0 histogram, digit counts all-gathered, partition level 0 =
  peer stores straight into the owners' buffers (``kb_shard_scatter``), then ``kb_shard_search``.
``exchange_mode="nccl"``: local partition + ``all_to_all_single`` (also what the gloo CPU tests drive, with a test double).

Survivor rows of different ranks are disjoint (set semantics: concatenate and sort, ``gather_rows``).  The reference has no
counterpart (it is single-host multiprocessing, krisp_fasta.py:86-123); the file fan-out mirrors ``sortedKmersParallel``: files
are independent extraction units.
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


def piece_tables(table, rank):
    """From the all-gathered digit counts ``table[source][digit]``: (records every shard receives, for `rank` the element offset
    of its piece of every digit in the owner's receive buffer, the piece counts of `rank`'s own shard in arrival order).
    A receive buffer is laid out by source rank, then digit."""
    table = np.asarray(table, dtype=np.int64)
    world, nd = table.shape
    firsts = [first_digit(s, world, nd) for s in range(world + 1)]
    need = [int(table[:, firsts[s]:firsts[s + 1]].sum()) for s in range(world)]
    piece_base = [0] * nd
    for s in range(world):
        sub = table[:, firsts[s]:firsts[s + 1]]
        flat = np.concatenate([[0], np.cumsum(sub.reshape(-1))])      # pieces in (source, digit) order
        dps = firsts[s + 1] - firsts[s]
        for j in range(dps):
            piece_base[firsts[s] + j] = int(flat[rank * dps + j])
    pieces = table[:, firsts[rank]:firsts[rank + 1]].reshape(-1).tolist()
    return need, piece_base, pieces


def shard_child_counts(children, n_digits, rank):
    """From the all-gathered two-level histograms ``children[source][digit << bits1 | next digit]`` (every rank's K1 counts over the
    whole key space): the level-1 child counts of the digits `rank` owns, summed over the sources, in digit order — what
    kb_shard_set_child_counts expects (the owner then needs no histogram pass over the records it received)."""
    children = np.asarray(children, dtype=np.int64)
    world, nch = children.shape
    per = nch // n_digits                                             # children per level-0 digit
    lo, hi = first_digit(rank, world, n_digits), first_digit(rank + 1, world, n_digits)
    return children[:, lo * per:hi * per].sum(axis=0).astype(np.uint64)


def direct_search(searcher, device, have_outgroup=True, group=None):
    """Steps 1-3 with the exchange FUSED into partition level 0: every rank stores each digit's run straight into the
    owner's receive buffer over NVLink peer memory (CUDA IPC mappings of library-owned buffers), so the only collectives
    left are two small ones (digit counts all-gather, one barrier).  GPUs of one box only."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    digits = searcher.shard_count()                                   # K1 (+ level-0 histogram)
    prof = list(searcher.last_profile())
    nd = len(digits)
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev[0].record()
    # K1's two-level histogram, when it has one: the level-1 child counts travel with the digit counts, so that the owner does not
    # have to re-read the records it received just to count them
    child = searcher.shard_child_counts() if hasattr(searcher, "shard_child_counts") else None
    nch = 0 if child is None else int(child.size)
    payload = np.asarray(digits, dtype=np.int64) if child is None else np.concatenate([np.asarray(digits, dtype=np.int64), child.astype(np.int64)])
    dg = torch.from_numpy(payload).to(device)
    alld = torch.empty(world * (nd + nch), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(alld, dg, group=group)
    both = alld.view(world, nd + nch).cpu().numpy()
    table = both[:, :nd]                                              # [source][digit]
    firsts = [first_digit(s, world, nd) for s in range(world + 1)]
    need, piece_base, pieces = piece_tables(table, rank)
    if nch:
        searcher.shard_set_child_counts(shard_child_counts(both[:, nd:], nd, rank))
    _ensure_ipc(searcher, device, group, need)
    searcher.shard_scatter(piece_base)                                # partition level 0 -> peer stores
    prof += [p for p in searcher.last_profile() if p[0].startswith("K2 partition 0")]
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    dist.all_reduce(flag, group=group)                                # every rank's stores have landed
    ev[1].record()
    res = searcher.shard_search(need[rank], pieces, have_outgroup=have_outgroup)
    ev[1].synchronize()
    prof.append(("K4 count all-gather + partition/exchange + barrier", ev[0].elapsed_time(ev[1])))
    res.profile = prof + list(res.profile)
    res.exchange = {"sent": int(sum(digits) - sum(digits[firsts[rank]:firsts[rank + 1]])), "received": need[rank]}
    return res


def _ensure_ipc(searcher, device, group, need):
    """Every rank's receive buffer holds at least need[rank] records and is mapped by every peer (CUDA IPC).  `need` is the same
    list on every rank, so all ranks take the (re)allocation branch together; buffers only grow."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    state = searcher.__dict__.setdefault("_ipc_state", {"cap": [0] * world, "world": world})
    if state["world"] == world and all(n <= c for n, c in zip(need, state["cap"])):
        return
    state["cap"] = [max(c, n + n // 8 + 4096) for n, c in zip(need, state["cap"] if state["world"] == world else [0] * world)]
    state["world"] = world
    if hasattr(searcher, "shard_ipc_close"):
        searcher.shard_ipc_close()                                    # nobody maps a buffer that is about to be reallocated
        dist.barrier(group=group)
    mine = torch.frombuffer(bytearray(searcher.shard_ipc_export(state["cap"][rank])), dtype=torch.uint8).to(device)
    allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allh, mine, group=group)
    blob = allh.cpu().numpy().tobytes()
    searcher.shard_ipc_import([blob[64 * r:64 * (r + 1)] for r in range(world)])


def _slab_groups(max_groups):
    """Digit groups of the pipelined exchange: up to 4 where the library allows it (two-level plan, tile-aligned slabs: every rank
    derives the same answer from the same plan), else one.  KRISP_SLAB_GROUPS overrides.  Measured (0.2 Gbp per GPU, own slabs
    first): 2 GPUs 6.04 ms with 4 groups, 6.22 with 8, 6.64 with 2; 8 GPUs 7.71 / 7.88 / 8.09 — fewer, larger launches per group
    win once this rank's own slabs cover the wait for the first group."""
    env = os.environ.get("KRISP_SLAB_GROUPS")
    return max(1, min(int(env) if env else 4, int(max_groups)))


def slab_search(searcher, device, have_outgroup=True, group=None, total_bases=None):
    """The exchange fused into K1 (one-word records): every rank's K1 + partition level 0 (csrc/kb_extract_part.cuh) store each
    level-0 digit's run straight into a fixed-capacity slab of the owner's receive buffer over NVLink peer memory.  There is no count
    exchange before the data and no separate partition pass; the slab fill levels stay on the device and are all-gathered there —
    the all-gather doubles as the barrier after which every rank's stores have landed — and the owner runs level 1 + the bucket hash
    on its slabs.  No host round trip between the stages.
    Returns the SearchResult, or None when the slab exchange does not apply (multi-word records, slab overflow on a repetitive
    input): the caller then uses the exact exchange (direct_search).  All ranks return the same kind of answer."""
    import torch
    import torch.distributed as dist
    from ._lib import UnsupportedError
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = int(getattr(searcher, "bases_added", 0))
    extra = int(searcher.__dict__.get("_shard_bb_extra", 0))
    # The plan (three small collectives + host reads) is reused while nothing it depends on changed.  Only when the caller passes
    # total_bases — on every rank alike — is "nothing changed" a collective fact; otherwise every search plans again.
    key = (world, rank, total_bases, mine, tuple(getattr(searcher, "lo", ()) or ()), getattr(searcher, "n_files", None), extra)
    cached = searcher.__dict__.get("_slab_plan") if total_bases is not None else None
    for _ in range(4):
        if cached is not None and cached[0] == key:
            nd, cap, max_groups = cached[1]
        else:
            t = torch.tensor([mine], dtype=torch.int64, device=device)
            if total_bases is None:
                s = torch.tensor([mine], dtype=torch.int64, device=device)
                dist.all_reduce(s, group=group)
                total = int(s.item())
            else:
                total = int(total_bases)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            max_rank_bases = int(t[0].item())
            searcher.set_option("shard_bb_extra", extra)
            try:
                nd, cap, max_groups = searcher.shard_slab_plan(world, rank, total, max_rank_bases)
                ok = 1
            except UnsupportedError:
                nd, cap, max_groups, ok = 0, 0, 1, 0
            caps = torch.tensor([cap if ok else -1], dtype=torch.int64, device=device)
            allc = torch.empty(world, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(allc, caps, group=group)
            need = [int(c) for c in allc.tolist()]
            if min(need) < 0:
                return None                                           # (some rank cannot: nobody does)
            _ensure_ipc(searcher, device, group, need)
            searcher.__dict__["_slab_plan"] = (key, (nd, cap, max_groups))
        cached = None
        n_groups = _slab_groups(max_groups)
        main = torch.cuda.current_stream(device)
        # streams: `side` carries nothing but the bulk copies, back to back; `vote` carries the tiny collectives ("every rank's copies
        # of group g have landed"), so that a collective waiting for a free SM never holds up the next group's copies
        # copy streams (= copy engines working at once; KRISP_COPY_STREAMS).  Two GPUs: the one peer's copy is cut into byte ranges;
        # more: every stream serves some of the peers.  The streams stay in step group by group (below), so that group g is complete
        # before anybody spends bandwidth on group g + 1.  One stream is the default: on 2 GPUs 1 / 2 / 3 streams land a 400 MB group
        # every 0.82 ms alike (~480 GB/s each way while level 1 and the bucket hash run) — the engines are not what limits it.
        n_copy = int(os.environ.get("KRISP_COPY_STREAMS", "1"))
        n_copy = max(1, min(n_copy, 4 if world == 2 else world - 1))
        sides = searcher.__dict__.setdefault("_copy_streams", [])
        while len(sides) < 1 + n_copy:
            sides.append(torch.cuda.Stream(device=device))
        vote, side = sides[0], sides[1]
        copies = sides[1:1 + n_copy]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        searcher._set_have_outgroup(have_outgroup)
        ev[0].record(main)
        cur_ptr = searcher.shard_slab_extract()                       # K1 + level 0 (asynchronous, main stream)
        ev[1].record(main)
        cur = _wrap(cur_ptr, nd, device)
        gathered = torch.empty(world * nd, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(gathered, cur, group=group)       # fill levels, on the device
        for cs in copies:
            cs.wait_event(ev[1])
        landed = []
        flag = searcher.__dict__.setdefault("_flag", torch.zeros(1, dtype=torch.int32, device=device))
        # "copy": bulk peer copies on the copy engines (no SM taken from the kernels that run underneath); "a2a": one NCCL all-to-all per
        # group.  Measured on 8 x B200 (0.2 Gbp per GPU): 8.08 ms vs 8.14 - 8.24 ms per search; on 2 GPUs 6.5 vs 8.7 ms
        mode = os.environ.get("KRISP_SLAB_EXCHANGE", "copy")
        if mode == "a2a":
            # one NCCL all-to-all per digit group between the library's buffers (send/recv over NVLink on NCCL's channels): its
            # completion on this rank IS "my slabs of the group have arrived", no separate vote
            stg, rcv, cap0, _ = searcher.shard_slab_buffers()
            ck = (stg, rcv, cap0, n_groups, world, rank, nd)
            lists = searcher.__dict__.get("_a2a_lists")
            if lists is None or lists[0] != ck:                       # (tensor views of the library's buffers: built once per plan)
                firsts = [first_digit(o, world, nd) for o in range(world + 1)]
                dps_me = firsts[rank + 1] - firsts[rank]
                empty = torch.empty(0, dtype=torch.int64, device=device)
                per_group = []
                for g in range(n_groups):
                    ins, outs = [], []
                    j0m, j1m = g * dps_me // n_groups, (g + 1) * dps_me // n_groups
                    for o in range(world):
                        dps_o = firsts[o + 1] - firsts[o]
                        j0, j1 = g * dps_o // n_groups, (g + 1) * dps_o // n_groups
                        ins.append(empty if o == rank or j1 <= j0 else _wrap(stg + 8 * (firsts[o] + j0) * cap0, (j1 - j0) * cap0, device))
                        outs.append(empty if o == rank or j1m <= j0m else _wrap(rcv + 8 * (o * dps_me + j0m) * cap0, (j1m - j0m) * cap0, device))
                    per_group.append((outs, ins))
                lists = (ck, per_group)
                searcher.__dict__["_a2a_lists"] = lists
            with torch.cuda.stream(side):
                for outs, ins in lists[1]:
                    dist.all_to_all(outs, ins, group=group)
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(side)
                    landed.append(e)
        else:
            prev = []
            for g in range(n_groups):
                cur = []
                for i, cs in enumerate(copies):
                    for j, e in enumerate(prev):
                        if j != i:
                            cs.wait_event(e)                           # the other streams' share of group g - 1 first
                    searcher.shard_slab_send(g, n_groups, cs.cuda_stream, i, n_copy if world == 2 else -n_copy)
                    sent = torch.cuda.Event()
                    sent.record(cs)
                    vote.wait_event(sent)
                    cur.append(sent)
                prev = cur if n_copy > 1 else []
                with torch.cuda.stream(vote):
                    dist.all_reduce(flag, group=group)
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(vote)
                    landed.append(e)
        timeline = os.environ.get("KRISP_TIMELINE") == "1"
        marks = []
        if n_groups > 1 and os.environ.get("KRISP_OWN_FIRST", "1") == "1" and hasattr(searcher, "shard_slab_own"):
            # this rank's own slabs are complete: their level 1 runs while the first digit group is still on the wire
            searcher.shard_slab_own(gathered.data_ptr())
        for g in range(n_groups):
            main.wait_event(landed[g])
            if timeline:
                a = torch.cuda.Event(enable_timing=True)
                a.record(main)
            searcher.shard_slab_level(gathered.data_ptr(), g, n_groups)
            if timeline:
                b = torch.cuda.Event(enable_timing=True)
                b.record(main)
                marks.append((a, b))
        ev[2].record(main)
        res, status = searcher.shard_slab_finish(have_outgroup=have_outgroup)
        st = searcher.__dict__.setdefault("_status", torch.zeros(1, dtype=torch.int64, device=device))
        st.fill_(status)                                              # (a fill kernel, not a pageable host -> device copy)
        dist.all_reduce(st, op=dist.ReduceOp.MAX, group=group)        # same decision everywhere; also: nobody starts the next
        status = int(st.item())                                       # search's copies while a peer still reads its buffer
        if status == 0:
            lo_d, hi_d = first_digit(rank, world, nd), first_digit(rank + 1, world, nd)
            items = searcher.shard_slab_buffers()[3]                  # one 8-byte item per window instead of two records
            elems = int(res.n_records) // (2 if items else 1)
            sent = elems - elems * (hi_d - lo_d) // max(nd, 1)        # (uniform digits: what does not stay)
            t_x = ev[1].elapsed_time(landed[-1])
            res.profile = list(res.profile) + [("K4 exchange (bulk peer copies, first send to last landed)", t_x)]
            res.exchange = {"slab": True, "digits": nd, "groups": n_groups, "records_extracted": int(res.n_records), "own_digits": [lo_d, hi_d],
                            "sent": sent, "sent_bytes": 8 * sent, "copied_bytes": 8 * (cap - 4096) * (world - 1) // max(world, 1),
                            "exchange_ms": t_x, "mode": mode, "window_items": items}
            if timeline:                                              # ms since the search began (this rank's clock)
                res.exchange["timeline"] = {"k1_end": ev[0].elapsed_time(ev[1]), "landed": [ev[0].elapsed_time(e) for e in landed],
                                            "level_begin": [ev[0].elapsed_time(a) for a, _ in marks],
                                            "level_end": [ev[0].elapsed_time(b) for _, b in marks], "levels_done": ev[0].elapsed_time(ev[2])}
            return res
        if status == 2:
            return None                                               # a slab overflowed somewhere: exact exchange for everybody
        if status == 3:
            cached = searcher.__dict__.get("_slab_plan")
            continue                                                  # a survivor table was grown: once more, same plan
        extra += 2                                                    # plan too coarse somewhere: two more bucket bits for everybody
        searcher.__dict__["_shard_bb_extra"] = extra
        key = key[:-1] + (extra,)
    return None


def shutdown(searcher, group=None):
    """Collective teardown: every rank unmaps its peers' receive buffers, then a barrier, then the contexts may be destroyed
    (exported memory must not be freed while a peer still maps it or still stores into it)."""
    import torch.distributed as dist
    searcher.__dict__.pop("_slab_plan", None)
    if hasattr(searcher, "synchronize"):
        searcher.synchronize()
    if hasattr(searcher, "shard_ipc_close"):
        searcher.shard_ipc_close()
    dist.barrier(group=group)
    searcher.__dict__.pop("_ipc_state", None)
    searcher.close()


def replicate_sequences(searcher, device, group=None):
    """Multi-word records (k > 28): every rank gets ALL sequences, in the same order (rank 0's files, rank 1's files, ...), and
    keeps extracting only its own.  The 8-byte elements that travel then carry a position that means the same on every GPU, and
    the owner of a key can rebuild the records it still needs from the bytes (csrc/kb_prefilter.cuh).
    One all-gather of the sequence bytes over NCCL (padded to the longest rank), then device-to-device adds."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ptr, n = searcher.sequence_buffer()
    table = [None] * world
    dist.all_gather_object(table, (list(searcher.added_ids), searcher.sequence_sizes(), n), group=group)
    longest = max(t[2] for t in table)
    mine = torch.zeros(max(longest, 1), dtype=torch.uint8, device=device)
    if n:
        mine[:n] = torch.as_tensor(_DeviceBytes(ptr, n), device=device)
    allb = torch.empty(world * max(longest, 1), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allb, mine, group=group)
    torch.cuda.current_stream().synchronize() if device.type == "cuda" else None
    searcher.clear_sequences()
    searcher.reserve(sum(t[2] for t in table) + 64)
    first = count = 0
    base = allb.data_ptr()
    n_added = 0
    for r, (ids, sizes, _) in enumerate(table):
        off = r * max(longest, 1)
        if r == rank:
            first, count = n_added, len(ids)
        for gid, sz in zip(ids, sizes):
            searcher.add_sequence(gid, (base + off, sz - 1))          # (the library appends the separator itself)
            off += sz
            n_added += 1
    searcher.synchronize()                                            # the adds copy out of `allb`, which dies with this frame
    searcher.shard_own_files(first, count)
    return sum(t[2] for t in table)


class _DeviceBytes:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def sharded_search(searcher, device, have_outgroup=True, group=None, total_bases=None, exchange_mode=None):
    """Steps 0-3 on the sequences this rank has added to `searcher`.  Returns this rank's SearchResult.
    `total_bases` = bases over all ranks (all-reduced from ``searcher.bases_added`` when not given).
    `exchange_mode`: "p2p" = fused partition + exchange over peer memory (default on CUDA), "nccl" = all-to-all."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    on_cuda = device is not None and getattr(device, "type", "cpu") == "cuda"
    if on_cuda and hasattr(searcher, "set_stream"):
        # the library's kernels must be ordered with the collectives issued here: both on torch's current stream
        searcher.set_stream(torch.cuda.current_stream(device).cuda_stream)
    lo = getattr(searcher, "lo", None)
    multiword = lo is not None and 2 * sum(lo) + 8 > 64
    if exchange_mode is None:
        exchange_mode = os.environ.get("KRISP_EXCHANGE")
    if exchange_mode in (None, "slab") and on_cuda and not multiword and hasattr(searcher, "shard_slab_plan"):
        res = slab_search(searcher, device, have_outgroup, group, total_bases)
        if res is not None:
            return res
        exchange_mode = "p2p"                                         # slab exchange not applicable / overflowed: exact exchange
    if exchange_mode == "slab":
        exchange_mode = None
    if lo is not None and 2 * sum(lo) + 8 > 64 and hasattr(searcher, "sequence_buffer"):
        if getattr(searcher, "_replicated", None) is not searcher.added_ids:
            replicate_sequences(searcher, device, group)
            searcher._replicated = searcher.added_ids                 # (a second search on the same sequences does not gather again)
        total_bases = int(searcher.bases_added)                       # every rank now holds everything
    if total_bases is None:
        t = torch.tensor([int(getattr(searcher, "bases_added", 0))], dtype=torch.int64, device=device)
        dist.all_reduce(t, group=group)
        total_bases = int(t.item())
    searcher.shard_plan(world, rank, total_bases)
    if exchange_mode is None:
        exchange_mode = "p2p" if on_cuda and hasattr(searcher, "shard_scatter") else "nccl"
    if exchange_mode == "p2p":
        return direct_search(searcher, device, have_outgroup, group)
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
