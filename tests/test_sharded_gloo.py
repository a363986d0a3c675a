"""world_size-2 `gloo` test of the multi-GPU host logic (file assignment, counts + records all-to-all,
per-shard search, row gather) on CPU.  The per-rank "searcher" here is a TEST DOUBLE built on the oracle
model: it packs records exactly like K1 (mixed flank key | middle | file id) and partitions them with the
host restatement of the device's level-0 digit, so the exchange carries the real record format."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from krisp_b200 import sharded
from krisp_b200.search import labels_for
from oracle import model
from tests.helpers import deduce_ldr, golden_paths, load_golden

MASK64 = (1 << 64) - 1
C1, C2 = 0x9E3779B97F4A7C15, 0xD6E8FEB86659FD93


def kb_mix(x, nb):
    """csrc/kb_common.cuh kb_mix, restated."""
    return (x * C1) & ((1 << nb) - 1)


class FakeSearcher:
    def __init__(self, L, D, R, is_ingroup, files, omit_soft):
        self.L, self.D, self.R, self.is_in, self.files, self.omit = L, D, R, is_ingroup, files, omit_soft
        self.FB = 2 * (L + R)

    def _pack(self, left, mid, right, fid):
        code = {"A": 0, "C": 1, "G": 2, "T": 3}
        key = 0
        for ch in left + right:
            key = (key << 2) | code[ch]
        m = 0
        for ch in mid:
            m = (m << 2) | code[ch]
        v = fid
        if self.FB:
            v |= kb_mix(key, self.FB) << (64 - self.FB)
        if self.D:
            v |= m << (64 - self.FB - 2 * self.D)
        return v

    def shard_plan(self, world, rank, total_bases):
        # any plan every rank agrees on: 2^bits0 level-0 digits, at least one per shard
        self.world, self.rank = world, rank
        self.bits0 = max(3, (world - 1).bit_length())
        assert self.FB > self.bits0 and total_bases > 0
        return 1 << self.bits0

    def shard_extract(self):
        world = self.world
        recs = []
        for fid, path in self.files:
            for line in model.kmer_lines(model.fasta_records(model.read_lines(path)), self.L, self.D, self.R, self.omit):
                left, mid, right = line.split(",")
                recs.append(self._pack(left, mid, right, fid))
        arr = np.array(recs, dtype=np.uint64)
        nd = 1 << self.bits0
        digit = sharded.digit_of_key(arr >> np.uint64(64 - self.FB), self.FB, self.bits0) if arr.size else np.zeros(0, np.uint64)
        order = np.argsort(digit, kind="stable")
        self.send = arr[order].view(np.int64).copy()
        digits = [int((digit == d).sum()) for d in range(nd)]
        firsts = [sharded.first_digit(s, world, nd) for s in range(world + 1)]
        return "send", [sum(digits[firsts[s]:firsts[s + 1]]) for s in range(world)], digits

    def shard_recv_buffer(self, n):
        self.recv = np.zeros(n, dtype=np.int64)
        return "recv"

    def wrap_records(self, handle, n, device):
        return torch.from_numpy(self.send if handle == "send" else self.recv)[:n]

    def shard_search(self, n, pieces, have_outgroup=True):
        L, D, R = self.L, self.D, self.R
        # the pieces must describe what arrived: (source, digit) runs in order, all digits inside this shard's range
        nd = 1 << self.bits0
        lo, hi = sharded.first_digit(self.rank, self.world, nd), sharded.first_digit(self.rank + 1, self.world, nd)
        assert len(pieces) == self.world * (hi - lo) and sum(pieces) == n
        got = self.recv[:n].view(np.uint64)
        pos = 0
        for i, c in enumerate(pieces):
            d = lo + i % (hi - lo)
            assert np.all((got[pos:pos + c] >> np.uint64(64 - self.bits0)) == np.uint64(d))
            pos += c
        groups = {}
        for v in self.recv[:n].view(np.uint64).tolist():
            key = v >> (64 - self.FB)
            mid = (v >> (64 - self.FB - 2 * D)) & ((1 << (2 * D)) - 1) if D else 0
            groups.setdefault(key, []).append((mid, v & 0xFF))
        rows, n_files = [], len(self.is_in)
        inv1, inv2 = pow(C1, -1, 1 << 64), pow(C2, -1, 1 << 64)
        for key, occ in groups.items():
            if len({f for _, f in occ}) != n_files:
                continue
            ins = [m for m, f in occ if self.is_in[f]]
            outs = [m for m, f in occ if not self.is_in[f]]
            col = lambda m, c: (m >> (2 * (D - 1 - c))) & 3
            if D and not any({col(m, c) for m in ins}.isdisjoint({col(m, c) for m in outs}) for c in range(D)):
                continue
            x = (key * inv1) & ((1 << self.FB) - 1)
            flank = "".join("ACGT"[(x >> (2 * (L + R - 1 - i))) & 3] for i in range(L + R))
            use = ins if have_outgroup else ins + outs
            cons = "".join(model.IUPAC_KEY[tuple(sorted({"ACGT"[col(m, c)] for m in use}))] for c in range(D))
            rows.append(f"{flank[:L]},{cons},{flank[L:]}")

        class R_:
            pass
        r = R_()
        r.rows = lambda: sorted(rows)
        return r


def _worker(rank, world, port, case, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ins, outs = golden_paths(case)
        L, D, R = deduce_ldr(case["flags"])
        files = ins + outs
        _, is_in = labels_for(ins, outs)
        owner = sharded.assign_files(len(files), world, sizes=[os.path.getsize(f) for f in files])
        mine = [(i, f) for i, f in enumerate(files) if owner[i] == rank]
        s = FakeSearcher(L, D, R, is_in, mine, case["omit_soft"])
        res = sharded.sharded_search(s, torch.device("cpu"), have_outgroup=len(outs) > 0, total_bases=sum(os.path.getsize(f) for f in files))
        rows = sharded.gather_rows(res.rows())
        if rank == 0:
            out_q.put(rows)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("name", ["c1_spacer_25_1_2", "p_10_1_2", "p_5_2_3"])
def test_two_rank_sharded_search_equals_reference_rows(name):
    case = next(c for c in load_golden()["cases"] if c["name"] == name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rows == case["rows"]
