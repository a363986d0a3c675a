"""torchrun worker for tests/test_gpu_sharded.py: flank-hash sharded search over NCCL on the golden cases."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from krisp_b200 import ingest, sharded  # noqa: E402
from krisp_b200.search import Searcher, labels_for  # noqa: E402
from tests.helpers import deduce_ldr, golden_paths, load_golden  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    s = Searcher(device=local)                         # (own stream: sharded_search moves it onto torch's current stream itself)
    bad = 0
    for case in load_golden()["cases"]:
        L, D, R = deduce_ldr(case["flags"])
        if R == 0 and D > 0:
            continue                                   # quirk S9: the host answers without a search
        ins, outs = golden_paths(case)
        files = ins + outs
        _, is_in = labels_for(ins, outs)
        owner = sharded.assign_files(len(files), world, sizes=[os.path.getsize(f) for f in files])
        s.configure(L, D, R, is_in, case["omit_soft"])
        s.clear_sequences()
        for i, f in enumerate(files):
            if owner[i] == rank:
                s.add_sequence(i, ingest.load_file(f)[0])
        for mode in ("slab", "p2p", "nccl"):           # K1-fused slab exchange / fused partition + exchange over peer memory / NCCL all-to-all
            res = sharded.sharded_search(s, dev, have_outgroup=len(outs) > 0, exchange_mode=mode)
            rows = sharded.gather_rows(res.rows())
            ok = len(rows) == case["n_rows"] and hashlib.sha256("\n".join(rows).encode()).hexdigest() == case["rows_sha256"]
            if rank == 0:
                print(("ok   " if ok else "FAIL ") + case["name"], mode, len(rows), flush=True)
            bad += 0 if ok else 1
    # seeded panels, big enough for multi-level plans; options force the plan shapes of the large runs: (4, 6, 6) bits = K1 counts the
    # level-1 children too and the counts travel with the exchange; rows must equal the single-GPU search of the whole panel
    from krisp_b200.panel import make_panel
    for (L, D, R), opts in [((25, 1, 2), {}), ((25, 1, 2), {"bucket_bits": 16, "shard_bits0": 4}), ((25, 1, 2), {"bucket_bits": 16, "shard_bits0": 4, "fused_hist": 0}),
                            ((25, 1, 2), {"bucket_bits": 12}), ((25, 1, 2), {"bucket_bits": 20, "shard_bits0": 5}), ((25, 1, 2), {"bucket_bits": 14, "hash_shared": 0, "hash_slots_log2": 6}),
                            ((25, 1, 2), {"slab_cap": 2}), ((25, 1, 2), {"slab": 0}), ((25, 1, 2), {"sym": 1}), ((25, 1, 2), {"sym": 1, "bucket_bits": 16}), ((12, 3, 12), {"sym": 1, "bucket_bits": 10}), ((12, 3, 12), {"bucket_bits": 10}),
                            ((32, 60, 32), {}), ((32, 60, 32), {"bucket_bits": 16, "shard_bits0": 4})]:
        gs = make_panel(5, 4, 200_000)
        is_in = [1 if g.is_ingroup else 0 for g in gs]
        s.configure(L, D, R, is_in)
        s.clear_sequences()
        for i, g in enumerate(gs):
            s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
        want = s.search(have_outgroup=True).rows()
        for k, v in opts.items():
            s.set_option(k, v)
        try:
            s.clear_sequences()
            for i, g in enumerate(gs):
                if i % world == rank:
                    s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
            res = sharded.sharded_search(s, dev, have_outgroup=True)
            rows = sharded.gather_rows(res.rows())
        finally:
            for k, v in (("bucket_bits", -1), ("shard_bits0", 0), ("fused_hist", 1), ("slab_cap", 0), ("slab", 1), ("sym", -1), ("hash_shared", -1), ("hash_slots_log2", 0)):
                s.set_option(k, v)
        ok = rows == want and len(want) > 0
        if rank == 0:
            print(("ok   " if ok else "FAIL ") + f"panel {L}/{D}/{R} {opts}", len(rows), flush=True)
        bad += 0 if ok else 1
    # a three-level plan with digit groups: level 2 runs per group too, into a buffer of its own (the staging buffer is still being
    # copied out of); the forced slab capacity is a multiple of the partition tile, which is what lets tiles map to parent slabs
    gs = make_panel(5, 4, 50_000)
    is_in = [1 if g.is_ingroup else 0 for g in gs]
    s.configure(25, 1, 2, is_in)
    s.clear_sequences()
    for i, g in enumerate(gs):
        s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
    want = s.search(have_outgroup=True).rows()
    for sym in (0, 1):
        for k, v in (("bucket_bits", 12), ("shard_bits0", 2), ("slab_cap", 262144), ("sym", sym), ("profile", 1)):
            s.set_option(k, v)
        try:
            s.clear_sequences()
            for i, g in enumerate(gs):
                if i % world == rank:
                    s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
            res = sharded.sharded_search(s, dev, have_outgroup=True)
            rows = sharded.gather_rows(res.rows())
        finally:
            for k, v in (("bucket_bits", -1), ("shard_bits0", 0), ("slab_cap", 0), ("sym", -1), ("profile", 0)):
                s.set_option(k, v)
        stages = [nm for nm, _ in res.profile]
        ex = getattr(res, "exchange", None) or {}
        grouped3 = any(nm.startswith("K2 partition 2 (group") for nm in stages)
        ok = rows == want and len(want) > 0 and (grouped3 or ex.get("groups", 1) == 1)
        if rank == 0:
            print(("ok   " if ok else "FAIL ") + f"three-level grouped plan, sym {sym}: groups {ex.get('groups')}, level 2 per group: {grouped3}", len(rows), flush=True)
        bad += 0 if ok else 1
    # divergent genomes (3 % private substitutions): the sharded plan is too coarse at first; every rank must vote for the re-plan
    # and the rows must still be the single-GPU rows
    gs = make_panel(7, 5, 400_000, noise=3e-2, snp_every=200)
    is_in = [1 if g.is_ingroup else 0 for g in gs]
    s.configure(25, 1, 2, is_in)
    s.clear_sequences()
    for i, g in enumerate(gs):
        s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
    want = s.search(have_outgroup=True).rows()
    s.clear_sequences()
    for i, g in enumerate(gs):
        if i % world == rank:
            s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
    for rep in range(2):
        res = sharded.sharded_search(s, dev, have_outgroup=True)
        rows = sharded.gather_rows(res.rows())
        ok = rows == want                                 # (possibly no row at all: almost every 28-mer is private)
        if rank == 0:
            print(("ok   " if ok else "FAIL ") + f"divergent panel, search {rep}, exchange {getattr(res, 'exchange', None)}", len(rows), flush=True)
        bad += 0 if ok else 1
    sharded.shutdown(s)
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
