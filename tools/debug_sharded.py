"""Debug helper: sharded search on a synthetic 20+20 panel under torchrun; prints stats per rank; checks vs the oracle when small."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from krisp_b200 import sharded
from krisp_b200.search import Searcher
import bench

def main():
    glen = int(sys.argv[1]); check = len(sys.argv) > 2 and sys.argv[2] == "check"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mine = bench.build_panel(glen, rank, world)
    s = Searcher(device=local, stream=torch.cuda.current_stream().cuda_stream)
    s.configure(25, 1, 2, [1] * 20 + [0] * 20)
    s.clear_sequences()
    for gid, _, arr in mine:
        s.add_sequence(gid, arr)
    t0 = time.time()
    res = sharded.sharded_search(s, dev)
    torch.cuda.synchronize()
    print(f"rank {rank}: n_records={res.n_records} groups={res.n_groups} stats={res.stats} dt={time.time()-t0:.3f}", flush=True)
    rows = sharded.gather_rows(res.rows())
    if rank == 0:
        print("total rows", len(rows), flush=True)
        if check:
            from krisp_b200.panel import make_panel
            from oracle import oracle
            gs = make_panel(20, 20, glen)
            recs = [[r.tobytes() for r in g.records] for g in gs]
            want, _ = oracle.search_records(recs, [g.name for g in gs], {g.name for g in gs if g.is_ingroup}, True, 25, 1, 2)
            print("oracle rows", len(want), "equal", want == rows, flush=True)
    s.close()
    dist.destroy_process_group()

main()
