import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
from krisp_b200 import sharded
from krisp_b200.search import Searcher
from krisp_b200.panel import make_panel
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
s = Searcher(device=local, stream=torch.cuda.current_stream().cuda_stream)
gs = make_panel(5, 4, 200_000)
s.configure(25, 1, 2, [1 if g.is_ingroup else 0 for g in gs])
s.clear_sequences()
for i, g in enumerate(gs):
    if i % world == rank: s.add_sequence(i, np.frombuffer(g.joined(), dtype=np.uint8))
t = torch.tensor([s.bases_added], dtype=torch.int64, device=dev); dist.all_reduce(t)
print("plan nd", s.shard_plan(world, rank, int(t.item())), flush=True)
d = s.shard_count(); print("digits", d, flush=True)
c = s.shard_child_counts(); print("child", None if c is None else (c.size, int(c.sum())), flush=True)
try:
    res = sharded.sharded_search(s, dev, have_outgroup=True)
    print("rows", len(res.rows()))
except Exception as e:
    print("ERR", e, flush=True)
dist.destroy_process_group()
