import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vector-database_b200"))
import numpy as np, torch, torch.distributed as dist
import vectorindex as vi
from vectorindex import synthetic as ds
from vectorindex.distributed import Collectives
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, d = int(sys.argv[1]), int(sys.argv[2])
ids, rows = ds.unit_gaussian(n, d, seed=3)
lo, hi = rank * n // world, (rank + 1) * n // world
ctx = vi.Context(rank)
ctx.reserve(hi - lo, d); ctx.add(ids[lo:hi], rows[lo:hi])
coll = Collectives(torch.device("cuda", rank)); coll.attach(ctx)
info = ctx.build(vi.MODE_FAST)
rid, dim, mid, oid = ctx.ranges()
sh = ctx.shared_rows
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: rows {len(rid)} shared {sh} calls {coll.calls}", flush=True)
        for k in range(min(sh, 8)):
            print("   ", k, int(rid[k]), int(dim[k]), float(mid[k]), int(oid[k]), flush=True)
        print("    first own rows:", [(int(rid[k]), int(dim[k])) for k in range(sh, min(sh + 4, len(rid)))], flush=True)
dist.barrier()
