"""Per-rank pinned H2D rate with every rank copying at once (run under torchrun): which GPUs of the box have the slow links,
and whether write-combined source buffers change it (MNV1_H2D_WC=1).  Prints one JSON line from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mnv1_b200 as mn  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mn.Context(local, mn.BF16)
nbytes = 256 * 224 * 224 * 3
out = {}
for reps in (21, 101):
    dist.barrier()
    g = torch.tensor([ctx.h2d_probe(nbytes, reps)], dtype=torch.float64, device="cuda")
    allg = [g.clone() for _ in range(world)]
    dist.all_gather(allg, g)
    out[f"reps{reps}"] = [round(float(t.item()), 1) for t in allg]
# one rank at a time: the link of each GPU alone
alone = []
for r in range(world):
    dist.barrier()
    v = ctx.h2d_probe(nbytes, 21) if r == rank else 0.0
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    alone.append(round(float(t.item()), 1))
out["alone"] = alone
if rank == 0:
    line = json.dumps({"wc": bool(os.environ.get("MNV1_H2D_WC")), **out})
    print(line, flush=True)
    if len(sys.argv) > 1:                      # NCCL prints its version after us: keep a copy in a file
        with open(sys.argv[1], "a") as f:
            f.write(line + "\n")
ctx.close()
dist.destroy_process_group()
