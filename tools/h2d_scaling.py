"""Pinned host-to-device copy bandwidth with N ranks copying at the same time (no kernels): the box's ceiling for the streaming
(e2e) path.  Launch with torchrun:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_scaling.py
Rank 0 prints one JSON line: per-rank and aggregate GB/s for one copy stream and for two, 1 GiB and 64 MiB transfers."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.empty(1 << 28, dtype=torch.float32).pin_memory()          # 1 GiB
x.fill_(1.0)                                                         # first touch by this rank
y = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
res = {}
for name, nstreams, elems in (("1GiB_1stream", 1, 1 << 28), ("1GiB_2streams", 2, 1 << 28), ("64MiB_chunks_1stream", 1, 1 << 24)):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]

    def once():
        per = (1 << 28) // nstreams
        for si, s in enumerate(streams):
            with torch.cuda.stream(s):
                for off in range(si * per, (si + 1) * per, min(elems, per)):
                    n = min(elems, per)
                    y[off:off + n].copy_(x[off:off + n], non_blocking=True)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(8):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    gbs = 8 * (1 << 30) / dt / 1e9
    if world > 1:
        all_g = [None] * world
        dist.all_gather_object(all_g, gbs)
    else:
        all_g = [gbs]
    res[name] = {"per_rank_gbs": [round(g, 2) for g in all_g], "aggregate_gbs": round(sum(all_g), 1)}
if rank == 0:
    print(json.dumps({"ranks": world, "h2d": res}), flush=True)
if world > 1:
    dist.destroy_process_group()
