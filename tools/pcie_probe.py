"""How long does one rank's text (13.2 MB, device -> pinned host) and input (3.4 MB, host -> device) take when the rank copies
alone and when all ranks of the box copy at the same moment?  (The e2e step of bench.py ends with that D2H on every rank.)
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D2H, H2D, REP = 13_200_000, 3_400_000, 50
dev = torch.empty(D2H, dtype=torch.uint8, device="cuda")
host = torch.empty(D2H, dtype=torch.uint8).pin_memory()
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
din = torch.empty(H2D, dtype=torch.uint8, device="cuda")


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(REP):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / REP * 1e3


def d2h():
    host.copy_(dev, non_blocking=True)


def h2d():
    din.copy_(hin, non_blocking=True)


for _ in range(5):
    d2h(); h2d()
torch.cuda.synchronize()
alone = torch.zeros(world, 2, device="cuda")
for r in range(world):                       # one rank at a time
    dist.barrier()
    if r == rank:
        alone[r, 0] = timed(d2h)
        alone[r, 1] = timed(h2d)
dist.barrier()
together = torch.tensor([timed(d2h), timed(h2d)], device="cuda")
dist.all_reduce(alone)
gathered = [torch.zeros(2, device="cuda") for _ in range(world)]
dist.all_gather(gathered, together)
if rank == 0:
    print("rank  D2H 13.2 MB alone / all ranks at once (ms)     H2D 3.4 MB alone / at once (ms)")
    for r in range(world):
        print("%4d  %8.3f / %8.3f   (%.1f / %.1f GB/s)        %8.3f / %8.3f" % (
            r, alone[r, 0].item(), gathered[r][0].item(), D2H / alone[r, 0].item() / 1e6, D2H / gathered[r][0].item() / 1e6,
            alone[r, 1].item(), gathered[r][1].item()))
    print("host cores:", os.cpu_count())
dist.destroy_process_group()
