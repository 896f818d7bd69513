"""Cohort mode (config 5) measured on N GPUs of one box, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/cohort_bench.py
Builds a hash-sharded synthetic table (+ the planted panel), then times, with CUDA events / max over ranks:
  * device-resident lookups through PEER LOADS (every rank queries the whole table over NVLink)
  * the same batch through the explicit NCCL all-to-all exchange
  * the panel (find plan) against the sharded table
Rank 0 prints one JSON line."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from km_b200 import cohort, synth          # noqa: E402
from km_b200._lib import check, lib        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--table-keys", type=int, default=2_000_000_000)
    ap.add_argument("--targets", type=int, default=10000)
    ap.add_argument("--queries", type=int, default=1 << 26)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    panel = synth.make_panel(a.targets, seed=synth.PANEL_SEED + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (panel.keys, panel.counts))
    keys = np.concatenate([g[0] for g in gathered])
    counts = np.concatenate([g[1] for g in gathered])
    t0 = time.time()
    shard = cohort.ShardedTable.create(rank, world, capacity_per_shard=(a.table_keys + len(keys)) // world + (1 << 20), device=local)
    shard.build_synthetic(synth.TABLE_SEED, a.table_keys)
    shard.insert(keys, counts, mode="overwrite")
    shard.attach(dist)
    build_s = time.time() - t0

    def tmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # lookups through peer loads
    q = torch.from_numpy(synth.lookup_queries(a.queries, synth.TABLE_SEED, a.table_keys, seed=7 + rank).view(np.int64)).to(dev)
    out = torch.empty(a.queries, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(dev)          # a real stream: a null handle would select the library's own
    torch.cuda.set_stream(s)

    def peer():
        check(lib().km_query_batch_device(shard._h, ctypes.c_void_p(q.data_ptr()), a.queries, ctypes.c_void_p(out.data_ptr()),
                                          ctypes.c_void_p(s.cuda_stream)))
    for _ in range(3):
        peer()
    dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(a.steps):
        peer()
    e1.record(s)
    torch.cuda.synchronize(dev); dist.barrier()
    peer_ms = tmax(e0.elapsed_time(e1)) / a.steps
    # the same with one rank at a time (the others idle): separates fabric contention from the rest
    solo_ms = []
    for r in range(world):
        dist.barrier(); torch.cuda.synchronize(dev)
        if r == rank:
            e0.record(s)
            for _ in range(a.steps):
                peer()
            e1.record(s)
            torch.cuda.synchronize(dev)
            solo_ms.append(e0.elapsed_time(e1) / a.steps)
        dist.barrier()
    solo_ms = tmax(solo_ms[0])
    hits = int((out != 0).sum().item())
    # the same through the explicit exchange (host-staged routing, NCCL all-to-all)
    qh = q.cpu().numpy().view(np.uint64)[: 1 << 22]
    shard.query_routed(qh, dist)
    dist.barrier(); torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    got = shard.query_routed(qh, dist)
    torch.cuda.synchronize(dev); dist.barrier()
    routed_ms = tmax(1e3 * (time.perf_counter() - t1))
    same = bool((got == out[: 1 << 22].cpu().numpy().view(np.uint32)).all())
    # the panel against the sharded table
    plan = shard.plan(panel.targets)
    for _ in range(3):
        plan.launch(s.cuda_stream)
    dist.barrier(); torch.cuda.synchronize(dev)
    e0.record(s)
    for _ in range(a.steps):
        plan.launch(s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize(dev); dist.barrier()
    panel_ms = tmax(e0.elapsed_time(e1)) / a.steps
    km = plan.kernel_ms()
    if rank == 0:
        print(json.dumps({
            "what": "cohort mode: table hash-sharded over %d GPUs, one rank per GPU" % world, "n_gpus": world,
            "table_keys": a.table_keys, "table_build_s": build_s, "shard_keys": shard.info()["n_keys"],
            "peer_loads": {"queries_per_rank": a.queries, "ms": peer_ms, "ms_one_rank_at_a_time": solo_ms, "lookups_per_s_whole_job": world * a.queries / peer_ms * 1e3,
                           "hit_frac": hits / a.queries},
            "all_to_all": {"queries_per_rank": len(qh), "ms": routed_ms, "lookups_per_s_whole_job": world * len(qh) / routed_ms * 1e3,
                           "equals_peer_loads": same, "note": "host-staged routing + 3 NCCL all_to_all_single"},
            "panel": {"targets_per_rank": a.targets, "ms_per_step": panel_ms, "targets_per_s_whole_job": world * a.targets / panel_ms * 1e3,
                      "kernel_ms": {"probe": km[0], "walk": km[1], "graph": km[2]}}}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
