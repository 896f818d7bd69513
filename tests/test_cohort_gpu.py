"""Cohort mode on real GPUs (BASELINE.json config 5): the table hash-sharded over 2 GPUs, one process per
GPU.  Peer loads (CUDA IPC over NVLink) and the explicit NCCL all-to-all exchange must both return
what one unsharded table returns, and find_batch over the sharded table must print the same rows.
Needs a box with >= 2 GPUs (`gpurun --gpus 2`); skipped elsewhere."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
WORLD = 2


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _say(rank, what):
    if os.environ.get("KM_TEST_VERBOSE"):
        print("[rank %d] %s" % (rank, what), flush=True)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from km_b200 import cohort, engine, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n_bg = 3_000_000
    panel = synth.make_panel(120, seed=5)
    # every rank streams the SAME keys; each keeps what it owns
    shard = cohort.ShardedTable.create(rank, world, capacity_per_shard=(n_bg + len(panel.keys)) // world + (1 << 16))
    shard.build_synthetic(synth.TABLE_SEED, n_bg)
    shard.insert(panel.keys, panel.counts, mode="overwrite")
    kept = shard.info()["n_keys"]
    _say(rank, "shard built: %d keys" % kept)
    total = torch.tensor([kept], dtype=torch.int64, device="cuda")
    dist.all_reduce(total)
    whole = engine.Table.create(capacity=n_bg + len(panel.keys), device=rank)       # the unsharded answer
    whole.build_synthetic(synth.TABLE_SEED, n_bg)
    whole.insert(panel.keys, panel.counts, mode="overwrite")
    assert int(total.item()) == whole.info()["n_keys"]
    assert 0.4 < kept / whole.info()["n_keys"] < 0.6                                 # the hash splits evenly

    q = synth.lookup_queries(1 << 18, synth.TABLE_SEED, n_bg, seed=100 + rank)
    want = whole.query_packed(q)
    _say(rank, "unsharded answers ready")
    routed = shard.query_routed(q, dist)                                             # NCCL all-to-all, no peer mapping yet
    _say(rank, "routed done")
    assert (routed == want).all()
    shard.attach(dist)
    _say(rank, "attached")
    assert (shard.query_packed(q) == want).all()                                     # peer loads over NVLink
    mine = cohort.shard_targets([len(s) for s in panel.targets], world)[rank]
    seqs = [panel.targets[i] for i in mine]
    names = [panel.names[i] for i in mine]
    a = shard.find_batch(seqs, want_graph=False).format_all("panel.jf", names)
    b = whole.find_batch(seqs, want_graph=False).format_all("panel.jf", names)
    assert a == b and a.count("\n") >= len(seqs)
    # default regime through the same ranks: contiguous shares of the targets, replicated table, ONE km_find_text per
    # rank, the byte buffers gathered on rank 0 in rank order = input order
    packed = engine.PackedTargets(panel.targets, panel.names)
    text, status = cohort.find_mutation_sharded(whole, packed, "panel.jf", dist)
    if rank == 0:
        one, st1 = whole.find_text(packed, "panel.jf", as_bytes=True)
        assert np.array_equal(text, one) and np.array_equal(status, st1)
    else:
        assert text is None and status is None
    # the same against the SHARDED table (peer loads inside the kernels)
    text2, _ = cohort.find_mutation_sharded(shard, packed, "panel.jf", dist)
    if rank == 0:
        assert np.array_equal(text2, one)
    # routed counting: every rank feeds ITS OWN reads, keys travel to their owner by atomics over NVLink
    counted = cohort.ShardedTable.create(rank, world, capacity_per_shard=(len(panel.keys) * 2) // world + (1 << 16))
    counted.attach(dist)
    counted.set_routing(True)
    share = list(range(rank, len(panel.targets), world))
    counted.count_text(synth.sample_reads(panel, share))
    torch.cuda.synchronize()
    dist.barrier()
    counted.drop_below(2)
    dist.barrier()
    n_here = torch.tensor([counted.info()["n_keys"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(n_here)
    assert int(n_here.item()) == len(panel.keys)
    assert (counted.query_packed(panel.keys) == panel.counts).all()          # peer loads: every key, wherever it lives
    # device-side explicit exchange == peer loads
    qd = torch.from_numpy(q.view(np.int64)).cuda()
    routed_dev = shard.query_routed_device(qd, dist).cpu().numpy().view(np.uint32)
    assert (routed_dev == want).all()
    dist.barrier()
    counted.close()
    _say(rank, "sharded targets done")
    dist.barrier()
    with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
        f.write("%d %d\n" % (kept, len(seqs)))
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < WORLD, reason="needs %d GPUs" % WORLD)
def test_sharded_table_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    from km_b200 import build as kb
    kb.build()
    mp.spawn(_worker, args=(WORLD, _free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(WORLD))


@pytest.mark.skipif(_n_gpus() < WORLD, reason="needs %d GPUs" % WORLD)
def test_cli_find_mutation_gpus_switch_prints_the_single_gpu_text(synth_small):
    """`km find_mutation --gpus 2 <targets> <db.jf>`: the targets dealt to two GPUs (one process each, table loaded on
    both), rank 0 prints -- byte for byte what one GPU prints, minus the volatile lines and the `#gpus:2` echo."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(extra):
        out = subprocess.run([sys.executable, "-m", "km_b200", "find_mutation", *extra, *synth_small["files"][:40], synth_small["jf"]],
                             cwd=root, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return [l for l in out.stdout.split("\n") if l and not l.startswith("#Elapsed") and not l.startswith("#func:") and not l.startswith("#gpus:")]
    one = run([])
    two = run(["--gpus", "2"])
    assert one == two and sum(1 for l in one if not l.startswith("#")) > 40
