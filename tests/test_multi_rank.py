"""Host-side multi-GPU logic on CPU: two ranks over gloo.  The device work is
replaced by stand-ins (a dict per shard, a formatter per rank); what is under test is the target
dealing, the in-order gather and the all-to-all query routing of km_b200/cohort.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from km_b200 import build as kb
from km_b200 import cohort

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, *args):
    kb.build()
    mp.spawn(fn, args=(WORLD, _free_port()) + args, nprocs=WORLD, join=True)


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def test_shard_targets_is_a_balanced_partition():
    rng = np.random.default_rng(3)
    lengths = rng.integers(62, 401, size=1000)
    for world in (1, 2, 4, 8):
        parts = cohort.shard_targets(lengths, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(1000))
        loads = [int((lengths[p] - 30).sum()) for p in parts]
        assert max(loads) - min(loads) <= 400            # within one target of each other
        assert all((np.diff(p) > 0).all() for p in parts)


def _both_worker(rank, world, port, n):
    """gather + routing in ONE pair of processes (starting them is most of the test's time)"""
    dist = _init(rank, world, port)
    _gather_body(rank, world, dist, n)
    _route_body(rank, world, dist)
    dist.barrier()
    dist.destroy_process_group()


def _gather_body(rank, world, dist, n):
    lengths = [62 + (7 * i) % 300 for i in range(n)]
    mine = cohort.shard_targets(lengths, world)[rank]
    texts = ["row-of-%d-by-%d\n" % (i, rank) for i in mine.tolist()]
    got = cohort.gather_in_order(texts, mine, n, dist)
    if rank == 0:
        assert len(got) == n
        owner = {}
        for r, part in enumerate(cohort.shard_targets(lengths, world)):
            for i in part.tolist():
                owner[i] = r
        assert got == ["row-of-%d-by-%d\n" % (i, owner[i]) for i in range(n)]
    else:
        assert got is None


def _route_body(rank, world, dist):
    import torch
    rng = np.random.default_rng(1234)                      # the same key universe on every rank
    keys = rng.integers(0, 1 << 62, size=5000, dtype=np.uint64)
    counts = rng.integers(1, 1 << 32, size=5000, dtype=np.uint64).astype(np.uint32)
    owner = cohort.shard_owner(keys, 31, False, world)
    assert owner.min() >= 0 and owner.max() < world and len(set(owner.tolist())) == world
    shard = {int(k): int(c) for k, c, o in zip(keys, counts, owner) if o == rank}     # what this rank stores

    def lookup_local(t):
        asked = t.numpy().view(np.uint64)
        assert all(int(cohort.shard_owner(asked[i:i + 1], 31, False, world)[0]) == rank for i in range(0, len(asked), 97))
        return torch.tensor([shard.get(int(k), 0) for k in asked], dtype=torch.int64)

    qrng = np.random.default_rng(99 + rank)                # different queries per rank: hits, misses, ragged sizes
    n_q = 700 + 300 * rank
    q = np.where(qrng.random(n_q) < 0.6, keys[qrng.integers(0, 5000, n_q)], qrng.integers(0, 1 << 62, n_q, dtype=np.uint64))
    got = cohort.route_queries(q, cohort.shard_owner(q, 31, False, world), dist, lookup_local)
    full = {int(k): int(c) for k, c in zip(keys, counts)}
    want = np.array([full.get(int(k), 0) for k in q], dtype=np.uint32)
    assert (got == want).all()
    # an empty batch on one rank must not stall the other
    e = np.zeros(0, dtype=np.uint64) if rank == 0 else q[:5]
    got = cohort.route_queries(e, cohort.shard_owner(e, 31, False, world), dist, lookup_local)
    assert len(got) == len(e)


def test_rows_come_back_in_order_and_all_to_all_routing_equals_direct_lookup():
    _spawn(_both_worker, 57)


class _FakeTable:
    """stands in for engine.Table on a machine without a GPU: find_text prints one line per target"""
    k = 31
    device = 0

    def __init__(self, fail_on_rank=None, rank=0):
        self.fail = fail_on_rank == rank

    def find_text(self, packed, db_name, as_bytes=False, **params):
        if self.fail:
            raise RuntimeError("boom")
        text = "".join("%s\t%s\t%d\n" % (db_name, n, len(s)) for n, s in zip(packed.names, packed.sequences)).encode()
        status = np.array([len(s) % 3 for s in packed.sequences], dtype=np.uint32)
        return np.frombuffer(text, dtype=np.uint8), status


def _sharded_worker(rank, world, port):
    from km_b200 import engine
    dist = _init(rank, world, port)
    rng = np.random.default_rng(5)
    seqs = ["A" * int(n) for n in rng.integers(62, 400, size=41)]
    names = ["t%03d" % i for i in range(41)]
    packed = engine.PackedTargets(seqs, names)
    text, status = cohort.find_mutation_sharded(_FakeTable(), packed, "db.jf", dist)
    if rank == 0:
        want, st = _FakeTable().find_text(packed, "db.jf")
        assert text.tobytes() == want.tobytes() and status.tolist() == st.tolist()
    else:
        assert text is None and status is None
    # a rank whose call fails must not leave the others waiting in the collective: everybody raises afterwards
    try:
        cohort.find_mutation_sharded(_FakeTable(fail_on_rank=1, rank=rank), packed, "db.jf", dist)
        raised = False
    except RuntimeError as e:
        raised = "rank 1: boom" in str(e)
    assert raised
    # an empty share (more ranks than targets)
    tiny = engine.PackedTargets(seqs[:1], names[:1])
    text, status = cohort.find_mutation_sharded(_FakeTable(), tiny, "db.jf", dist)
    if rank == 0:
        assert text.tobytes().count(b"\n") == 1 and len(status) == 1
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_find_mutation_gathers_text_in_input_order_and_survives_a_failing_rank():
    _spawn(_sharded_worker)


def test_shard_ranges_are_contiguous_and_balanced():
    rng = np.random.default_rng(3)
    lengths = rng.integers(62, 401, size=1000)
    for world in (1, 2, 4, 8):
        cuts = cohort.shard_ranges(lengths, world)
        assert cuts[0] == 0 and cuts[-1] == 1000 and len(cuts) == world + 1 and all(b >= a for a, b in zip(cuts, cuts[1:]))
        loads = [int((lengths[a:b] - 30).sum()) for a, b in zip(cuts, cuts[1:])]
        assert max(loads) - min(loads) <= 800
    assert cohort.shard_ranges([100], 4) in ([0, 0, 0, 0, 1], [0, 1, 1, 1, 1], [0, 0, 0, 1, 1], [0, 0, 1, 1, 1])


def test_owner_is_strand_independent_for_canonical_tables():
    kb.build()
    from km_b200.engine import pack_kmer
    rng = np.random.default_rng(8)
    comp = str.maketrans("ACGT", "TGCA")
    fwd, rev = [], []
    for _ in range(200):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, 31))
        fwd.append(pack_kmer(s))
        rev.append(pack_kmer(s.translate(comp)[::-1]))
    a = cohort.shard_owner(np.array(fwd, dtype=np.uint64), 31, True, 8)
    b = cohort.shard_owner(np.array(rev, dtype=np.uint64), 31, True, 8)
    assert (a == b).all() and len(set(a.tolist())) > 4
