"""Multi-GPU host logic (the reference has no counterpart -- it is single-threaded).

Two regimes, one process per GPU under torch.distributed:

* default -- TARGETS shard, the table is replicated: `shard_targets` deals targets to ranks balanced
  by reference-k-mer count, every rank runs its own batch, `gather_in_order` brings the per-target
  text back to rank 0 in input order.  No collective on the data path.

* cohort / huge table (BASELINE.json config 5) -- the TABLE is hash-sharded, one shard per GPU
  (`ShardedTable`).  Two ways to reach a remote key:
    - peer loads: after `attach()` every shard is mapped into every process (CUDA virtual-memory API, file descriptors) and the probe
      kernels read remote buckets over NVLink directly -- `find_batch` / `query_packed` work
      unchanged and there is no exchange step;
    - explicit exchange: `query_routed` sends each k-mer to its owner and the count back with two
      all-to-all rounds (NCCL on GPUs; the same code runs on gloo for the CPU tests).
"""
import ctypes
import heapq

import numpy as np

from . import engine
from ._lib import check, lib


# ---- targets sharded, table replicated -----------------------------------------------------------
def shard_targets(lengths, world, k=31):
    """Deal targets to `world` ranks balanced by reference-k-mer count (longest first, always to the
    least-loaded rank; ties by rank then by input order, so every rank computes the same answer).
    Returns a list of index arrays, each in ascending input order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    weight = np.maximum(lengths - k + 1, 1)
    order = np.lexsort((np.arange(len(weight)), -weight))
    heap = [(0, r) for r in range(world)]
    mine = [[] for _ in range(world)]
    for i in order.tolist():
        load, r = heapq.heappop(heap)
        mine[r].append(i)
        heapq.heappush(heap, (load + int(weight[i]), r))
    return [np.array(sorted(m), dtype=np.int64) for m in mine]


def gather_in_order(local_items, local_index, n_total, dist=None, dst=0):
    """local_items[j] belongs to input position local_index[j]; rank `dst` gets the list of all
    n_total items in input order, the other ranks get None."""
    pairs = list(zip([int(i) for i in local_index], local_items))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        gathered = [pairs]
        rank = dst
    else:
        rank = dist.get_rank()
        gathered = [None] * dist.get_world_size() if rank == dst else None
        dist.gather_object(pairs, gathered, dst=dst)
    if rank != dst:
        return None
    out = [None] * n_total
    for part in gathered:
        for i, item in part:
            out[i] = item
    if any(x is None for x in out):
        raise RuntimeError("gather_in_order: some targets were not produced by any rank")
    return out


def shard_ranges(lengths, world, k=31):
    """CONTIGUOUS ranges of targets, one per rank, balanced by reference-k-mer count: rank r gets [cut[r], cut[r+1]).
    With contiguous ranges the text of the whole run is the ranks' texts one after the other -- rank 0 joins
    buffers, nothing is re-ordered per target."""
    lengths = np.asarray(lengths, dtype=np.int64)
    weight = np.maximum(lengths - k + 1, 1)
    csum = np.concatenate([[0], np.cumsum(weight)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        i = int(np.searchsorted(csum, total * r / world, side="left"))
        cuts.append(max(cuts[-1], min(i, len(lengths))))
    cuts.append(len(lengths))
    return cuts


class ShardPlan:
    """This rank's share of a batch of targets (a contiguous range) packed once, for repeated sharded runs."""

    def __init__(self, packed, world, rank, k=31):
        if not isinstance(packed, engine.PackedTargets) or packed.names is None:
            raise TypeError("ShardPlan takes PackedTargets(sequences, names)")
        self.n_total = len(packed)
        self.cuts = shard_ranges([len(s) for s in packed.sequences], world, k)
        self.lo, self.hi = self.cuts[rank], self.cuts[rank + 1]
        self.mine = packed.slice(self.lo, self.hi)
        self.world, self.rank = world, rank
        self._bufs = {}

    def buffer(self, name, n_bytes, device, pinned=False):
        """A byte buffer of at least n_bytes kept across calls (device or pinned host memory)."""
        import torch
        cur = self._bufs.get(name)
        if cur is None or cur.numel() < n_bytes or cur.device != device:
            cur = torch.empty(max(64, n_bytes + n_bytes // 4), dtype=torch.uint8, device=device,
                              pin_memory=bool(pinned and torch.cuda.is_available()))
            self._bufs[name] = cur
        return cur


def find_mutation_sharded(table, targets, db_name, dist=None, device=None, raise_errors=True, **params):
    """`km find_mutation` for a batch of targets on all ranks of `dist` (the per-target loop of
    km/tools/find_mutation.py:47-58 dealt to the GPUs): every rank holds a replica of the table (or a shard of a
    peer-mapped cohort table), runs ONE km_find_text call on its contiguous share, and rank 0 receives the ranks'
    texts and per-target statuses with two collectives (sizes, then one padded gather of byte buffers) and joins
    them in rank order = input order.  `targets`: a ShardPlan, or PackedTargets(sequences, names).

    Returns (text: uint8 array, status: uint32 array over all targets) on rank 0, (None, None) elsewhere; when `targets`
    is a ShardPlan the two arrays are views of buffers the plan keeps (valid until its next call).
    A rank whose call fails still takes part in both collectives, so nobody hangs; the error then surfaces on
    every rank (raise_errors) after the gather, like the reference's -- which prints the rows of the targets
    before the failing one first (the statuses of targets that failed on their own are returned, not raised:
    callers print the text, then engine.raise_for_status in target order)."""
    import torch
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    plan = targets if isinstance(targets, ShardPlan) else ShardPlan(targets, world, rank, table.k)
    err = None
    text = np.zeros(0, dtype=np.uint8)
    status = np.zeros(0, dtype=np.uint32)
    try:
        text, status = table.find_text(plan.mine, db_name, as_bytes=True, **params)
    except Exception as e:                         # a failing CALL (not a failing target): reported after the collectives
        err = "rank %d: %s" % (rank, e)
    if world == 1:
        if err and raise_errors:
            raise RuntimeError(err)
        return np.array(text), np.array(status)
    dev = device if device is not None else (torch.device("cuda", table.device) if dist.get_backend() == "nccl" else torch.device("cpu"))
    on_gpu = dev.type == "cuda"
    n_text, n_stat = int(text.size), int(status.size)
    sizes = torch.tensor([n_text, n_stat, 1 if err else 0], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(3 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes)
    all_sizes = all_sizes.cpu().numpy().reshape(world, 3)
    text_sizes = all_sizes[:, 0].tolist()
    stat_sizes = (4 * all_sizes[:, 1]).tolist()
    # TWO collectives move every rank's bytes to rank 0 (all_to_all_single with everything addressed to rank 0): the
    # texts, which land one after the other in rank order -- that IS the text of the whole run, nothing is joined or
    # copied again on the host -- and the (small) statuses.  The text is read where the library left it (pinned memory).
    send_t = plan.buffer("send_text", n_text, dev)
    send_s = plan.buffer("send_stat", 4 * n_stat, dev)
    if n_text:
        send_t[:n_text].copy_(torch.from_numpy(text), non_blocking=on_gpu)
    if n_stat:
        send_s[:4 * n_stat].copy_(torch.from_numpy(np.ascontiguousarray(status, dtype=np.uint32).view(np.uint8)), non_blocking=on_gpu)
    total_t = int(sum(text_sizes)) if rank == 0 else 0
    total_s = int(sum(stat_sizes)) if rank == 0 else 0
    recv_t = plan.buffer("recv_text", total_t, dev)
    recv_s = plan.buffer("recv_stat", total_s, dev)
    zeros = [0] * world
    dist.all_to_all_single(recv_t[:total_t], send_t[:n_text], text_sizes if rank == 0 else zeros, [n_text] + [0] * (world - 1))
    dist.all_to_all_single(recv_s[:total_s], send_s[:4 * n_stat], stat_sizes if rank == 0 else zeros, [4 * n_stat] + [0] * (world - 1))
    any_err = bool(all_sizes[:, 2].any())
    if any_err:
        msgs = [None] * world
        dist.all_gather_object(msgs, err)
        if raise_errors:
            raise RuntimeError("; ".join(m for m in msgs if m))
    if rank != 0:
        if on_gpu:
            torch.cuda.current_stream(dev).synchronize()      # the send buffers are reused by the next call
        return None, None
    if on_gpu:
        # one copy back each, into pinned memory that lives on the plan: the returned text is a VIEW of it, valid until the
        # next call with this plan
        host_t = plan.buffer("host_text", total_t, torch.device("cpu"), pinned=True)
        host_s = plan.buffer("host_stat", total_s, torch.device("cpu"), pinned=True)
        host_t[:total_t].copy_(recv_t[:total_t], non_blocking=True)
        host_s[:total_s].copy_(recv_s[:total_s], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return host_t[:total_t].numpy(), host_s[:total_s].numpy().view(np.uint32)
    return recv_t[:total_t].numpy(), recv_s[:total_s].numpy().view(np.uint32)


# ---- table sharded -----------------------------------------------------------------------------------
def shard_owner(kmers, k, canonical, n_shards):
    """Owner shard of each packed forward-strand k-mer (the library's own arithmetic, on the host)."""
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    out = np.empty(kmers.size, dtype=np.int32)
    check(lib().km_shard_owner(kmers.ctypes.data, kmers.size, int(k), int(bool(canonical)), int(n_shards), out.ctypes.data))
    return out


def route_queries(kmers, owners, dist, lookup_local, device="cpu"):
    """Explicit exchange: every rank holds some k-mers to look up; each goes to its owner, the owner
    answers from its shard, the counts come back in the caller's order.
        kmers    uint64 array, this rank's queries
        owners   int32 array, owner rank of each (shard_owner)
        lookup_local(keys: int64 torch tensor on `device`) -> int64/uint32 torch tensor of counts
    Three collectives: all_to_all of the per-destination sizes, of the keys, of the counts."""
    import torch
    world = dist.get_world_size()
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    order = np.argsort(owners, kind="stable")
    send_counts = np.bincount(owners, minlength=world).astype(np.int64)
    send = torch.from_numpy(kmers[order].view(np.int64)).to(device)
    sc = torch.from_numpy(send_counts).to(device)
    rc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(rc, sc)
    recv_counts = rc.cpu().tolist()
    recv = torch.empty(int(sum(recv_counts)), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, recv_counts, send_counts.tolist())
    answers = lookup_local(recv).to(torch.int64) & 0xFFFFFFFF        # counts are uint32 bit patterns
    back = torch.empty(kmers.size, dtype=torch.int64, device=device)
    dist.all_to_all_single(back, answers, send_counts.tolist(), recv_counts)
    out = np.empty(kmers.size, dtype=np.uint32)
    out[order] = back.cpu().numpy().astype(np.uint32)
    return out


class ShardedTable(engine.Table):
    """Shard `rank` of a table hash-sharded over `world` GPUs (one process per GPU)."""

    @classmethod
    def create(cls, rank, world, k=31, canonical=True, capacity_per_shard=1 << 20, device=None):
        h = ctypes.c_void_p()
        dev = rank if device is None else device
        check(lib().km_table_create_shard(int(dev), int(k), int(bool(canonical)), int(capacity_per_shard), int(rank),
                                          int(world), ctypes.byref(h)))
        t = cls(h)
        t.rank, t.world = int(rank), int(world)
        return t

    def attach(self, dist):
        """Map every other rank's shard into this process.  Each shard is exported as a POSIX file
        descriptor (CUDA virtual-memory API) and handed to the peers over a Unix socket (SCM_RIGHTS);
        the ranks take turns serving so that nobody connects before the socket listens."""
        import os
        import socket
        import time
        fd = ctypes.c_int(-1)
        check(lib().km_table_shard_export_fd(self._h, ctypes.byref(fd)))
        token = [os.getpid() if self.rank == 0 else None]
        dist.broadcast_object_list(token, src=0)
        path = lambda r: "\0km_b200_%d_%d" % (token[0], r)          # abstract namespace: nothing to unlink
        for server in range(self.world):
            if server == self.rank:
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as srv:
                    srv.bind(path(server))
                    srv.listen(self.world)
                    dist.barrier()
                    for _ in range(self.world - 1):
                        conn, _addr = srv.accept()
                        with conn:
                            socket.send_fds(conn, [b"fd"], [fd.value])
            else:
                dist.barrier()
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                    for attempt in range(200):
                        try:
                            c.connect(path(server))
                            break
                        except (FileNotFoundError, ConnectionRefusedError):
                            time.sleep(0.01)
                    _msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
                check(lib().km_table_shard_attach_fd(self._h, server, fds[0]))
            dist.barrier()
        os.close(fd.value)

    def set_routing(self, on=True):
        """Routed inserts: every key given to insert / count_* on this rank goes to its OWNER's shard by atomics over
        NVLink (all peers attached).  Each rank then feeds its own part of the stream -- its sample's reads."""
        check(lib().km_table_set_routing(self._h, int(bool(on))))

    def query_routed_device(self, kmers_dev, dist):
        """The explicit exchange with everything on the device: `kmers_dev` is a CUDA int64 tensor of packed k-mers;
        owners are computed and the batch is grouped by owner by the library's kernels (km_route_partition), the keys
        travel with one NCCL all_to_all_single, the owner answers from its shard (km_query_batch_device), the counts
        travel back and are put in the caller's order (km_route_unpermute).  Only the per-owner counts (8 numbers)
        visit the host, because NCCL wants its split sizes there.  Returns a CUDA int32 tensor (uint32 bit patterns)."""
        import torch
        dev = kmers_dev.device
        cur = torch.cuda.current_stream(dev)
        if cur.cuda_stream == 0:
            # the legacy default stream has a null handle, which the library reads as "the table's own stream": run on
            # a real stream of ours instead, ordered after and before the caller's
            if getattr(self, "_stream", None) is None:
                self._stream = torch.cuda.Stream(dev)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                out = self.query_routed_device(kmers_dev, dist)
            cur.wait_stream(self._stream)
            out.record_stream(cur)
            return out
        n = kmers_dev.numel()
        s = cur
        sp = ctypes.c_void_p(s.cuda_stream)
        sorted_k = torch.empty(n, dtype=torch.int64, device=dev)
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        scratch = torch.empty(24, dtype=torch.int64, device=dev)
        check(lib().km_route_partition(self._h, ctypes.c_void_p(kmers_dev.data_ptr()), n, ctypes.c_void_p(sorted_k.data_ptr()),
                                       ctypes.c_void_p(perm.data_ptr()), ctypes.c_void_p(scratch.data_ptr()), sp))
        send_counts_dev = scratch[:self.world].clone()
        recv_counts_dev = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv_counts_dev, send_counts_dev)
        both = torch.stack([send_counts_dev, recv_counts_dev]).cpu()          # the one host round trip: 2 x world integers
        send_counts, recv_counts = both[0].tolist(), both[1].tolist()
        recv = torch.empty(int(sum(recv_counts)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, sorted_k, recv_counts, send_counts)
        answers = torch.empty(recv.numel(), dtype=torch.int32, device=dev)
        if recv.numel():
            check(lib().km_query_batch_device(self._h, ctypes.c_void_p(recv.data_ptr()), recv.numel(),
                                              ctypes.c_void_p(answers.data_ptr()), sp))
        back = torch.empty(n, dtype=torch.int32, device=dev)
        dist.all_to_all_single(back, answers, send_counts, recv_counts)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        check(lib().km_route_unpermute(self._h, ctypes.c_void_p(back.data_ptr()), ctypes.c_void_p(perm.data_ptr()), n,
                                       ctypes.c_void_p(out.data_ptr()), sp))
        self.last_route = {"send_counts": send_counts, "recv_counts": recv_counts}
        return out

    def query_routed(self, kmers, dist):
        """Counts of `kmers` through the explicit all-to-all exchange (no peer mapping needed)."""
        import torch
        dev = torch.device("cuda", self.device)
        # one real (non-default) stream carries the collectives AND the lookup kernel, so they are ordered
        # without host synchronisation (a null stream handle would mean "the library's own stream")
        if getattr(self, "_stream", None) is None:
            self._stream = torch.cuda.Stream(dev)

        def lookup_local(keys):
            out = torch.empty(keys.numel(), dtype=torch.int32, device=dev)
            if keys.numel():
                check(lib().km_query_batch_device(self._h, ctypes.c_void_p(keys.data_ptr()), keys.numel(),
                                                  ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(self._stream.cuda_stream)))
            return out

        owners = shard_owner(kmers, self.k, self.canonical, self.world)
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            out = route_queries(kmers, owners, dist, lookup_local, device=dev)
        torch.cuda.current_stream(dev).wait_stream(self._stream)
        return out
