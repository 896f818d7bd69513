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


def find_mutation_sharded(table, sequences, names, db_name, dist=None, **params):
    """`km find_mutation` for a list of targets on all ranks of `dist`: every rank holds a replica of
    the table, takes its share of the targets and formats their rows; rank 0 receives one text block
    per target in input order (what the CLI prints)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_targets([len(s) for s in sequences], world, table.k)[rank]
    res = table.find_batch([sequences[i] for i in mine], want_graph=False, **params)
    for j, i in enumerate(mine.tolist()):
        engine.raise_for_status(res.status[j], names[i], params.get("nodes", 10000))
    texts = [res.format_target(j, db_name, names[i]) for j, i in enumerate(mine.tolist())]
    return gather_in_order(texts, mine, len(sequences), dist)


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
