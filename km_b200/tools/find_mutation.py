"""`km find_mutation <target(s)> <db.jf>` -- drop-in for km/tools/find_mutation.py.

Same stdout: the `#name:value` echo of every argument, the header, one sorted block of rows
per target in target order, `#Elapsed time:`.  The per-target loop of the reference
(find_mutation.py:47-58) becomes ONE km_find_batch launch over all targets; rows are formatted
by the library.
"""
import logging as log
import os
import sys
import time

from ..utils import MutationFinder as umf
from ..utils import Sequence as us
from ..utils import common as uc
from ..utils.Jellyfish import Jellyfish


def _read_targets(target_fn, k):
    refpaths = []
    for seq_f in uc.target_2_seqfiles(target_fn):
        ref_name = os.path.splitext(os.path.basename(seq_f))[0]
        ref_seqs, _ = uc.file_2_seq(seq_f)
        refpaths.append(us.RefSeq("".join(ref_seqs), ref_name, k))    # multi-record FASTA is concatenated
    return refpaths


def _gpu_worker(rank, world, port, opts):
    """One process per GPU (--gpus N): the table is loaded on every GPU, the targets are dealt in contiguous
    shares, rank 0 prints the joined text and then raises for the first failing target, like the serial loop."""
    import torch
    import torch.distributed as dist
    from .. import cohort, engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # stdout is the report (km's callers parse it).  Whatever a library has to say on file descriptor 1 -- NCCL prints its
    # "NCCL version ..." banner there at NCCL_DEBUG=VERSION / WARN -- goes to stderr: the worker keeps a private copy of
    # the real stdout for the report and points descriptor 1 at descriptor 2.
    sys.stdout.flush()
    report = os.fdopen(os.dup(1), "wb")
    os.dup2(2, 1)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        jf = Jellyfish(opts["jellyfish_fn"], cutoff=opts["ratio"], n_cutoff=opts["count"], device=rank)
        refpaths = _read_targets(opts["target_fn"], jf.k)
        packed = engine.PackedTargets([r.seq for r in refpaths], [r.name for r in refpaths])
        text, status = cohort.find_mutation_sharded(jf.jf, packed, opts["jellyfish_fn"], dist, count=opts["count"],
                                                    ratio=opts["ratio"], steps=opts["steps"], branchs=opts["branchs"],
                                                    nodes=opts["nodes"])
        if rank == 0:
            # rows of the targets before a failing one are printed, then the error surfaces
            bad = [i for i, st in enumerate(status.tolist()) if st & ~engine.ST_TOUCHED_LIMIT]
            if bad:
                first = refpaths[bad[0]].name.encode()
                keep = []
                for ln in text.tobytes().split(b"\n"):
                    if ln and ln.split(b"\t")[1] == first:
                        break
                    if ln:
                        keep.append(ln)
                report.write(b"\n".join(keep) + (b"\n" if keep else b""))
                report.flush()
                engine.raise_for_status(status[bad[0]], refpaths[bad[0]].name, opts["nodes"])
            report.write(text.tobytes())
            report.flush()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _main_multi_gpu(args):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    opts = {k: getattr(args, k) for k in ("jellyfish_fn", "target_fn", "ratio", "count", "steps", "branchs", "nodes")}
    sys.stdout.flush()
    mp.spawn(_gpu_worker, args=(args.gpus, port, opts), nprocs=args.gpus, join=True)


def main_find_mut(args, argparser):
    time_start = time.time()
    if args.verbose:
        log.basicConfig(level=log.INFO, format="VERBOSE: %(message)s")
    if args.debug:
        log.basicConfig(level=log.DEBUG, format="VERBOSE: %(message)s")
    if getattr(args, "graphical", False):
        sys.exit("ERROR: -g/--graphical is not available in km_b200 (plots are outside the GPU path)")

    gpus = int(getattr(args, "gpus", 1) or 1)
    for name, value in vars(args).items():
        if name == "gpus" and gpus == 1:
            continue
        sys.stdout.write("#" + str(name) + ":" + str(value) + "\n")

    if gpus > 1:
        umf.MutationFinder.output_header()
        _main_multi_gpu(args)
        sys.stdout.write("#Elapsed time:" + str(time.time() - time_start) + "\n")
        return

    jf = Jellyfish(args.jellyfish_fn, cutoff=args.ratio, n_cutoff=args.count)
    umf.MutationFinder.output_header()
    refpaths = _read_targets(args.target_fn, jf.k)

    # rows of the targets before a failing one are still printed, then the error surfaces
    for finder in umf.find_batch(refpaths, jf, args.steps, args.branchs, args.nodes):
        sys.stdout.write(finder.format_rows())

    sys.stdout.write("#Elapsed time:" + str(time.time() - time_start) + "\n")
