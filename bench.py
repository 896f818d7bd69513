#!/usr/bin/env python
"""bench.py -- km find_mutation on the synthetic panel of BASELINE.json config 4.

    python bench.py --gpus N --steps K --warmup W            (one rank per GPU under torchrun)
    python bench.py --impl reference ...                     (the CPU arm: oracle port on host cores)

A STEP = one pass of the hot path over one batch: the 10,000-target panel (planted
SNV/insertion/deletion/tandem-duplication variants) against the ~2e9-distinct-31-mer table.
Targets shard across ranks (each rank works on its own 10,000-target panel -> weak scaling), the
table is replicated, there is no collective on the data path.

  value   targets/s, whole job, inputs resident in HBM: K launches of the resident plan (eight kernels per
          launch: reference probe, walk, scheduler, two bubble passes, three graph passes; one CUDA graph), CUDA events
          on the launch stream, max over ranks
  e2e     the same through the reference-facing call with HOST buffers: km_find_text (H2D of sequences and
          names, kernels, the text `km find_mutation` prints formatted on the device, D2H of that text), wall
          clock bracketed by barriers + synchronize; its text is compared with the host formatter's every run
  roofline      reference-probe kernel: the lookups it issues x 32 B / its CUDA-event duration, against
                MEASURED_PEAKS.json hbm_gbs and against the random-gather rate measured in the same run
  lookup        2^30 device-resident canonical k-mer lookups (km_query_batch_device)
  cpu_baseline  oracle/ (the CPU restatement of the reference) on a bounded sample, all host cores
"""
import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "targets/sec"
UNIT = "targets/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed regions run."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own Python (baseline/_ref, unmodified km 2.2.2) one process per core, each running
# main_find_mut (km/tools/find_mutation.py:17-60) over a contiguous slice of the targets; Jellyfish -- absent from
# the reference tree -- is oracle/jellyfish_standin over oracle/kmer_store.c with the analytic 2e9-key background.
# kind "port" runs oracle/km_oracle.py (the CPU restatement) over the same slices instead.
# ---------------------------------------------------------------------------------------------
def _cpu_worker(conn, job, keys, counts, bg_seed, bg_n, kind, workdir):
    """One process = one core = one contiguous slice of the targets, for the life of the measurement."""
    try:
        import warnings
        warnings.simplefilter("ignore")                  # the reference's own SyntaxWarnings (Graph.py docstring)
        from oracle import km_oracle as ko
        from oracle.store import KmerStore
        store = KmerStore(31, True, len(keys))
        store.set_background(bg_seed, bg_n)
        store.insert(keys, counts)
        names, seqs = [n for n, _ in job], [s for _, s in job]
        if kind == "reference":
            from oracle import reference_runner as rr
            session = rr.ReferenceSession(store, os.path.join(workdir, "w%d" % os.getpid()))
            files = session.write_targets(names, seqs)   # untimed: the reference reads its targets from disk
        else:
            jf = ko.OracleJellyfish(store, "panel.jf", 0.05, 5)
        conn.send("ready")
        while True:
            n_take = conn.recv()
            if n_take is None:
                break
            if kind == "reference":
                rows, issued = session.find_mutation(files[:n_take]) if n_take else ([], 0)
            else:
                q0 = jf.n_queries
                rows = []
                for name, seq in job[:n_take]:
                    f = ko.OracleFinder(ko.Target(seq, name, 31), jf).run()
                    rows.extend(str(r) for r in f.get_paths())
                issued = jf.n_queries - q0
            conn.send((min(n_take, len(job)), issued, rows))
    except Exception as e:                               # surfaces in the parent instead of a hang
        import traceback
        conn.send(("error", "%s\n%s" % (e, traceback.format_exc())))


def workload_config(workload, targets, table_keys, world):
    """`config` of the JSON line: names the workload; identical on the native and the --impl reference arm."""
    return {"workload": workload, "targets_per_gpu": targets, "table_keys": table_keys,
            "parallelism": "targets sharded x%d, table replicated, no data-path collective" % world,
            "l2": "table (%.0f GB of 32-byte buckets) and per-step visited sets are far larger than the 126 MB L2; "
                  "no explicit flush" % (table_keys * 32 / 1e9)}


def cpu_arm(panel, bg_seed, bg_n, n_sample, steps, warmup, cores=None, kind="reference", budget_s=None, table_keys=None):
    """Times find_mutation on the CPU over `n_sample` targets of the panel, one process per core, each over a
    contiguous 1/P slice.  budget_s: if the warm-up predicts that `steps` passes exceed it,
    every slice shrinks.  Returns a dict: value (targets/s), cores, sample, issued_lookups_per_s, ms_per_step, rows
    (of the last step: one list per slice), names (targets of each slice that were run)."""
    import multiprocessing as mp
    import shutil
    import tempfile
    cores = cores or os.cpu_count() or 1
    n_sample = min(n_sample, len(panel.targets))
    jobs_targets = list(zip(panel.names[:n_sample], panel.targets[:n_sample]))
    per = (n_sample + cores - 1) // cores
    chunks = [jobs_targets[i:i + per] for i in range(0, n_sample, per)]
    keys, counts = table_keys if table_keys is not None else (panel.keys, panel.counts)
    workdir = tempfile.mkdtemp(prefix="km_ref_")
    ctx = mp.get_context("fork")
    procs, conns = [], []
    try:
        for c in chunks:
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_cpu_worker, args=(b, c, keys, counts, bg_seed, bg_n, kind, workdir), daemon=True)
            pr.start()
            procs.append(pr)
            conns.append(a)

        def gather():
            out = [c.recv() for c in conns]
            for o in out:
                if o and o[0] == "error":
                    raise RuntimeError("CPU arm worker failed: " + o[1])
            return out
        gather()                                          # "ready"

        def step(takes):
            t0 = time.perf_counter()
            for c, n in zip(conns, takes):
                c.send(n)
            out = gather()
            return time.perf_counter() - t0, out
        takes = [len(c) for c in chunks]
        probe = [min(8, t) for t in takes]
        warm = 0.0
        for _ in range(max(1, warmup)):
            warm, _o = step(probe)
        if budget_s is not None:
            predicted = warm * (max(takes) / max(1, max(probe))) * steps
            if predicted > budget_s:
                scale = budget_s / predicted
                takes = [max(8, int(t * scale)) for t in takes]
        times, issued, rows, done = [], 0, [], 0
        for _ in range(steps):
            dt, out = step(takes)
            times.append(dt)
            issued = sum(o[1] for o in out)
            done = sum(o[0] for o in out)
            rows = [o[2] for o in out]
        for c in conns:
            c.send(None)
        for pr in procs:
            pr.join(timeout=10)
    finally:
        for pr in procs:
            if pr.is_alive():
                pr.terminate()
        shutil.rmtree(workdir, ignore_errors=True)
    total = sum(times)
    what = ("unmodified reference Python (km 2.2.2 main_find_mut, baseline/_ref) + stand-in k-mer store"
            if kind == "reference" else "oracle port (Python restatement + C k-mer store)")
    sample = "%d of %d panel targets per step, %d processes (contiguous slices), %s, analytic %d-key background" % (
        done, len(panel.targets), len(chunks), what, bg_n)
    return {"value": done * steps / total, "cores": len(chunks), "sample": sample, "issued_lookups_per_s": issued * steps / total,
            "ms_per_step": 1e3 * total / steps, "rows": rows, "names": [[n for n, _ in c[:t]] for c, t in zip(chunks, takes)],
            "n_done": done, "kind": kind}


def reference_available():
    from oracle import reference_runner as rr
    return rr.locate_reference() is not None


def compare_with_cpu_rows(cpu, text):
    """Rows the CPU arm produced (per slice, per target) against the GPU arm's text: every target the CPU ran."""
    from oracle.compare import compare_rows
    by_target = {}
    for ln in text.split("\n"):
        if ln:
            by_target.setdefault(ln.split("\t")[1], []).append(ln)
    bad, flips, checked, bad_names = 0, 0, 0, []
    for rows, names in zip(cpu["rows"], cpu["names"]):
        mine = {}
        for r in rows:
            mine.setdefault(r.split("\t")[1], []).append(r)
        for n in names:
            errs, fl = compare_rows(mine.get(n, []), by_target.get(n, []))
            checked += 1
            flips += fl
            if errs:
                bad += 1
                if len(bad_names) < 8:
                    bad_names.append(n)
    return {"targets_checked": checked, "mismatching": bad, "printed_digit_flips": flips, "mismatching_names": bad_names,
            "against": cpu["kind"]}


# ---------------------------------------------------------------------------------------------
# N > 1 only: strong scaling through the product's sharded call, and cohort mode (BASELINE.json config 5)
# ---------------------------------------------------------------------------------------------
def strong_leg(args, table, dist, dev, rank, world, panel, text_rank0, barrier, max_over_ranks, stream):
    """The SAME 10,000-target panel (rank 0's) dealt over the N ranks in contiguous shares, through
    cohort.find_mutation_sharded: one km_find_text per rank (host buffers in), the texts gathered on rank 0
    -- gather included in the timed region.  Also the device-only time of each rank's share."""
    import torch
    from km_b200 import cohort, engine, synth
    box = [(panel.targets, panel.names)] if rank == 0 else [None]
    dist.broadcast_object_list(box, src=0)
    targets0, names0 = box[0]
    packed0 = engine.PackedTargets(targets0, names0)
    plan = cohort.ShardPlan(packed0, world, rank, 31)
    for _ in range(max(3, args.warmup)):
        text, status = cohort.find_mutation_sharded(table, plan, "panel.jf", dist)
    gc.collect()
    gc.disable()                # (as for the e2e loop of main(): no generation-2 collection inside a region of a few ms)
    barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        text, status = cohort.find_mutation_sharded(table, plan, "panel.jf", dist)
    torch.cuda.synchronize(dev)
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
    gc.enable()
    same = bool(np.array_equal(text, text_rank0)) if rank == 0 else None
    # device-only: every rank's share resident in HBM, K launches, CUDA events, max over ranks
    share = table.plan(plan.mine.sequences)
    for _ in range(3):
        share.launch(stream.cuda_stream)
    torch.cuda.synchronize(dev)
    share.fetch(want_graph=False)
    share.launch(stream.cuda_stream)
    barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        share.launch(stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    km = [max_over_ranks(x) for x in share.kernel_ms()]
    share.close()
    n = len(targets0)
    worst = max(range(3), key=lambda i: km[i])
    return {"what": "the same %d-target panel dealt over %d ranks in contiguous shares (cohort.find_mutation_sharded: one "
                    "km_find_text per rank, texts gathered on rank 0 with 2 collectives)" % (n, world),
            "targets": n, "e2e_ms_per_step": e2e_ms, "e2e_value": n / (e2e_ms / 1e3), "device_ms_per_step": dev_ms,
            "device_value": n / (dev_ms / 1e3), "unit": UNIT, "text_equals_single_gpu_text": same,
            "kernel_ms_max_over_ranks": {"ref_probe": km[0], "walks": km[1], "graph": km[2]},
            "limited_by": ("e2e: per-call host work (enqueue, staging, the gather's two collectives) on a share of %d targets; "
                           "device: the %s, whose tail is a handful of long dependent chains per share"
                           % (plan.hi - plan.lo, ("reference-probe kernel", "walk kernels", "graph kernels")[worst]))}


def cohort_leg(args, table, dist, dev, rank, world, local, panel, table_keys, barrier, max_over_ranks, sum_over_ranks, stream):
    """BASELINE.json config 5: N synthetic samples (one per rank) -> reads -> canonical 31-mers counted ON THE DEVICE
    into ONE table hash-sharded over the N GPUs, every k-mer sent to its owner by atomics over NVLink from inside the
    counting kernel (`jellyfish count -m 31 -C`), counts < 2 dropped (-L 2); then a ~2e9-key synthetic background is
    added (the 'huge table'), and the table is queried through (a) peer loads inside the probe kernel, (b) the explicit
    exchange with device-side routing + NCCL all-to-all, (c) the panel's find_mutation; each checked."""
    import torch
    from km_b200 import cohort, engine, synth
    from km_b200._lib import lib, check
    n_t = min(args.cohort_targets, len(panel.targets))
    sub = synth.make_panel(n_t, seed=synth.PANEL_SEED + rank + args.panel_offset)        # the first n_t targets of this rank's panel
    assert sub.targets == panel.targets[:n_t]
    t0 = time.perf_counter()
    reads = synth.sample_reads(sub)
    gen_s = time.perf_counter() - t0
    n_reads = reads.count(b"\n")
    n_kmers = len(reads) - 31 * n_reads
    counted_keys_total = int(sum_over_ranks(len(sub.keys)))
    cap = (table_keys + 2 * counted_keys_total) // world + (1 << 20)
    shard = cohort.ShardedTable.create(rank, world, capacity_per_shard=cap, device=local)
    shard.attach(dist)
    shard.set_routing(True)
    barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    shard.count_text(reads)                          # host buffer in: H2D + count kernel with routed inserts, chunks overlapped
    torch.cuda.synchronize(dev)
    barrier()
    count_s = max_over_ranks(time.perf_counter() - t0)
    kmers_total = sum_over_ranks(n_kmers)
    bytes_total = sum_over_ranks(len(reads))
    shard.drop_below(2)                              # -L 2 (also recounts this shard's keys)
    barrier()
    shard.set_routing(False)
    kept_total = int(sum_over_ranks(shard.info()["n_keys"]))
    # the sample's analytic content (what the reads must have produced) through peer loads -- every rank, every key
    got = shard.query_packed(sub.keys)
    analytic_ok = bool(sum_over_ranks(int(not (got == sub.counts).all())) == 0)
    # the huge-table filler: the config-4 background, every rank keeps what it owns (insert-if-absent)
    t0 = time.perf_counter()
    shard.build_synthetic(synth.TABLE_SEED, table_keys)
    barrier()
    fill_s = time.perf_counter() - t0
    keys_total = int(sum_over_ranks(shard.info()["n_keys"]))
    # host recount of the actual read bytes of a few targets (rank 0): the counting kernel against a plain numpy count
    recount = None
    if rank == 0:
        from oracle import count_oracle
        some = list(range(min(6, n_t)))
        hk, hc = count_oracle.count_stream(synth.sample_reads(sub, some))
        keep = hc >= 2
        gk = shard.query_packed(hk)
        # (keys below 2 were dropped; a background key may sit there instead -- decided by the analytic background)
        from oracle.store import KmerStore
        st = KmerStore(31, True, 16)
        st.set_background(synth.TABLE_SEED, table_keys)
        want = np.where(keep, hc, st.query_batch(hk).astype(np.int64))
        recount = {"targets": len(some), "kmers_distinct": int(len(hk)), "equal": bool((gk.astype(np.int64) == want).all())}
    # ---- (a) lookups through peer loads, (b) through the device-side exchange --------------------------------------
    nq = args.cohort_queries
    q = torch.empty(nq, dtype=torch.int64, device=dev)
    check(lib().km_bench_make_queries(shard._h, ctypes.c_void_p(q.data_ptr()), nq, synth.TABLE_SEED, table_keys,
                                      synth.QUERY_SEED + 17 * rank, ctypes.c_void_p(stream.cuda_stream)))
    out_peer = torch.empty(nq, dtype=torch.int32, device=dev)
    out_repl = torch.empty(nq, dtype=torch.int32, device=dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def peer():
        check(lib().km_query_batch_device(shard._h, ctypes.c_void_p(q.data_ptr()), nq, ctypes.c_void_p(out_peer.data_ptr()), sp))
    for _ in range(3):
        peer()
    barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        peer()
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    peer_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    # the replicated table holds the same background (+ planted keys the query mix never asks for): equal answers expected
    check(lib().km_query_batch_device(table._h, ctypes.c_void_p(q.data_ptr()), nq, ctypes.c_void_p(out_repl.data_ptr()), sp))
    torch.cuda.synchronize(dev)
    equals_unsharded = bool(sum_over_ranks(int(not torch.equal(out_peer, out_repl))) == 0)
    hit_frac = float((out_peer != 0).sum().item()) / nq
    with torch.cuda.stream(stream):
        for _ in range(2):
            out_a2a = shard.query_routed_device(q, dist)
        barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_a2a = shard.query_routed_device(q, dist)
        torch.cuda.synchronize(dev)
        barrier()
    a2a_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
    a2a_equal = bool(sum_over_ranks(int(not torch.equal(out_a2a, out_peer))) == 0)
    del q, out_peer, out_repl, out_a2a
    # ---- (c) the panel against the sharded table ------------------------------------------------------------------
    packed = engine.PackedTargets(sub.targets, sub.names)
    text_sh, _ = shard.find_text(packed, "cohort.jf", as_bytes=True)
    text_rep, _ = table.find_text(packed, "cohort.jf", as_bytes=True)
    text_equal = bool(sum_over_ranks(int(not np.array_equal(text_sh, text_rep))) == 0)
    plan = shard.plan(sub.targets)
    for _ in range(3):
        plan.launch(stream.cuda_stream)
    torch.cuda.synchronize(dev)
    plan.fetch(want_graph=False)
    plan.launch(stream.cuda_stream)
    barrier()
    torch.cuda.synchronize(dev)
    ev0.record(stream)
    for _ in range(args.steps):
        plan.launch(stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    panel_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    km = [max_over_ranks(x) for x in plan.kernel_ms()]
    plan.close()
    oracle = None
    if rank == 0:
        from oracle import km_oracle as ko
        from oracle.compare import compare_rows
        from oracle.store import KmerStore
        st = KmerStore(31, True, len(sub.keys))
        st.set_background(synth.TABLE_SEED, table_keys)
        st.insert(sub.keys, sub.counts)
        jf = ko.OracleJellyfish(st, "cohort.jf", 0.05, 5)
        by_target = {}
        for ln in text_sh.tobytes().decode().split("\n"):
            if ln:
                by_target.setdefault(ln.split("\t")[1], []).append(ln)
        bad = 0
        picks = list(range(0, n_t, max(1, n_t // 48)))
        for i in picks:
            f = ko.OracleFinder(ko.Target(sub.targets[i], sub.names[i], 31), jf).run()
            errs, _ = compare_rows([str(r) for r in f.get_paths()], by_target.get(sub.names[i], []))
            bad += 1 if errs else 0
        oracle = {"targets_checked": len(picks), "mismatching": bad}
    barrier()
    shard.close()
    remote = (world - 1) / world
    return {
        "what": "config 5: %d samples' reads counted on device into one table hash-sharded over %d GPUs (owner = top bits of the "
                "key's hash), k-mers routed to their owner by system-scope atomics over NVLink inside the counting kernel; "
                "+ %.3g-key synthetic background" % (world, world, table_keys),
        "samples": world, "targets_per_sample": n_t, "reads_per_sample": n_reads, "read_bytes_total": int(bytes_total),
        "read_generation_s_host": gen_s,
        "count": {"kmers_total": int(kmers_total), "seconds": count_s, "kmers_per_s_whole_job": kmers_total / count_s,
                  "read_GBps_whole_job": bytes_total / count_s / 1e9,
                  "what": "km_table_count_text from HOST buffers on every rank at once: H2D + km_count_text_kernel, inserts routed to the owner shard",
                  "distinct_keys_kept_after_L2": kept_total, "counts_equal_analytic_model": analytic_ok, "host_recount_of_read_bytes": recount},
        "background_fill_s": fill_s, "table_keys_total": keys_total,
        "peer_loads": {"queries_per_rank": nq, "ms": peer_ms, "lookups_per_s_whole_job": world * nq / (peer_ms / 1e3),
                       "lookups_per_s_per_gpu": nq / (peer_ms / 1e3), "hit_frac": hit_frac, "equals_unsharded": equals_unsharded,
                       "nvlink_bytes_per_lookup": 32 * remote, "remote_fraction": remote,
                       "peer_gather_ceiling_per_gpu": 6.6e9,
                       "frac_of_peer_gather_ceiling": (nq * remote / (peer_ms / 1e3)) / 6.6e9 if remote else None},
        "all_to_all": {"queries_per_rank": nq, "ms": a2a_ms, "lookups_per_s_whole_job": world * nq / (a2a_ms / 1e3),
                       "equals_peer_loads": a2a_equal, "nvlink_bytes_per_lookup": 12 * remote,
                       "what": "km_route_partition (owner + grouping on the device) -> NCCL all_to_all_single of keys -> "
                               "km_query_batch_device at the owner -> all_to_all_single of counts -> km_route_unpermute; only the "
                               "per-owner counts visit the host"},
        "panel": {"targets_per_rank": n_t, "ms_per_step": panel_ms, "targets_per_s_whole_job": world * n_t / (panel_ms / 1e3),
                  "kernel_ms_max_over_ranks": {"ref_probe": km[0], "walks": km[1], "graph": km[2]},
                  "text_equals_replicated_table": text_equal, "equals_oracle": oracle},
    }


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--targets", type=int, default=10000)
    ap.add_argument("--table-keys", type=int, default=2_000_000_000)
    ap.add_argument("--lookup-queries", type=int, default=1 << 30)
    ap.add_argument("--cpu-sample", type=int, default=0, help="targets in the CPU sample (0 = the whole panel, shrunk if it would take minutes)")
    ap.add_argument("--cpu-kind", default="reference", choices=["reference", "port"],
                    help="CPU arm: the unmodified reference Python from baseline/_ref (default) or the oracle port")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookup", action="store_true")
    ap.add_argument("--no-tier2", action="store_true")
    ap.add_argument("--parts", type=int, default=0,
                    help="`value`: the resident panel is launched as this many plans on as many streams, so that the phases of one "
                         "part overlap those of the others (0 = try 1 and 2, report the better and say which)")
    ap.add_argument("--no-cohort", action="store_true", help="skip the cohort (config 5) and strong-scaling legs at N > 1")
    ap.add_argument("--cohort-targets", type=int, default=1000, help="targets per sample whose reads are counted in the cohort leg")
    ap.add_argument("--cohort-queries", type=int, default=1 << 26)
    ap.add_argument("--n-sub", type=int, default=0, help="sub-batches in flight in km_find_text (0 = library default)")
    ap.add_argument("--panel-offset", type=int, default=0, help="debug: use the panel rank R would get (seed offset)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3                       # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    from km_b200 import synth
    workload = "synthetic %d-target panel (SNV/ins/del/dup planted) x %.3g-distinct-31-mer table" % (
        args.targets, args.table_keys)

    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import store
        store.build()
        try:
            from tools import stage_reference
            stage_reference.stage()
        except Exception:
            pass
        panel = synth.make_panel(args.targets, seed=synth.PANEL_SEED)
        cores = os.cpu_count() or 1
        kind = "reference" if reference_available() and args.cpu_kind != "port" else "port"
        # every step = the whole panel unless the warm-up predicts that K steps would take more than ~2.5 minutes
        r = cpu_arm(panel, synth.TABLE_SEED, args.table_keys, args.cpu_sample or args.targets, max(1, args.steps),
                    min(max(args.warmup, 1), 2), cores, kind=kind, budget_s=150.0)
        v = r["value"]
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / u32 counts / f32 graph weights / f64 solver",
            "data": "synthetic", "config": workload_config(workload, args.targets, args.table_keys, max(1, args.gpus)),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": kind, "sample": r["sample"],
                             "issued_lookups_per_s": r["issued_lookups_per_s"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---- native arm ---------------------------------------------------------------------------
    # the bench prints ONE line on stdout: everything else that writes to fd 1 (NCCL's version banner,
    # library chatter) goes to stderr; the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import __graft_entry__ as ge
    ge.build()
    from km_b200 import engine
    from km_b200._lib import lib, check

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # every rank has its own panel; the planted k-mers of ALL panels go into every table (replicated)
    panel = synth.make_panel(args.targets, seed=synth.PANEL_SEED + rank + args.panel_offset)
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, (panel.keys, panel.counts))
        all_keys = np.concatenate([g[0] for g in gathered])
        all_counts = np.concatenate([g[1] for g in gathered])
    else:
        all_keys, all_counts = panel.keys, panel.counts

    free, total_mem = torch.cuda.mem_get_info(dev)
    table_keys = args.table_keys
    per_key = 52 if os.environ.get("KM_TABLE_LINES", "0") not in ("", "0") else 32      # family lines store every key twice
    need = (table_keys + len(all_keys)) * per_key + (12 << 30) + (0 if args.no_lookup else args.lookup_queries * 12)
    if need > free:
        table_keys = max(1 << 20, int((free - (16 << 30)) // per_key // 2))
        args.lookup_queries = min(args.lookup_queries, 1 << 28)
    t_build = time.time()
    table = engine.Table.create(k=31, canonical=True, capacity=table_keys + len(all_keys), device=local)
    table.build_synthetic(synth.TABLE_SEED, table_keys)
    table.insert(all_keys, all_counts, mode="overwrite")
    t_build = time.time() - t_build
    info = table.info()

    plan = table.plan(panel.targets)
    stream = torch.cuda.Stream(dev)           # a real (non-default) stream: its handle is passed to the library
    torch.cuda.set_stream(stream)
    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: K launches of the resident batch ------------------------------------------------
    for _ in range(args.warmup):
        plan.launch(stream.cuda_stream)
    torch.cuda.synchronize(dev)
    first = plan.fetch(want_graph=False)          # also settles capacities (retries happen here, untimed)
    retries = first.timing["retries"]
    n_rows = int(first.row_count.sum())
    n_nodes = first.n_nodes.astype(np.int64)
    n_ref = np.array([max(0, len(s) - 30) for s in panel.targets], dtype=np.int64)
    ok = first.status & ~np.uint32(16) == 0
    algorithmic = int((n_ref + 4 * np.maximum(n_nodes - 2, 0))[ok].sum())
    issued = int(first.lookups.sum())
    for _ in range(2):
        plan.launch(stream.cuda_stream)

    def time_parts(n_parts):
        """K steps of the whole resident panel, launched as n_parts plans (contiguous shares balanced by reference k-mers)
        on n_parts streams: a step starts when the previous step's parts have all finished (event join), CUDA events on
        the launch stream, max over ranks."""
        if n_parts == 1:
            parts, streams = [plan], [stream]
        else:
            from km_b200 import cohort
            cuts = cohort.shard_ranges([len(s) for s in panel.targets], n_parts, 31)
            parts = [table.plan(panel.targets[a:b]) for a, b in zip(cuts, cuts[1:])]
            streams = [stream] + [torch.cuda.Stream(dev) for _ in range(n_parts - 1)]
            for p_, s_ in zip(parts, streams):
                for _ in range(3):
                    p_.launch(s_.cuda_stream)
            torch.cuda.synchronize(dev)
            for p_ in parts:
                p_.fetch(want_graph=False)           # settles capacities (untimed)

        def one_step():
            for s_ in streams[1:]:
                s_.wait_stream(stream)
            for p_, s_ in zip(parts, streams):
                p_.launch(s_.cuda_stream)
            for s_ in streams[1:]:
                stream.wait_stream(s_)
        for _ in range(3):
            one_step()
        barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            one_step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        if n_parts > 1:
            for p_ in parts:
                p_.close()
        return ms
    tried = {}
    for n_parts in ([1, 2] if args.parts <= 0 else [args.parts]):
        tried[n_parts] = time_parts(n_parts)
    best_parts = min(tried, key=tried.get)
    ms_per_step = tried[best_parts]
    value = world * args.targets / (ms_per_step / 1e3)
    # per-kernel durations, averaged over K more launches (CUDA events inside the library, on the launch stream)
    probe_ms, walk_ms, graph_ms = [], [], []
    for _ in range(args.steps):
        plan.launch(stream.cuda_stream)
        a, b, c = plan.kernel_ms()
        probe_ms.append(a)
        walk_ms.append(b)
        graph_ms.append(c)
    probe_ms, walk_ms, graph_ms = float(np.mean(probe_ms)), float(np.mean(walk_ms)), float(np.mean(graph_ms))
    # lookups the reference-probe kernel issues per launch: per reference k-mer its own count + the three
    # successors off the reference (all four for a target's last k-mer); the successor along the reference
    # comes from the neighbouring lane, except for the last lane of a 32-k-mer chunk, which fetches it
    L = n_ref[n_ref > 0]
    probe_lookups = int((L + 3 * (L - 1) + 4 + (L - 1) // 32).sum())

    # ---- e2e: host buffers in, TSV text out ---------------------------------------------------
    e2e_split = {"find_batch_ms": 0.0, "format_ms": 0.0}

    # the step's inputs as HOST buffers (sequences + offsets, names + offsets): what the C ABI takes
    packed = engine.PackedTargets(panel.targets, panel.names)

    def e2e_step():
        # ONE library call: H2D of the sequences, kernels, D2H of rows and spelled paths, text building;
        # sub-batches in flight so that these overlap (km_find_text)
        return table.find_text(packed, "panel.jf", as_bytes=True, n_sub=args.n_sub)

    def two_step():
        t_a = time.perf_counter()
        res = table.find_batch(packed, want_graph=False)
        t_b = time.perf_counter()
        text = res.format_all("panel.jf", packed, as_bytes=True)
        t_c = time.perf_counter()
        e2e_split["find_batch_ms"] += 1e3 * (t_b - t_a)
        e2e_split["format_ms"] += 1e3 * (t_c - t_b)
        return res, text

    for _ in range(args.warmup):
        text, status = e2e_step()
    # wall-clock region of a few milliseconds in a process that holds the panel, the parity rows and the CPU arm's results as
    # Python objects: a generation-2 pass of the cyclic collector (10-30 ms, tools/e2e_jitter.py saw one in 900 calls)
    # would be most of it -- collect now, and not inside
    gc.collect()
    gc.disable()
    barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        text, status = e2e_step()
    torch.cuda.synchronize(dev)
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
    gc.enable()
    e2e_value = world * args.targets / (e2e_ms / 1e3)
    h2d, d2h = table.last_timing["h2d_bytes"], table.last_timing["d2h_bytes"]
    # the same work as two calls (km_find_batch, then km_result_text), for the split between them
    for _ in range(2):
        res, text2 = two_step()
    e2e_split = {k: 0.0 for k in e2e_split}
    for _ in range(args.steps):
        res, text2 = two_step()
    e2e_breakdown = {k: res.timing[k] for k in ("h2d_ms", "walk_ms", "graph_ms", "d2h_ms")}
    e2e_breakdown.update({k: v / args.steps for k, v in e2e_split.items()})
    e2e_breakdown["text_identical_to_one_call"] = bool(np.array_equal(text, text2))
    text_bytes_rank0 = np.array(text)
    text = text.tobytes().decode()

    # ---- tier 2: the same panel against a REDUCED background (the first 1e7 keys of the same
    # stream, so full is a superset).  A target's text may differ only if its walk touched a background key the
    # reduced table lacks (chance ~ queried k-mers x 2e9 / 2^61 per panel); such targets are counted and listed.
    tier2 = None
    if rank == 0 and not args.no_tier2:
        n_small = min(10_000_000, table_keys)
        small = engine.Table.create(k=31, canonical=True, capacity=n_small + len(all_keys), device=local)
        small.build_synthetic(synth.TABLE_SEED, n_small)
        small.insert(all_keys, all_counts, mode="overwrite")
        text_small, _ = small.find_text(packed, "panel.jf")
        small.close()

        def per_target(tx):
            d = {}
            for ln in tx.split("\n"):
                if ln:
                    d.setdefault(ln.split("\t")[1], []).append(ln)
            return d
        full_rows, small_rows = per_target(text), per_target(text_small)
        differing = [n for n in panel.names if full_rows.get(n) != small_rows.get(n)]
        tier2 = {"reduced_background_keys": n_small, "full_background_keys": table_keys,
                 "targets_compared": len(panel.names), "targets_differing": len(differing), "names": differing[:16]}

    # ---- lookup microbenchmark (device-resident queries) ---------------------------------------
    lookup = None
    gather = None
    if not args.no_lookup:
        best, mean, hits = ctypes.c_float(), ctypes.c_float(), ctypes.c_uint64()
        check(lib().km_bench_lookup(table._h, synth.TABLE_SEED, table_keys, args.lookup_queries, synth.QUERY_SEED, 5,
                                    ctypes.byref(best), ctypes.byref(mean), ctypes.byref(hits)))
        lookup = {"lookups_per_s": world * args.lookup_queries / (max_over_ranks(mean.value) / 1e3),
                  "n_queries": args.lookup_queries, "mean_ms": mean.value, "best_ms": best.value,
                  "hit_frac": hits.value / args.lookup_queries,
                  "mix": "50% table keys (random strand) / 50% random 31-mers, submitted non-canonical"}
        table_bytes = info["bytes"]
        if rank == 0:
            ms = ctypes.c_float()
            free2, _ = torch.cuda.mem_get_info(dev)
            span = min(table_bytes, max(1 << 30, free2 - (4 << 30)))
            check(lib().km_bench_random_gather(local, span, 1 << 28, 3, ctypes.byref(ms)))
            gather = {"gsectors_per_s": (1 << 28) / ms.value / 1e6, "GBps": (1 << 28) * 32 / ms.value / 1e6,
                      "span_gb": span / 1e9}
    strong = cohort_out = None
    if world > 1 and not args.no_cohort:
        strong = strong_leg(args, table, dist, dev, rank, world, panel, text_bytes_rank0, barrier, max_over_ranks, stream)
        cohort_out = cohort_leg(args, table, dist, dev, rank, world, local, panel, table_keys, barrier, max_over_ranks,
                                sum_over_ranks, stream)
    clocks = sampler.stop()

    # ---- parity against the CPU arm (not timed): EVERY target of rank 0's panel --------------------------------
    # N=1: the rows are those of the cpu_baseline run itself (the unmodified reference Python when it is staged);
    # N>1: the oracle port over the whole panel on rank 0's host cores.  The 2e9-key background is decided
    # analytically on the CPU, so the comparison is exact at full scale (tier 1).
    parity = None
    cpu = None
    if rank == 0:
        cores = os.cpu_count() or 1
        kind = "reference" if reference_available() and args.cpu_kind != "port" else "port"
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_arm(panel, synth.TABLE_SEED, table_keys, args.cpu_sample or args.targets, 1, 1, cores, kind=kind,
                        budget_s=30.0, table_keys=(all_keys, all_counts))
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind, "sample": r["sample"],
                   "issued_lookups_per_s": r["issued_lookups_per_s"]}
            parity = compare_with_cpu_rows(r, text)
            if kind == "reference":        # the oracle port beside it: a second, labelled figure
                r2 = cpu_arm(panel, synth.TABLE_SEED, table_keys, args.cpu_sample or args.targets, 1, 1, cores, kind="port",
                             budget_s=30.0, table_keys=(all_keys, all_counts))
                cpu["oracle_port_value"] = r2["value"]
                cpu["oracle_port_parity"] = {k: v for k, v in compare_with_cpu_rows(r2, text).items() if k != "against"}
        elif not args.no_cpu_baseline:
            r = cpu_arm(panel, synth.TABLE_SEED, table_keys, args.targets, 1, 1, max(1, cores // 2), kind="port",
                        budget_s=20.0, table_keys=(all_keys, all_counts))
            parity = compare_with_cpu_rows(r, text)
        if parity is not None and tier2 is not None:
            parity["tier2_full_vs_reduced_background"] = tier2

    peak, peak_src = load_peaks()
    achieved = probe_lookups * 32 / (probe_ms / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("targets") == args.targets and tj.get("table_keys") == table_keys:
            traffic = tj["km_ref_probe_kernel"]["dram_bytes_per_launch"]
    except Exception:
        pass
    issued_total = sum_over_ranks(issued)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 keys / u32 counts / f32 graph weights / f64 solver",
            "data": "synthetic",
            # `config` names the workload and is the same object on both arms (--impl reference prints it too); what this run
            # measured about it sits in `config_measured`
            "config": workload_config(workload, args.targets, table_keys, world),
            "config_measured": {"table_distinct": info["n_keys"], "table_gb": info["bytes"] / 1e9, "table_build_s": t_build,
                       "table_layout": "family lines (128 B, two copies per k-mer)" if info["layout"] else "sector buckets (32 B)",
                       "ref_kmers": int(n_ref.sum()), "rows": n_rows, "capacity_retries": retries,
                       "launch_parts": best_parts,
                       "ms_per_step_by_parts": {str(k): v for k, v in tried.items()}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "device_ms": e2e_breakdown,
                    "what": "km_find_text(host buffers: sequences + offsets, names) -> the TSV text km find_mutation prints; "
                            "device_ms = the same work as two calls (km_find_batch, km_result_text), not pipelined"},
            # per step and part: reference probe, walk, scheduler, the two bubble passes, the three CTA-per-target graph
            # passes (plan_launch; replayed as one CUDA graph from a plan's third launch on)
            "gpu_launches": (6 if os.environ.get("KM_NO_BUBBLE_KERNEL", "0") not in ("", "0") else 8) * args.steps * best_parts,
            "kernels": {"km_ref_probe_kernel_ms": probe_ms, "km_walk_kernels_ms": walk_ms, "km_graph_kernels_ms": graph_ms,
                        "what": "reference probe (HBM-bound: ~87% of the panel's lookups), shared-memory + general walk "
                                "(latency-bound tails), graph/paths/quantification (shared memory, latency-bound)"},
            "roofline": {"kernel": "km_ref_probe_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "units_per_launch": probe_lookups, "bytes_per_unit": 32,
                         "unit_is": "one canonical k-mer lookup ANSWERED = one 32-byte HBM sector (SURVEY.md 8d): per reference "
                                    "k-mer its own count + the successors off the reference; with the table's neighbour masks "
                                    "most 'absent' answers need no read (issued_table_reads)",
                         "traffic_over_algorithmic": (traffic / (probe_lookups * 32.0)) if traffic else None,
                         "random_gather_GBps": gather["GBps"] if gather else None,
                         "frac_of_random_gather": (achieved / gather["GBps"]) if gather else None,
                         "panel": {"algorithmic_lookups": algorithmic, "issued_lookups": issued,
                                   "GBps_over_whole_step": algorithmic * 32 / (ms_per_step / 1e3) / 1e9,
                                   "frac_of_random_gather_over_whole_step":
                                       (algorithmic * 32 / (ms_per_step / 1e3) / 1e9 / gather["GBps"]) if gather else None}},
            "lookup": lookup, "random_gather": gather, "strong": strong, "cohort": cohort_out,
            "lookups_per_s_in_panel": issued_total / (ms_per_step / 1e3),
            "clocks": clocks, "parity": parity, "cpu_baseline": cpu,
        }
        if lookup and gather:
            lookup["sector_GBps"] = lookup["lookups_per_s"] / world * 32 / 1e9
            lookup["frac_of_random_gather"] = lookup["sector_GBps"] / gather["GBps"]
            lookup["frac_of_hbm_peak"] = lookup["sector_GBps"] / peak
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
