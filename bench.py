#!/usr/bin/env python
"""bench.py -- km find_mutation on the synthetic panel of BASELINE.json config 4.

    python bench.py --gpus N --steps K --warmup W            (one rank per GPU under torchrun)
    python bench.py --impl reference ...                     (the CPU arm: oracle port on host cores)

A STEP = one pass of the hot path over one batch: the 10,000-target panel (planted
SNV/insertion/deletion/tandem-duplication variants) against the ~2e9-distinct-31-mer table.
Targets shard across ranks (each rank works on its own 10,000-target panel -> weak scaling), the
table is replicated, there is no collective on the data path.

  value   targets/s, whole job, inputs resident in HBM: K launches of the resident plan (seven kernels per
          launch: reference probe, two walks, schedule, three graph passes; km_find_plan_launch), CUDA events
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
# CPU arm: the oracle port, one process per core
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(keys, counts, bg_seed, bg_n):
    from oracle import km_oracle as ko
    from oracle.store import KmerStore
    store = KmerStore(31, True, len(keys))
    store.set_background(bg_seed, bg_n)
    store.insert(keys, counts)
    _W["jf"] = ko.OracleJellyfish(store, "panel.jf", 0.05, 5)
    _W["ko"] = ko


def _cpu_work(job):
    ko, jf = _W["ko"], _W["jf"]
    q0 = jf.n_queries
    nodes = 0
    text = []
    for name, seq in job:
        f = ko.OracleFinder(ko.Target(seq, name, 31), jf).run()
        text.extend(str(r) for r in f.get_paths())
        nodes += len(f.refpath.ref_mer) + 4 * (f.num_k - 2)
    return len(job), jf.n_queries - q0, nodes, len(text)


def cpu_arm(panel, bg_seed, bg_n, n_sample, steps, warmup, cores=None):
    """Times the oracle port over `n_sample` targets of the panel with one process per core.
    Returns (targets/s, cores, sample text, issued lookups/s, ms per step)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n_sample = min(n_sample, len(panel.targets))
    jobs_targets = list(zip(panel.names[:n_sample], panel.targets[:n_sample]))
    chunks = [jobs_targets[i::cores] for i in range(cores)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks), initializer=_cpu_init, initargs=(panel.keys, panel.counts, bg_seed, bg_n)) as pool:
        for _ in range(warmup):
            pool.map(_cpu_work, [c[:1] for c in chunks])
        times, issued = [], 0
        for _ in range(steps):
            t0 = time.perf_counter()
            out = pool.map(_cpu_work, chunks)
            times.append(time.perf_counter() - t0)
            issued = sum(o[1] for o in out)
    total = sum(times)
    sample = "%d of %d panel targets, %d processes, oracle port (Python + C k-mer store, analytic %d-key background)" % (
        n_sample, len(panel.targets), len(chunks), bg_n)
    return n_sample * steps / total, len(chunks), sample, issued * steps / total, 1e3 * total / steps


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
    ap.add_argument("--cpu-sample", type=int, default=0, help="targets in the CPU sample (0 = 640 per core, at most the panel: ~5-10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookup", action="store_true")
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
        import __graft_entry__ as ge
        from oracle import store
        store.build()
        panel = synth.make_panel(args.targets, seed=synth.PANEL_SEED)
        cores = os.cpu_count() or 1
        n_sample = args.cpu_sample or 640 * cores
        v, used, sample, lps, ms = cpu_arm(panel, synth.TABLE_SEED, args.table_keys, n_sample, max(1, args.steps),
                                           min(args.warmup, 1), cores)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / u32 counts / f64 solver",
            "data": "synthetic", "config": {"workload": workload, "step": "bounded sample: " + sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": "port", "sample": sample,
                             "issued_lookups_per_s": lps},
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
    barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        plan.launch(stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
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
    barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        text, status = e2e_step()
    torch.cuda.synchronize(dev)
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
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
    text = text.tobytes().decode("ascii")

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
    clocks = sampler.stop()

    # ---- parity spot check against the oracle (not timed) ----------------------------------------
    parity = None
    cpu = None
    if rank == 0:
        from oracle import km_oracle as ko
        from oracle.compare import compare_rows
        from oracle.store import KmerStore
        store = KmerStore(31, True, len(panel.keys))
        store.set_background(synth.TABLE_SEED, table_keys)
        store.insert(all_keys, all_counts)
        jf = ko.OracleJellyfish(store, "panel.jf", 0.05, 5)
        lines = text.split("\n")
        by_target = {}
        for ln in lines:
            if ln:
                by_target.setdefault(ln.split("\t")[1], []).append(ln)
        bad = flips = checked = 0
        for i in range(0, args.targets, max(1, args.targets // 64)):
            f = ko.OracleFinder(ko.Target(panel.targets[i], panel.names[i], 31), jf).run()
            errs, fl = compare_rows([str(r) for r in f.get_paths()], by_target.get(panel.names[i], []))
            bad += 1 if errs else 0
            flips += fl
            checked += 1
        parity = {"targets_checked": checked, "mismatching": bad, "printed_digit_flips": flips}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, used, sample, lps, _ = cpu_arm(panel, synth.TABLE_SEED, table_keys, args.cpu_sample or 640 * cores, 1, 1, cores)
            cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "port", "sample": sample,
                   "issued_lookups_per_s": lps}

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
            "config": {"workload": workload, "targets_per_gpu": args.targets, "table_keys": table_keys,
                       "table_distinct": info["n_keys"], "table_gb": info["bytes"] / 1e9, "table_build_s": t_build,
                       "table_layout": "family lines (128 B, two copies per k-mer)" if info["layout"] else "sector buckets (32 B)",
                       "parallelism": "targets sharded x%d, table replicated, no data-path collective" % world,
                       "l2": "table (%.0f GB) and per-step visited sets are far larger than the 126 MB L2; no explicit flush"
                             % (info["bytes"] / 1e9),
                       "ref_kmers": int(n_ref.sum()), "rows": n_rows, "capacity_retries": retries},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "device_ms": e2e_breakdown,
                    "what": "km_find_text(host buffers: sequences + offsets, names) -> the TSV text km find_mutation prints; "
                            "device_ms = the same work as two calls (km_find_batch, km_result_text), not pipelined"},
            "gpu_launches": 7 * args.steps,
            "kernels": {"km_ref_probe_kernel_ms": probe_ms, "km_walk_kernels_ms": walk_ms, "km_graph_kernels_ms": graph_ms,
                        "what": "reference probe (HBM-bound: ~87% of the panel's lookups), shared-memory + general walk "
                                "(latency-bound tails), graph/paths/quantification (shared memory, latency-bound)"},
            "roofline": {"kernel": "km_ref_probe_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "units_per_launch": probe_lookups, "bytes_per_unit": 32,
                         "unit_is": "one canonical k-mer lookup = one 32-byte HBM sector (SURVEY.md 8d)",
                         "random_gather_GBps": gather["GBps"] if gather else None,
                         "frac_of_random_gather": (achieved / gather["GBps"]) if gather else None,
                         "panel": {"algorithmic_lookups": algorithmic, "issued_lookups": issued,
                                   "GBps_over_whole_step": algorithmic * 32 / (ms_per_step / 1e3) / 1e9,
                                   "frac_of_random_gather_over_whole_step":
                                       (algorithmic * 32 / (ms_per_step / 1e3) / 1e9 / gather["GBps"]) if gather else None}},
            "lookup": lookup, "random_gather": gather,
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
