"""One-off measurements on a B200: random 32-B sector gather ceiling and the device-resident
lookup rate at a few table sizes.  Writes gpurun_out/probe.json."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from km_b200 import engine, synth          # noqa: E402
from km_b200._lib import lib, check        # noqa: E402

out = {"gather": [], "lookup": []}
L = lib()
for gib in (1, 8, 32, 96):
    ms = ctypes.c_float()
    n_loads = 1 << 28
    check(L.km_bench_random_gather(0, gib << 30, n_loads, 3, ctypes.byref(ms)))
    out["gather"].append({"gib": gib, "n_loads": n_loads, "ms": ms.value,
                          "gsectors_per_s": n_loads / ms.value / 1e6, "GBps": n_loads * 32 / ms.value / 1e6})
    print(out["gather"][-1], flush=True)
for n_keys in (1 << 24, 1 << 28, 2_000_000_000):
    t0 = time.time()
    t = engine.Table.create(capacity=n_keys)
    t.build_synthetic(synth.TABLE_SEED, n_keys)
    build_s = time.time() - t0
    best, mean, hits = ctypes.c_float(), ctypes.c_float(), ctypes.c_uint64()
    nq = 1 << 28
    check(L.km_bench_lookup(t._h, synth.TABLE_SEED, n_keys, nq, synth.QUERY_SEED, 5, ctypes.byref(best),
                            ctypes.byref(mean), ctypes.byref(hits)))
    info = t.info()
    out["lookup"].append({"n_keys": n_keys, "distinct": info["n_keys"], "table_gb": info["bytes"] / 1e9,
                          "build_s": build_s, "n_queries": nq, "best_ms": best.value, "mean_ms": mean.value,
                          "hit_frac": hits.value / nq, "glookups_per_s": nq / best.value / 1e6,
                          "sector_GBps": nq * 32 / best.value / 1e6})
    print(out["lookup"][-1], flush=True)
    t.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
