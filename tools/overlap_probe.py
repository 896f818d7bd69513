"""How much do the find kernels gain from running sub-batches side by side?  The panel is cut into S
interleaved slices, each a resident plan launched on its own stream; the whole panel is timed with CUDA
events on a master stream that forks to / joins from the S streams.  (Probe = HBM request rate, walk =
dependent-load latency, graph = issue-bound shared-memory work: different resources.)

    python tools/overlap_probe.py [--targets 10000] [--table-keys 2000000000] [--subs 1,2,3,4,6,8]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=10000)
    ap.add_argument("--table-keys", type=int, default=2_000_000_000)
    ap.add_argument("--subs", default="1,2,3,4,6,8")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from km_b200 import engine, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    panel = synth.make_panel(args.targets, seed=synth.PANEL_SEED)
    table = engine.Table.create(k=31, canonical=True, capacity=args.table_keys + len(panel.keys), device=0)
    table.build_synthetic(synth.TABLE_SEED, args.table_keys)
    table.insert(panel.keys, panel.counts, mode="overwrite")
    out = {"layout": table.info()["layout"], "runs": []}
    master = torch.cuda.Stream(dev)
    for S in [int(x) for x in args.subs.split(",")]:
        for how in ("interleaved", "contiguous"):
            if S == 1 and how == "contiguous":
                continue
            if how == "interleaved":
                slices = [panel.targets[i::S] for i in range(S)]
            else:
                step = (len(panel.targets) + S - 1) // S
                slices = [panel.targets[i * step:(i + 1) * step] for i in range(S)]
            plans = [table.plan(s) for s in slices]
            streams = [torch.cuda.Stream(dev) for _ in range(S)]

            def launch_all():
                fork = torch.cuda.Event()
                fork.record(master)
                for p, s in zip(plans, streams):
                    s.wait_event(fork)
                    p.launch(s.cuda_stream)
                    e = torch.cuda.Event()
                    e.record(s)
                    master.wait_event(e)

            for _ in range(3):
                launch_all()
            torch.cuda.synchronize(dev)
            for p in plans:
                p.fetch(want_graph=False)        # settle capacities
            for _ in range(2):
                launch_all()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(master)
            for _ in range(args.steps):
                launch_all()
            e1.record(master)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / args.steps
            out["runs"].append({"subs": S, "how": how, "ms_per_panel": ms, "targets_per_s": args.targets / ms * 1e3})
            print(out["runs"][-1], file=sys.stderr)
            for p in plans:
                p.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
