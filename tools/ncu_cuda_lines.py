"""Top source lines of a kernel in an ncu report by stall samples / executed instructions, from ncu's own CUDA<->SASS
correlation (--print-source cuda,sass; the report must have been taken with --import-source on and -lineinfo).
    python tools/ncu_cuda_lines.py report.ncu-rep <kernel substring> [top N] [column]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
col = sys.argv[4] if len(sys.argv) > 4 else "# Samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = csv.reader(out.split("\n"))
cur_file, cur_fn, hdr = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur_fn and kern in cur_fn and r[0].isdigit() and len(r) >= len(hdr) - 1:
        try:
            v = float(r[hdr.index(col)] or 0)
            ie = float(r[hdr.index("Instructions Executed")] or 0)
        except (ValueError, IndexError):
            continue
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0.0, 0.0, r[1]])
        a[0] += v
        a[1] += ie
tot = sum(a[0] for a in agg.values()) or 1.0
toti = sum(a[1] for a in agg.values()) or 1.0
print("kernel ~ %s: total %s %.0f, instructions %.0f" % (kern, col, tot, toti))
for (f, ln), (v, ie, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100 * v / tot, 100 * ie / toti, f, ln, src.strip()[:105]))
