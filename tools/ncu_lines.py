"""Attribute ncu warp-stall samples to source lines without the GUI (NCU_COLUMN="Instructions Executed"
ranks lines by executed warp-instructions instead).

    python tools/ncu_lines.py report.ncu-rep km_b200/libkm_b200.so <kernel substring> [top N]

ncu's CSV source page is SASS-only, so the SASS offsets are joined with `nvdisasm -g` line
annotations of the cubin inside the .so (built with -lineinfo)."""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    ncu_name = sys.argv[5] if len(sys.argv) > 5 else kern      # demangled base name for ncu -k
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, capture_output=True)
        # one cubin per translation unit: disassemble them all (the kernel is in one of them)
        sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
                       for f in sorted(os.listdir(d)) if f.endswith(".cubin"))
    # offset -> (file, line) for the wanted kernel
    line_of = {}
    inside = False
    cur = None
    for ln in sass.split("\n"):
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = None
    base = None
    per_line, per_ins = {}, []
    total = 0
    wanted = False
    for r in rows:
        if r and r[0] == "Kernel Name":
            wanted = ncu_name in r[1]
            hdr = None
            continue
        if not wanted:
            continue
        if r and r[0] == "Address":
            hdr = r
            base = None
            continue
        if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
            a = int(r[0], 16)
            if base is None:
                base = a
            s = int(float(r[hdr.index(os.environ.get("NCU_COLUMN", "# Samples"))] or 0))
            total += s
            key = line_of.get(a - base, (("?", 0), r[1]))[0]
            per_line[key] = per_line.get(key, 0) + s
            per_ins.append((s, a - base, key, r[1].strip()))
    print("total samples", total)
    srcs = {}
    for (f, l), s in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
            srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
        print("%7d %5.1f%%  %s:%d  %s" % (s, 100.0 * s / max(1, total), f, l, text))
    print("-- hottest instructions")
    for s, off, key, ins in sorted(per_ins, reverse=True)[:15]:
        print("%7d  +%05x %s:%d  %s" % (s, off, key[0], key[1], ins[:80]))


if __name__ == "__main__":
    main()
