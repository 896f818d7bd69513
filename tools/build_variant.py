"""Build a tuning variant of libkm_b200.so beside the product's: tools/variants/libkm_b200_<tag>.so, compiled with
extra -D flags (e.g. -DKM_CTA=64).  Select it at run time with KM_B200_LIB=<path>.  Measurement aid only.
    python tools/build_variant.py <tag> -DKM_CTA=64 -DKM_GRAPH_TINY_MINB=16 ...
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from km_b200 import build as kb      # noqa: E402


def main():
    tag, extra = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "tools", "variants")
    obj_dir = os.path.join(out_dir, ".obj_" + tag)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [s for s in kb.SOURCES if os.path.exists(os.path.join(kb.CSRC, s))]

    def run(s):
        obj = os.path.join(obj_dir, s[:-3] + ".o")
        r = subprocess.run([nvcc, *kb.NVCC_FLAGS, *extra, "-c", "-o", obj, os.path.join(kb.CSRC, s)], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise SystemExit("nvcc failed on " + s)
        return obj, r.stderr
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
        res = list(pool.map(run, srcs))
    lib = os.path.join(out_dir, "libkm_b200_%s.so" % tag)
    subprocess.check_call([nvcc, "-shared", "-o", lib, *[o for o, _ in res], "-lcudart"])
    with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
        f.write("".join(log for _, log in res))
    print(lib)


if __name__ == "__main__":
    main()
