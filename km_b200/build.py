"""Builds km_b200/libkm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m km_b200.build [--force]

One object per translation unit (kernel families and API parts compile side by side and only what
changed is recompiled), then one link.  Objects live in km_b200/csrc/.obj/ (git-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, ".obj")
LIB = os.path.join(HERE, "libkm_b200.so")
SOURCES = ["table_api.cu", "io_api.cu", "plan_api.cu", "text_api.cu", "walk_kernels.cu", "graph_kernels.cu",
           "format_kernels.cu", "cohort_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
           [os.path.join(HERE, "..", "include", "km_b200.h")]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    if os.environ.get("KM_PHASE_TIMERS") and "-DKM_PHASE_TIMERS" not in NVCC_FLAGS:
        NVCC_FLAGS.append("-DKM_PHASE_TIMERS")
        force = True
    extra = os.environ.get("KM_NVCC_EXTRA", "").split()        # experiments: e.g. KM_NVCC_EXTRA=-DKM_WALK_WARPS=1
    if extra and not set(extra) <= set(NVCC_FLAGS):
        NVCC_FLAGS.extend(extra)
        force = True
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest(_headers())
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flag_tag = os.path.join(OBJ, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS)
    if not os.path.exists(flag_tag) or open(flag_tag).read() != flags_now:
        force = True

    def stale(src, obj):
        return force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_time, os.path.getmtime(src))

    jobs = []
    for s in sources:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if stale(src, obj):
            jobs.append((s, [nvcc, *NVCC_FLAGS, "-c", "-o", obj, src]))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources]
    if not jobs and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(objs):
        return LIB

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    logs = []
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for name, r in pool.map(run, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed compiling %s" % name)
            logs.append("==== %s\n%s" % (name, r.stderr))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libkm_b200.so")
    with open(flag_tag, "w") as f:
        f.write(flags_now)
    with open(os.path.join(CSRC, ".ptxas.log"), "a" if not force else "w") as f:
        f.write("".join(logs))
    if verbose:
        sys.stderr.write("".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
