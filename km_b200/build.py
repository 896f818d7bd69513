"""Builds km_b200/libkm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m km_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkm_b200.so")
SOURCES = ["api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "km_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if os.environ.get("KM_PHASE_TIMERS") and "-DKM_PHASE_TIMERS" not in NVCC_FLAGS:
        NVCC_FLAGS.append("-DKM_PHASE_TIMERS")
        force = True
    extra = os.environ.get("KM_NVCC_EXTRA", "").split()        # experiments: e.g. KM_NVCC_EXTRA=-DKM_WALK_WARPS=1
    if extra and not set(extra) <= set(NVCC_FLAGS):
        NVCC_FLAGS.extend(extra)
        force = True
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libkm_b200.so")
    with open(os.path.join(HERE, "csrc", ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
