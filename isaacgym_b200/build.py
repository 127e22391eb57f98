"""Build the CUDA library in-tree: isaacgym_b200/_lib/libppk.so (sm_100a only).

    python -m isaacgym_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  `-fmad=false`: the reference's flags compare fp32 values that
each come from one correctly rounded ATen op; FMA contraction would move values across thresholds.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libppk.so")
STAMP = os.path.join(LIB_DIR, "libppk.stamp")
SOURCES = ["ppk_api.cu", "ppk_host.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--shared", "-cudart", "shared",
] + os.environ.get("PPK_NVCC_EXTRA", "").split()      # e.g. -DPPK_LIBM_HEADING, -DPPK_TRACE for A/B builds


def _source_hash():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/ppk.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def stamp_matches():
    return os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == _source_hash()


def build(force=False, verbose=False, out=None):
    """`out`: build a variant (PPK_NVCC_EXTRA) somewhere else, e.g. scratch/libs/x.so for A/B runs with PPK_LIB."""
    os.makedirs(LIB_DIR, exist_ok=True)
    want = _source_hash()
    if out is None and not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == want:
        return LIB_PATH
    nvcc = nvcc_path()
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libppk.so")
    if verbose:
        print(res.stdout + res.stderr)
    if out is not None:
        return out
    with open(STAMP, "w") as f:
        f.write(want)
    return LIB_PATH


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out)
    print(p)
