"""Build the CUDA library (sm_100a only) in-tree: bilby_b200/_lib/libbilby_b200.so.

    python -m bilby_b200.build            # compile if sources are newer than the library
    python -m bilby_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libbilby_b200.so")
SOURCES = ["bb_kernels.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, name)))
    return m


def build(force=False, verbose=False):
    """BB_NVCC_EXTRA (extra nvcc flags, e.g. -DBB_K1_UNROLL=2) and BB_LIB_OUT (output path) serve kernel experiments."""
    global LIB
    LIB = os.environ.get("BB_LIB_OUT", LIB)
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + os.environ.get("BB_NVCC_EXTRA", "").split()
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
