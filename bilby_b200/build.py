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
TORCH_LIB = os.path.join(LIB_DIR, "libbilby_b200_torch.so")
SOURCES = ["bb_kernels.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false",
              # one soname for the library and every experiment build (BB_LIB_OUT): whichever copy ctypes loads first also
              # satisfies the torch shim's DT_NEEDED, so the shim never pulls a second copy in beside it
              "-Xlinker", "-soname=libbilby_b200.so"]


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
    if "BB_LIB_OUT" not in os.environ:
        build_torch_shim()
    return LIB


def build_torch_shim():
    """csrc/bb_torch.cpp -> _lib/libbilby_b200_torch.so: TORCH_LIBRARY ops over the C ABI (g++, ~15 s)."""
    import torch
    from torch.utils import cpp_extension
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", os.path.join(CSRC, "bb_torch.cpp"), "-o", TORCH_LIB]
    cmd += [f"-I{i}" for i in cpp_extension.include_paths()] + ["-I/usr/local/cuda/include"]
    cmd += [f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", f"-L{LIB_DIR}", "-lbilby_b200",
            "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}",
            "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed: " + " ".join(cmd))
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
