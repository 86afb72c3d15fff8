"""Builds libngp_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch headers are involved, so a full rebuild takes seconds-to-a-minute instead of the
~12 minutes the reference's three pybind extensions need.  Translation units are compiled in
parallel and cached by (source mtime, flags).  ``freq_encode.cu`` alone gets ``--use_fast_math``
because the reference builds its freqencoder that way (freqencoder/backend.py:9) and bit-parity
of ``__sinf(scalbnf(x)+phase)`` depends on it.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(LIB_DIR, "libngp_b200.so")
INCLUDE = os.path.abspath(os.path.join(HERE, "..", "..", "include"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "186",
          "-I" + INCLUDE, "-I" + CSRC]
PER_FILE = {"freq_encode.cu": ["--use_fast_math"]}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "ngp_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, src + ".o")
    stamp = obj + ".stamp"
    flags = ARCH + COMMON + PER_FILE.get(src, [])
    path = os.path.join(CSRC, src)
    key = hashlib.sha1((" ".join(flags) + str(os.path.getmtime(path)) + str(_deps_mtime())).encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    cmd = ["nvcc", "-c", path, "-o", obj] + flags
    if verbose:
        print("[ngp_b200.build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(key)
    return obj, True


def build(verbose=False, force=False):
    """Compile every .cu under csrc/ for sm_100a and link lib/libngp_b200.so. Returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [r[0] for r in res]
    if any(r[1] for r in res) or not os.path.exists(LIB_PATH):
        cmd = ["nvcc", "-shared", "-o", LIB_PATH] + objs + ARCH
        if verbose:
            print("[ngp_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
