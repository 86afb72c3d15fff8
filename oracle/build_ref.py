#!/usr/bin/env python
"""TEST INFRASTRUCTURE - builds the reference's own CUDA extensions as the parity oracle.

Compiles the UNMODIFIED sources where they lie under /root/reference
(gridencoder/src, raymarching/src, freqencoder/src) into pybind modules
``oracle/_ref/_gridencoder.so``, ``_raymarching.so`` and ``_freqencoder.so``.
Nothing is copied into the repo: only the built binaries land in
``oracle/_ref/`` (git-ignored, but shipped to the GPU box by gpurun).

The single deviation from the reference's own build flags
(/root/reference/gridencoder/backend.py:6-9, raymarching/backend.py:6-9,
freqencoder/backend.py:6-10) is ``-std=c++17`` instead of ``-std=c++14``:
torch 2.11's ATen headers refuse to compile below C++17.  freqencoder keeps
its ``-use_fast_math``.  We do not run the reference's build system; this is
the short recipe the task allows.

Only tests/, bench.py's reference legs and __graft_entry__.build() may call this.
"""
import os
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NGP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "_gridencoder": ("gridencoder/src", ["gridencoder.cu", "bindings.cpp"], []),
    "_raymarching": ("raymarching/src", ["raymarching.cu", "bindings.cpp"], []),
    "_freqencoder": ("freqencoder/src", ["freqencoder.cu", "bindings.cpp"], ["-use_fast_math"]),
}


def _torch_paths():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths("cuda")
    lib = ce.library_paths("cuda")
    return inc, lib


def build_one(name, force=False, verbose=True):
    sub, files, extra = EXTS[name]
    so = os.path.join(OUT, name + ".so")
    srcs = [os.path.join(REF, sub, f) for f in files]
    if not all(os.path.exists(s) for s in srcs):
        return None
    if os.path.exists(so) and not force:
        if all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
            return so
    os.makedirs(OUT, exist_ok=True)
    inc, lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    objs = []
    t0 = time.time()
    for s in srcs:
        obj = os.path.join(OUT, name + "_" + os.path.basename(s) + ".o")
        cmd = ["nvcc", "-c", s, "-o", obj, "-O3", "-std=c++17",
               "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
               "-U__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr",
               "-gencode", "arch=compute_100,code=sm_100",
               "-Xcompiler", "-fPIC", "-DTORCH_EXTENSION_NAME=" + name,
               "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1",
               "-I" + pyinc] + ["-I" + i for i in inc] + extra
        if verbose:
            print("[build_ref]", name, os.path.basename(s), flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = ["g++", "-shared", "-o", so] + objs + ["-L" + l for l in lib] + [
        "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
        "-Wl,-rpath," + lib[0]]
    subprocess.check_call(cmd)
    for o in objs:
        os.remove(o)
    if verbose:
        print("[build_ref] built %s in %.0fs" % (so, time.time() - t0), flush=True)
    return so


def build_all(force=False, verbose=True):
    """Build every reference extension (in parallel). Returns {name: path or None}."""
    if not os.path.isdir(REF):
        return {n: (os.path.join(OUT, n + ".so") if os.path.exists(os.path.join(OUT, n + ".so")) else None)
                for n in EXTS}
    with ThreadPoolExecutor(max_workers=3) as ex:
        futs = {n: ex.submit(build_one, n, force, verbose) for n in EXTS}
        return {n: f.result() for n, f in futs.items()}


if __name__ == "__main__":
    res = build_all(force="--force" in sys.argv)
    print(res)
