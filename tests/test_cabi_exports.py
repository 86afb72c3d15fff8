"""CPU test: libngp_b200.so loads and exports exactly what include/ngp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ngp_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    names = _declared()
    # one per native function the reference binds (raymarching.h:7-17, gridencoder.h:12-13, freqencoder.h:7,10)
    for n in ["grid_encode_forward", "grid_encode_backward", "near_far_from_aabb", "sph_from_ray", "morton3D",
              "morton3D_invert", "packbits", "march_rays_train", "composite_rays_train_forward",
              "composite_rays_train_backward", "march_rays", "composite_rays", "freq_encode_forward",
              "freq_encode_backward"]:
        assert "ngp_" + n in names


def test_library_exports_every_declared_symbol():
    from ngp_b200 import _cabi
    lib = _cabi.load()
    for name in _declared():
        assert hasattr(lib, name), "missing export " + name
        assert name in _cabi.SIGNATURES, "ctypes signature missing for " + name
    assert set(_cabi.SIGNATURES) == set(_declared())
    assert lib.ngp_version() >= 100
    assert "unsupported" in _cabi.error_string(-2)


def test_no_torch_types_in_the_abi():
    text = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "at::" not in code and "torch" not in code and "std::" not in code
    assert 'extern "C"' in text


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "single-stable-dreamfusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
                assert "ngp_oracle" not in src, os.path.join(dp, f)


def test_ops_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import raymarching
    with pytest.raises(Exception):
        raymarching.near_far_from_aabb(torch.zeros(4, 3), torch.ones(4, 3), torch.tensor([-1., -1, -1, 1, 1, 1]))


def test_ctypes_signatures_match_the_header_argument_for_argument():
    """Every prototype of include/ngp_b200.h against its ctypes signature in ngp_b200/_cabi.py: same number of arguments
    and, position by position, the same kind (pointer / uint32_t / uint64_t / int / float).  A drifted signature would
    otherwise only show up on a GPU, as a garbage argument."""
    import ctypes as C
    from ngp_b200 import _cabi
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = re.findall(r"\b(int|uint64_t|const char\s*\*)\s+(ngp_[A-Za-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(decls) >= 60

    def kind_of_c(param):
        p = " ".join(param.split())
        if "*" in p:
            return "ptr"
        for t in ("uint64_t", "uint32_t", "float", "int"):
            if re.match(r"(const )?%s\b" % t, p):
                return t
        raise AssertionError("unparsed parameter: " + p)

    def kind_of_ctypes(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_uint32: "uint32_t", C.c_uint64: "uint64_t", C.c_int: "int", C.c_float: "float"}[t]

    for ret, name, params in decls:
        res, args = _cabi.SIGNATURES[name]
        params = params.strip()
        plist = [] if params in ("", "void") else [p for p in params.split(",")]
        assert len(plist) == len(args), "%s: header has %d arguments, ctypes %d" % (name, len(plist), len(args))
        for i, (p, a) in enumerate(zip(plist, args)):
            assert kind_of_c(p) == kind_of_ctypes(a), "%s argument %d: header `%s`, ctypes %s" % (name, i, p.strip(), a)
        want = {"int": C.c_int, "uint64_t": C.c_uint64}.get(ret, C.c_char_p)
        assert res is want, name
