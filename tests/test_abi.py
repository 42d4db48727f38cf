"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/bdf_b200.h
declares; the ctypes table covers the same set; without a GPU the engine fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bdf_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bdf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import bdf_b200
    from bdf_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 20
    if not os.path.exists(_lib.LIB_PATH):
        from bdf_b200 import build  # noqa: F401

        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in bdf_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.bdf_version() >= 100


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bdf_b200

    with pytest.raises(bdf_b200.BDFError) as ei:
        bdf_b200.Engine(10)
    assert "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bayesiandatafusion.jl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_julia_glue_binds_only_declared_symbols_with_matching_arity():
    """julia/BDFCuda.jl is the reference-side binding a maintainer adds; it cannot run here, so at least
    every `ccall((:sym, LIB), Cint, (argtypes...), ...)` must name a declared entry and pass as many arguments as the
    C prototype takes."""
    src = open(os.path.join(ROOT, "julia", "BDFCuda.jl")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(bdf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    calls = re.findall(r"ccall\(\(:(bdf_[a-z0-9_]+), LIB\),\s*\w+,\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)", src, flags=re.S)
    assert len(calls) >= 25
    for name, argtypes in calls:
        assert name in protos, f"BDFCuda.jl binds {name}, which include/bdf_b200.h does not declare"
        nargs = len([a for a in re.split(r",\s*(?![^{}]*\})", argtypes.strip().rstrip(",")) if a.strip()])
        assert nargs == protos[name], f"{name}: Julia passes {nargs} arguments, the C prototype takes {protos[name]}"


def test_julia_glue_argument_types_match_the_c_prototypes():
    """Every ccall of julia/BDFCuda.jl and julia/macau_cuda.jl passes, position by position, the Julia type that corresponds to the C
    parameter type of include/bdf_b200.h (Ptr{Cdouble} for double*, Int64 for int64_t, ...): the glue cannot be run here, so its
    signatures are checked against the header mechanically."""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    hdr = re.sub(r"//.*", "", hdr)

    def ctype(a):
        a = a.strip()
        if not a.endswith("*"):
            a = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", a).strip()   # drop the parameter name
        return a.replace("const ", "").replace(" ", "")

    protos = {}
    for m in re.finditer(r"\b(bdf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = [] if args in ("", "void") else [ctype(x) for x in args.split(",")]
    julia_of = {"bdf_t*": {"Ptr{Void}"}, "bdf_t**": {"Ptr{Ptr{Void}}"}, "int": {"Cint"}, "int64_t": {"Int64"}, "double": {"Cdouble"},
                "double*": {"Ptr{Cdouble}"}, "int64_t*": {"Ptr{Int64}"}, "int32_t*": {"Ptr{Int32}", "Ptr{Cint}"}, "int*": {"Ptr{Cint}", "Ptr{Int32}"},
                "uint64_t": {"UInt64"}, "void*": {"Ptr{Void}"}, "void**": {"Ptr{Ptr{Void}}"}, "unsignedchar*": {"Ptr{UInt8}"}, "char*": {"Ptr{UInt8}"},
                "float*": {"Ptr{Float32}", "Ptr{Cfloat}"}}
    checked = 0
    for f in ("BDFCuda.jl", "macau_cuda.jl"):
        src = open(os.path.join(ROOT, "julia", f)).read()
        for name, argtypes in re.findall(r"ccall\(\(:(bdf_[a-z0-9_]+), LIB\),\s*\w+,\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)", src, flags=re.S):
            jt = [a.strip() for a in re.split(r",\s*(?![^{}]*\})", argtypes.strip().rstrip(",")) if a.strip()]
            ct = protos[name]
            assert len(jt) == len(ct), name
            for i, (j, c) in enumerate(zip(jt, ct)):
                assert c in julia_of, f"{name}: no Julia mapping for C type {c}"
                assert j in julia_of[c], f"{f}: {name} argument {i + 1} is {j} in Julia but {c} in C"
                checked += 1
    assert checked >= 150
