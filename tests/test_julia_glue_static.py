"""Julia is not installed here, so julia/DynamicSparseArraysB200.jl cannot be executed.  This static check keeps it honest:
every `ccall((:dsa_xxx, libdsa), Ret, (ArgTypes...), ...)` in the glue must name a function declared in include/dsa.h with the
same number of parameters, and each Julia argument type must be a legal spelling of the C parameter type."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# C parameter type (normalised) -> Julia ccall argument types that are ABI-compatible with it
COMPAT = {
    "int64_t": {"Int64"},
    "int": {"Cint"},
    "double": {"Float64", "Cdouble"},
    "const int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "const double*": {"Ptr{Float64}", "Ref{Float64}"},
    "double*": {"Ptr{Float64}", "Ref{Float64}"},
    "uint8_t*": {"Ptr{UInt8}", "Ref{UInt8}"},
    "void*": {"Ptr{Cvoid}"},
    "const void*": {"Ptr{Cvoid}"},
    "char*": {"Ptr{UInt8}", "Cstring"},
    "handle": {"Ptr{Cvoid}"},                 # dsa_vec_t* / dsa_matrix_t* / dsa_dist_t* / dsa_dmatrix_t* (const or not)
    "handle*": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
}
RET = {"int": "Cint", "const char*": "Cstring", "int64_t": "Int64", "dsa_matrix_t*": "Ptr{Cvoid}"}


def _header_decls():
    text = open(os.path.join(ROOT, "include", "dsa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(dsa_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        norm = []
        for p in plist:
            p = re.sub(r"\s+", " ", p)
            ty = re.sub(r"\s*[A-Za-z_][A-Za-z0-9_]*$", "", p).strip() if not p.endswith("*") else p   # drop the parameter name
            ty = ty.replace(" *", "*")
            stars = ty.count("*")
            base = ty.replace("*", "").strip()
            if re.search(r"dsa_(vec|matrix|dist|dmatrix)_t", base):
                norm.append("handle" + "*" * (stars - 1))
            elif base == "const void":
                norm.append("const void" + "*" * stars)
            else:
                norm.append(base + "*" * stars)
        decls[name] = (re.sub(r"\s+", " ", ret).replace(" *", "*"), norm)
    return decls


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ccalls():
    src = open(os.path.join(ROOT, "julia", "DynamicSparseArraysB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(dsa_[a-z0-9_]+),\s*libdsa\),\s*([A-Za-z0-9_{}]+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                      # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = _split_top(src[m.end():i - 1])
        calls.append((m.group(1), m.group(2), types))
    return calls


def test_every_ccall_matches_the_header():
    decls, calls = _header_decls(), _ccalls()
    assert len(calls) >= 20, "the glue should bind the whole hot path"
    for name, ret, types in calls:
        assert name in decls, f"{name} is not declared in include/dsa.h"
        cret, cparams = decls[name]
        assert RET[cret] == ret, (name, cret, ret)
        assert len(types) == len(cparams), (name, types, cparams)
        for jt, ct in zip(types, cparams):
            assert jt in COMPAT[ct], f"{name}: Julia {jt} does not match C {ct}"


def test_glue_binds_the_reference_exports():
    """the exported names of the reference (src/DynamicSparseArrays.jl:5-16) are all defined by the glue"""
    src = open(os.path.join(ROOT, "julia", "DynamicSparseArraysB200.jl")).read()
    for name in ("DynamicSparseVector", "DynamicSparseMatrix", "dynamicsparsevec", "dynamicsparse", "nbpartitions", "deletecolumn!",
                 "deleterow!", "addrow!", "closefillmode!", "shrink_size!"):
        assert re.search(r"\b" + re.escape(name), src), name
