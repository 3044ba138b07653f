"""The C-ABI library loads and exports every symbol include/reach.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from reach_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "reach.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reach_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), "libreach_cuda.so does not export %s" % name
        assert name in _lib.SIGNATURES, "%s is not bound in _lib.SIGNATURES" % name
    assert sorted(_lib.SIGNATURES) == names
    assert lib.reach_version() >= 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.reach_term) == 4 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.reach_setexpr) == 8 + 8 + 8
    assert ctypes.sizeof(_lib.reach_lgg09_opts) == 8 * 4
    assert ctypes.sizeof(_lib.reach_invariant) == 8 + 8 + 8          # kind + m share the first 8 bytes


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    lib = _lib.load()
    if lib.reach_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.ReachError, match="no usable CUDA device"):
        _lib.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "reachabilityanalysis.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def _c_param_counts():
    """name -> number of parameters, parsed from the prototypes in include/reach.h."""
    src = open(os.path.join(ROOT, "include", "reach.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(reach_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_ctypes_signatures_have_the_header_arity():
    counts = _c_param_counts()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert counts[name] == len(argtypes), "%s: header has %d parameters, ctypes binding %d" % (
            name, counts[name], len(argtypes))


def test_julia_wrapper_ccalls_match_the_header():
    """julia/ReachCUDA.jl cannot be executed here (no julia binary); at least every ccall must name a declared symbol
    and pass as many arguments as the C prototype takes (type tuple and value list)."""
    counts = _c_param_counts()
    src = open(os.path.join(ROOT, "julia", "ReachCUDA.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = 0
    for m in re.finditer(r"ccall\(\(:(reach_[a-z0-9_]+),\s*lib\),\s*\w+,\s*\(", src):
        name = m.group(1)
        assert name in counts, "ReachCUDA.jl calls %s, which include/reach.h does not declare" % name
        i, depth = m.end(), 1                       # scan the type tuple
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = src[m.end():i - 1].strip().rstrip(",")
        ntypes = 0 if not types else len(_split_top_level(types))
        j, depth = i, 1                             # the value list runs to the closing paren of ccall
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(src[j], 0)
            j += 1
        values = src[i:j - 1].strip().lstrip(",").strip()
        nvalues = 0 if not values else len(_split_top_level(values))
        assert ntypes == counts[name], "%s: %d argument types, header has %d" % (name, ntypes, counts[name])
        assert nvalues == counts[name], "%s: %d argument values, header has %d" % (name, nvalues, counts[name])
        calls += 1
    assert calls >= 8


def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def _julia_code_only(s):
    """Julia source with string literals emptied and comments removed."""
    out, i, n = [], 0, len(s)
    while i < n:
        if s.startswith('"""', i):
            j = s.find('"""', i + 3)
            i = j + 3 if j >= 0 else n
            out.append('""')
        elif s[i] == '"':
            j = i + 1
            while j < n and s[j] != '"':
                j += 2 if s[j] == "\\" else 1
            i = j + 1
            out.append('""')
        elif s.startswith("#=", i):
            j = s.find("=#", i + 2)
            i = j + 2 if j >= 0 else n
        elif s[i] == "#":
            j = s.find("\n", i)
            i = j if j >= 0 else n
        else:
            out.append(s[i])
            i += 1
    return "".join(out)


def test_julia_wrapper_blocks_and_brackets_balance():
    """No julia binary here: at least every block opener has its `end`, brackets nest, the module closes, and every
    name the wrapper takes from ReachabilityAnalysis / LazySets in a qualified import is one it uses."""
    code = _julia_code_only(open(os.path.join(ROOT, "julia", "ReachCUDA.jl")).read())
    openers = {"function", "if", "for", "while", "let", "begin", "struct", "module", "do", "try", "quote", "macro"}
    closing = {"(": ")", "[": "]", "{": "}"}
    blocks, brackets, line = [], [], 1
    for m in re.finditer(r":?[^\W\d][\w!]*|[\[\](){}]|\n", code):
        w = m.group(0)
        if w == "\n":
            line += 1
        elif w in closing:
            brackets.append((w, line))
        elif w in closing.values():
            assert brackets, "line %d: unmatched %s" % (line, w)
            o, l = brackets.pop()
            assert closing[o] == w, "line %d: %s closes the %s of line %d" % (line, w, o, l)
        elif w in openers:
            if brackets and w in ("for", "if"):        # comprehension / generator: no `end`
                continue
            blocks.append((w, line))
        elif w == "end":
            if brackets and brackets[-1][0] == "[":    # indexing: x[end]
                continue
            assert blocks, "line %d: `end` without an open block" % line
            blocks.pop()
    assert not brackets, "unclosed brackets: %s" % brackets
    assert not blocks, "unclosed blocks: %s" % blocks
    for m in re.finditer(r"^import [\w.]+:\s*((?:[^\n,]+,\s*\n?)*[^\n,]+)$", code, flags=re.M):
        for name in re.split(r"[,\s]+", m.group(1).strip()):
            assert len(re.findall(r"(?<![\w.])%s(?![\w!])" % re.escape(name), code)) >= 2, \
                "ReachCUDA.jl imports %s and never uses it" % name


def _c_prototypes():
    """name -> (return type, [parameter types]) from include/reach.h, parameter names and qualifiers stripped."""
    src = open(os.path.join(ROOT, "include", "reach.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:^|[;}\n])\s*((?:const\s+)?[A-Za-z_][\w]*(?:\s*\*+)?)\s*\b(reach_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = re.sub(r"\bconst\b", "", a).strip()
                stars = a.count("*")
                base = re.sub(r"\*", " ", a).split()
                base = base[0] if len(base) == 1 else " ".join(base[:-1])      # drop the parameter name
                params.append(base.strip() + "*" * stars)
        out[name] = (re.sub(r"\bconst\b", "", ret).replace(" ", ""), params)
    return out


_JULIA_SCALARS = {"int": {"Cint", "Int32"}, "int32_t": {"Cint", "Int32"}, "int64_t": {"Int64", "Clonglong"},
                  "size_t": {"Csize_t"}, "double": {"Cdouble", "Float64"}}
_JULIA_STRUCTS = {"reach_setexpr": "ReachSetExpr", "reach_invariant": "ReachInvariant", "reach_lgg09_opts": "ReachLGG09Opts",
                  "reach_term": "ReachTerm"}


def _julia_types_for(ctype):
    """The Julia ccall argument types that are layout-compatible with a C parameter type."""
    stars = ctype.count("*")
    base = ctype.rstrip("*")
    if stars == 0:
        return _JULIA_SCALARS[base]
    if base in _JULIA_STRUCTS:
        inner = {_JULIA_STRUCTS[base]}
    elif base in _JULIA_SCALARS:
        inner = set(_JULIA_SCALARS[base])
    elif base in ("void", "char") or base.startswith("reach_"):
        inner = {"Cvoid"} if base != "char" else {"UInt8", "Cchar"}
    else:
        raise AssertionError("unmapped C type %r" % ctype)
    for _ in range(stars - 1):
        inner = {"Ptr{%s}" % t for t in inner}
    out = {"Ptr{%s}" % t for t in inner} | {"Ref{%s}" % t for t in inner}
    if base == "char" and stars == 1:
        out |= {"Cstring"}
    return out


def test_julia_wrapper_ccall_types_match_the_header():
    """Beyond arity: every ccall's return type and each argument type in julia/ReachCUDA.jl is layout-compatible
    with the C prototype (int <-> Cint, int64_t <-> Int64, double* <-> Ptr{Float64}, opaque handles <-> Ptr{Cvoid},
    handle out-parameters <-> Ref{Ptr{Cvoid}}, the mirrored structs by name)."""
    protos = _c_prototypes()
    assert len(protos) >= 30 and protos["reach_lgg09"][1][3] == "int64_t"
    src = _julia_code_only(open(os.path.join(ROOT, "julia", "ReachCUDA.jl")).read())
    checked = 0
    for m in re.finditer(r"ccall\(\(:(reach_[a-z0-9_]+),\s*lib\),\s*(\w+),\s*\(", src):
        name, jret = m.group(1), m.group(2)
        cret, cparams = protos[name]
        if cret == "void":
            assert jret == "Cvoid", name
        elif cret == "int":
            assert jret == "Cint", name
        elif cret == "char*":
            assert jret == "Cstring", name
        else:
            assert jret in _julia_types_for(cret), (name, jret, cret)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        jtypes = [t.strip() for t in _split_top_level(src[m.end():i - 1].strip().rstrip(","))]
        assert len(jtypes) == len(cparams), name
        for k, (jt, ct) in enumerate(zip(jtypes, cparams)):
            assert jt in _julia_types_for(ct), "%s argument %d: Julia %s against C %s" % (name, k + 1, jt, ct)
        checked += 1
    assert checked >= 20


def _ctypes_for(ctype):
    """The ctypes argument types that are layout-compatible with a C parameter type."""
    stars = ctype.count("*")
    base = ctype.rstrip("*")
    scalars = {"int": {ctypes.c_int, ctypes.c_int32}, "int32_t": {ctypes.c_int, ctypes.c_int32},
               "int64_t": {ctypes.c_int64, ctypes.c_longlong}, "size_t": {ctypes.c_size_t}, "double": {ctypes.c_double}}
    structs = {"reach_setexpr": _lib.reach_setexpr, "reach_invariant": _lib.reach_invariant,
               "reach_lgg09_opts": _lib.reach_lgg09_opts, "reach_term": _lib.reach_term}
    if stars == 0:
        return scalars[base]
    if base == "char" and stars == 1:
        return {ctypes.c_char_p}
    if base in structs:
        return {ctypes.POINTER(structs[base])}
    if base in scalars and stars == 1:
        return {ctypes.POINTER(t) for t in scalars[base]} | {ctypes.c_void_p}
    if base == "void" or base.startswith("reach_"):                 # opaque handle / raw device pointer
        return {ctypes.c_void_p} if stars == 1 else {ctypes.POINTER(ctypes.c_void_p)}
    if base in scalars and stars == 2:
        return {ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.POINTER(next(iter(scalars[base]))))}
    raise AssertionError("unmapped C type %r" % ctype)


def test_ctypes_signatures_match_the_header_types():
    """Width and kind of every bound argument and return type against the C prototype: a c_int where the header has
    int64_t would pass the arity check and corrupt the call."""
    protos = _c_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, (res, argtypes) in _lib.SIGNATURES.items():
        cret, cparams = protos[name]
        if cret == "void":
            assert res is None, name
        elif cret == "char*":
            assert res is ctypes.c_char_p, name
        else:
            assert res in _ctypes_for(cret), (name, res, cret)
        for k, (a, ct) in enumerate(zip(argtypes, cparams)):
            assert a in _ctypes_for(ct), "%s argument %d: ctypes %s against C %s" % (name, k + 1, a, ct)
