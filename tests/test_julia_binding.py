"""julia/TNQSCuda.jl cannot run here (no Julia in the image); what can be checked mechanically is: it binds every entry
point of include/tnqs_cuda.h with `ccall`, with the header's return type, argument count and argument types, and its option
structs have the header's fields in the header's order (and the ctypes twin in lib.py agrees)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def header_prototypes():
    text = open(os.path.join(ROOT, "include", "tnqs_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(tnqs_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        protos[name] = (ret, [] if args == "void" else _split_top(args))
    return protos


def c_to_julia(decl):
    """Acceptable Julia ccall types for one C parameter declaration (or return type)."""
    d = decl.replace("const ", "").strip()
    stars = d.count("*")
    base = re.sub(r"[\*\s]+[A-Za-z_][A-Za-z0-9_]*$", "", d).replace("*", "").strip() if " " in d or "*" in d else d
    base = base.split()[0] if base else d.split()[0]
    scalar = {"int32_t": "Int32", "int64_t": "Int64", "uint64_t": "UInt64", "double": "Float64"}
    if stars == 0:
        return {scalar[base]}
    if base == "tnqs_ctx":
        return {"Ctx"} if stars == 1 else {"Ref{Ctx}"}
    if base == "tnqs_bp_opts":
        return {"Ref{BpOpts}"}
    if base == "tnqs_apply_opts":
        return {"Ref{ApplyOpts}"}
    if base == "char":
        return {"Cstring"}
    if base == "void":
        return {"Ptr{Cvoid}", "Ptr{UInt8}"} if stars == 1 else {"Ref{Ptr{Cvoid}}"}
    if base == "double":
        if stars == 2:
            return {"Ptr{Ptr{ComplexF64}}"}
        return {"Ptr{Float64}", "Ptr{ComplexF64}", "Ref{Float64}"}
    j = scalar[base]
    return {f"Ptr{{{j}}}", f"Ref{{{j}}}"}


def julia_ccalls():
    text = open(os.path.join(ROOT, "julia", "TNQSCuda.jl")).read()
    text = re.sub(r"#[^\n]*", "", text)
    calls = {}
    for m in re.finditer(r"ccall\(\(:(tnqs_[a-z0-9_]+), LIB\),\s*([A-Za-z0-9_{}]+),\s*\(", text):
        start = m.end()
        depth, i = 1, start
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        types = _split_top(" ".join(text[start:i - 1].split()))
        types = [t for t in types if t]
        calls.setdefault(m.group(1), []).append((m.group(2), types))
    return calls


def test_julia_module_binds_every_header_symbol_with_matching_types():
    protos, calls = header_prototypes(), julia_ccalls()
    assert len(protos) >= 45
    assert sorted(calls) == sorted(protos)
    for name, (ret, args) in protos.items():
        for jret, jargs in calls[name]:
            assert jret in c_to_julia(ret) or (ret.endswith("*") and "void" in ret and jret == "Ptr{Cvoid}"), (name, ret, jret)
            assert len(jargs) == len(args), (name, args, jargs)
            for a, ja in zip(args, jargs):
                assert ja in c_to_julia(a), (name, a, ja)


def test_option_structs_match_header_and_ctypes():
    header = open(os.path.join(ROOT, "include", "tnqs_cuda.h")).read()
    julia = open(os.path.join(ROOT, "julia", "TNQSCuda.jl")).read()
    from tensornetworkquantumsimulator_b200 import lib
    ctype = {"int32_t": "Int32", "double": "Float64"}
    for cname, jname, pyt in (("tnqs_bp_opts", "BpOpts", lib.BpOpts), ("tnqs_apply_opts", "ApplyOpts", lib.ApplyOpts)):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname, header, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        cfields = [(ctype[t], n) for t, n in re.findall(r"(int32_t|double)\s+([a-z_]+)\s*;", body)]
        jbody = re.search(r"struct " + jname + r"\b(.*?)\nend", julia, flags=re.S).group(1)
        jfields = [(t, n) for n, t in re.findall(r"^\s*([a-z_]+)::([A-Za-z0-9]+)", jbody, flags=re.M)]
        assert jfields == cfields, (cname, cfields, jfields)
        assert [n for n, _ in pyt._fields_] == [n for _, n in cfields]


def test_julia_module_mirrors_the_python_host_surface():
    """Every reference-facing function of the executable host (engine.py) that the Julia module claims to cover exists there."""
    julia = open(os.path.join(ROOT, "julia", "TNQSCuda.jl")).read()
    from tensornetworkquantumsimulator_b200 import engine
    for fn in ("update", "apply_gates", "expect", "rescale", "symmetric_gauge", "truncate", "entanglement", "inner", "network",
               "message", "messages", "partitionfunction", "maxvirtualdim"):
        assert hasattr(engine, fn), fn
        assert re.search(r"\b" + fn + r"[!]?\(", julia), fn
