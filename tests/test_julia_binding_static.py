"""The Julia binding (thunderbolt.jl_b200/julia/ThunderboltB200Ext.jl) cannot run here -- there is no Julia in the image --
so check statically what can be checked: every `ccall` / `@tb` names an entry point include/tbolt_b200.h declares, passes
as many arguments as the C prototype has, and each Julia argument type is one the C parameter type accepts (Float64 for
double, Int32 / Int64 for the sized integers, pointers for pointers).  Catches the typos an unexecuted file collects."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
JL = (ROOT / "thunderbolt.jl_b200" / "julia" / "ThunderboltB200Ext.jl").read_text()
HDR = (ROOT / "include" / "tbolt_b200.h").read_text()


def c_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int32_t|int64_t|const char \*)\s*(tb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = params
    return protos


def split_top(s):
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


def julia_calls():
    calls = []
    for m in re.finditer(r"@tb\s+(tb_[a-z0-9_]+)\s+\(", JL):
        start = m.end() - 1
        depth, i = 0, start
        while True:
            depth += JL[i] == "("
            depth -= JL[i] == ")"
            if depth == 0:
                break
            i += 1
        calls.append((m.group(1), split_top(JL[start + 1:i])))
    for m in re.finditer(r"ccall\(\(:(tb_[a-z0-9_]+),\s*LIB\[\]\),\s*(\w+),\s*\(", JL):
        start = m.end() - 1
        depth, i = 0, start
        while True:
            depth += JL[i] == "("
            depth -= JL[i] == ")"
            if depth == 0:
                break
            i += 1
        calls.append((m.group(1), split_top(JL[start + 1:i])))
    return calls


def compatible(jl: str, c: str) -> bool:
    c = c.replace("const ", "").strip()
    is_ptr = "*" in c
    base = c.split()[0] if not is_ptr else None
    if is_ptr:
        return jl.startswith(("Ptr{", "Ref{")) or jl == "Cstring"
    return {"double": jl == "Float64", "int32_t": jl == "Int32", "int64_t": jl == "Int64"}.get(base, False)


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_calls()
    assert len(protos) >= 100 and len(calls) >= 40
    for name, types in calls:
        assert name in protos, f"{name} is not declared in include/tbolt_b200.h"
        params = protos[name]
        assert len(types) == len(params), f"{name}: Julia passes {len(types)} arguments, the prototype has {len(params)}: {params}"
        for k, (jt, cp) in enumerate(zip(types, params)):
            assert compatible(jt, cp), f"{name}: argument {k} is {jt} in Julia but `{cp}` in C"


def test_binding_covers_the_path():
    """the entry points SURVEY 8b's hooks need are all bound"""
    bound = {n for n, _ in julia_calls()}
    for need in ("tb_ctx_create", "tb_vec_create", "tb_vec_upload", "tb_vec_download", "tb_mesh_create", "tb_csr_create",
                 "tb_assemble_mass", "tb_assemble_diffusion", "tb_assemble_source_program", "tb_assemble_source_qp",
                 "tb_csr_axpby_values", "tb_spmv", "tb_cg_solve_pc", "tb_cell_step", "tb_vec_axpy"):
        assert need in bound, need
