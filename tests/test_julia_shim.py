"""Static checks of the Julia host shim (julia-ocean-modelling_b200/julia/src).

There is no Julia toolchain in the image, so the shim cannot be executed here.  What can be
checked without one, and is: every `ccall` names a symbol declared in include/qgb200.h with the
same arity and the same argument / return type classes; `QGParams` lists the fields of
`struct qg_params` in the same order with matching types; in load order (includes expanded)
every type is defined before the first `ccall` signature that names it; no docstring is left
dangling in front of another docstring (Julia: "cannot document the following expression")."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "julia-ocean-modelling_b200", "julia", "src")
HEADER = os.path.join(ROOT, "include", "qgb200.h")


def c_class(t):
    t = t.strip()
    if "*" in t:
        return "cstr" if re.fullmatch(r"const\s+char\s*\*", t) else "ptr"
    t = t.replace("const", "").strip()
    return {"int": "int", "int32_t": "int", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "void": "void"}[t]


def jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")):
        return "ptr"
    return {"Cint": "int", "Int32": "int", "Int64": "i64", "UInt64": "u64", "Cdouble": "f64", "Float64": "f64",
            "Cstring": "cstr", "Cvoid": "void"}[t]


def header_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\n\s*([A-Za-z_][\w\s]*?[\w\*])\s*\**\s*\b(qg_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", src):
        full_ret = ret + ("*" if re.search(r"\*\s*" + name, src) else "")
        argl = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
        classes = []
        for a in argl:
            # drop the parameter name: the last identifier that is not part of the type
            m = re.match(r"(.*?)([A-Za-z_]\w*)?$", a)
            typ = m.group(1).strip() if m.group(2) and m.group(1).strip() else a
            classes.append(c_class(typ))
        protos[name] = (c_class(full_ret), classes)
    return protos


def flatten(path, seen=None):
    """Source text in load order with include(...) expanded; returns a list of (file, text) chunks."""
    seen = seen if seen is not None else set()
    if path in seen:
        return []
    seen.add(path)
    text = open(path).read()
    out, pos = [], 0
    for m in re.finditer(r'^include\("([^"]+)"\)', text, flags=re.M):
        out.append((path, text[pos:m.start()]))
        out += flatten(os.path.normpath(os.path.join(os.path.dirname(path), m.group(1))), seen)
        pos = m.end()
    out.append((path, text[pos:]))
    return out


def strip_jl_comments(text):
    return re.sub(r"#[^\n]*", "", text)


CCALL = re.compile(r"ccall\(\(:(\w+),\s*libqgb200\),\s*(\w+),\s*\(([^)]*)\)")


def all_ccalls():
    calls = []
    for root, _, files in os.walk(JL):
        for f in files:
            if f.endswith(".jl"):
                src = strip_jl_comments(open(os.path.join(root, f)).read())
                for name, ret, args in CCALL.findall(src):
                    types = [a.strip() for a in args.split(",") if a.strip()]
                    calls.append((f, name, ret, types))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    assert len(protos) >= 25 and "qg_create" in protos and protos["qg_last_error"][0] == "cstr"
    calls = all_ccalls()
    assert len(calls) >= 14
    used = set()
    for f, name, ret, types in calls:
        assert name in protos, f"{f}: ccall to {name}, which include/qgb200.h does not declare"
        cret, cargs = protos[name]
        assert jl_class(ret) == cret, f"{f}: {name} returns {ret}, header says {cret}"
        assert len(types) == len(cargs), f"{f}: {name} called with {len(types)} argument types, header has {len(cargs)}"
        for k, (jt, ct) in enumerate(zip(types, cargs)):
            assert jl_class(jt) == ct, f"{f}: {name} argument {k}: {jt} vs {ct}"
        used.add(name)
    # the step path and the state transfers must all be bound
    for must in ("qg_create", "qg_destroy", "qg_upload_state", "qg_upload_initial_state", "qg_download_state", "qg_step",
                 "qg_evolve_zeta", "qg_evolve_psi", "qg_solve", "qg_init_state", "qg_snapshot_begin", "qg_snapshot_end",
                 "qg_extrema", "qg_diagnostics", "qg_last_error"):
        assert must in used, f"the Julia shim never calls {must}"


def test_qgparams_mirrors_struct_qg_params():
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct qg_params \{(.*?)\} qg_params;", hdr, flags=re.S).group(1)
    cfields = []
    for typ, names in re.findall(r"\b(int32_t|double)\s+([^;]+);", body):
        for n in names.split(","):
            n = n.strip()
            m = re.match(r"(\w+)\[(\d+)\]", n)
            cfields.append((m.group(1), typ, int(m.group(2))) if m else (n, typ, 1))
    jl = strip_jl_comments(open(os.path.join(JL, "model.jl")).read())
    sbody = re.search(r"struct QGParams\n(.*?)\nend", jl, flags=re.S).group(1)
    jfields = []
    for name, typ in re.findall(r"^\s*(\w+)::([\w\{\},]+)", sbody, flags=re.M):
        m = re.match(r"NTuple\{(\d+),Float64\}", typ)
        jfields.append((name, "double", int(m.group(1))) if m else (name, {"Int32": "int32_t", "Float64": "double"}[typ], 1))
    assert jfields == cfields


def test_types_are_defined_before_the_ccalls_that_name_them():
    for entry in ("model.jl", "run_model.jl", "run_model_no_output.jl"):
        text = "".join(strip_jl_comments(t) for _, t in flatten(os.path.join(JL, entry)))
        for typ in ("QGParams",):
            first_use = min((m.start() for m in CCALL.finditer(text) if typ in m.group(3)), default=None)
            assert first_use is not None or entry != "model.jl"
            if first_use is not None:
                d = re.search(r"^struct " + typ + r"\b", text, flags=re.M)
                assert d and d.start() < first_use, f"{entry}: {typ} is used in a ccall signature before it is defined"
        for fn in ("qg_check", "qg_error", "libqgb200"):
            d = re.search(r"^(function |const )?" + fn + r"\b[^\n]*=|^function " + fn + r"\b", text, flags=re.M)
            u = re.search(r"\b" + fn + r"\b", text)
            assert d, f"{entry}: {fn} is never defined"
        # single definition of every struct (a second include of model.jl would redefine them)
        for s in ("QGParams", "BaroclinicModel", "RectangularDomain", "SpectralPlan"):
            assert len(re.findall(r"^(?:mutable )?struct " + s + r"\b", text, flags=re.M)) == 1, (entry, s)


def test_no_dangling_docstrings():
    for root, _, files in os.walk(JL):
        for f in files:
            if not f.endswith(".jl"):
                continue
            src = open(os.path.join(root, f)).read()
            for m in re.finditer(r'"""(?:.|\n)*?"""', src):
                tail = src[m.end():].lstrip()
                assert not tail.startswith('"""'), f"{f}: two docstrings in a row near offset {m.start()}"
                assert tail, f"{f}: docstring at end of file"
