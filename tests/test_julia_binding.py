"""The Julia host (julia/NoNameB200.jl) cannot run in this image (no julia binary): check statically that every
`ccall` in it binds an entry point of include/nnpm.h with the right arity and argument classes, that every declared
entry point is bound, and that its NNPMConfig / NNPMStats structs list the header's fields in order with matching
types."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nnpm.h")
JULIA = os.path.join(ROOT, "julia", "NoNameB200.jl")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def _c_class(t):
    t = t.strip()
    if "*" in t or "[" in t:
        return "cstr" if re.match(r"^const\s+char\s*\*", t) else "ptr"
    base = re.sub(r"\b(const|unsigned)\b", "", t).split()[0]
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "float": "f32", "void": "void"}[base]


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cdouble": "f64",
            "Float32": "f32", "Cstring": "cstr"}[t]


def header_prototypes():
    h = _strip_c_comments(open(HEADER, encoding="utf-8").read())
    protos = {}
    for ret, name, args in re.findall(r"\b(int64_t|int|const char \*)\s*(nnpm_\w+)\s*\(([^;{]*?)\)\s*;", h):
        args = args.strip()
        classes = [] if args in ("", "void") else [_c_class(a) for a in args.split(",")]
        protos[name] = (_c_class(ret), classes)
    return protos


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def julia_ccalls():
    j = open(JULIA, encoding="utf-8").read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(nnpm_\w+),\s*libnnpm\),\s*(\w+),\s*\(", j):
        # argument-type tuple: balanced parentheses from m.end() - 1
        i, depth = m.end() - 1, 0
        while True:
            depth += j[i] == "("
            depth -= j[i] == ")"
            if depth == 0:
                break
            i += 1
        types = [t for t in _split_top(j[m.end():i]) if t.strip()]
        # values passed: the rest of the ccall's argument list
        k, depth = i + 1, 1          # we are inside ccall( ... ; depth 1
        start = k
        while depth:
            depth += j[k] == "("
            depth -= j[k] == ")"
            k += 1
        vals = [v for v in _split_top(j[start:k - 1]) if v.strip()]
        calls.append((m.group(1), _jl_class(m.group(2)), [_jl_class(t) for t in types], len(vals)))
    return calls


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= len(protos)
    for name, ret, types, nvals in calls:
        assert name in protos, f"{name} is not declared in nnpm.h"
        cret, cargs = protos[name]
        want_ret = {"cstr": "cstr"}.get(cret, cret)
        assert ret == want_ret, f"{name}: Julia returns {ret}, header {cret}"
        got = ["ptr" if t == "cstr" and c == "ptr" else t for t, c in zip(types, cargs)]
        want = ["cstr" if t == "cstr" else c for t, c in zip(types, cargs)]      # const char* parameters bind as Cstring
        assert len(types) == len(cargs), f"{name}: {len(types)} argument types, header has {len(cargs)}"
        assert got == want, f"{name}: Julia {types} vs header {cargs}"
        assert nvals == len(types), f"{name}: {nvals} values for {len(types)} argument types"


def test_every_entry_point_is_bound():
    bound = {c[0] for c in julia_ccalls()}
    assert bound == set(header_prototypes()), sorted(set(header_prototypes()) ^ bound)


def _header_struct(name):
    h = _strip_c_comments(open(HEADER, encoding="utf-8").read())
    end = re.search(r"\}\s*" + name + r"\s*;", h).start()
    body = h[h.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"^(.*?)(\w+)\s*(\[(\w+)\])?$", decl)
        typ, fname, dim = m.group(1).strip(), m.group(2), m.group(4)
        if dim and not dim.isdigit():
            dim = re.search(r"\b" + dim + r"\s*=\s*(\d+)", h).group(1)
        fields.append((fname, "ptr" if "*" in typ else _c_class(typ), int(dim) if dim else 0))
    return fields


def _julia_struct(name):
    j = open(JULIA, encoding="utf-8").read()
    body = re.search(r"struct " + name + r"\n(.*?)\nend", j, flags=re.S).group(1)
    fields = []
    for line in body.splitlines():
        fname, typ = [q.strip() for q in line.split("::")]
        m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", typ)
        fields.append((fname, _jl_class(m.group(2)) if m else _jl_class(typ), int(m.group(1)) if m else 0))
    return fields


def test_struct_layouts_match_the_header():
    assert _julia_struct("NNPMConfig") == _header_struct("nnpm_config")
    assert _julia_struct("NNPMStats") == _header_struct("nnpm_stats")
    j = open(JULIA, encoding="utf-8").read()
    h = open(HEADER, encoding="utf-8").read()
    ver = re.search(r"#define NNPM_VERSION (\d+)", h).group(1)
    assert f"Cint, ()) == {ver}" in j              # the ABI version the binding was written against


def test_documents_cite_an_existing_julia_file():
    for doc in ("README.md", "DESIGN.md", "INTEGRATION.md", os.path.join("include", "nnpm.h"), os.path.join("noname_b200", "_lib.py")):
        txt = open(os.path.join(ROOT, doc), encoding="utf-8").read()
        if "julia/NoNameB200.jl" in txt:
            assert os.path.isfile(JULIA)


def _julia_code_only(src):
    """the source with comments, string and character literals blanked (enough of a lexer for the checks below)"""
    out, i, n = [], 0, len(src)
    while i < n:
        ch = src[i]
        if ch == "#":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if ch == '"':
            if src.startswith('"""', i):
                i = src.index('"""', i + 3) + 3
            else:
                j = i + 1
                while src[j] != '"':
                    j += 2 if src[j] == "\\" else 1
                i = j + 1
            out.append('""')
            continue
        if ch == "'" and i + 2 < n and (src[i + 2] == "'" or (src[i + 1] == "\\" and src[i + 3] == "'")):
            i += 3 if src[i + 2] == "'" else 4
            out.append("' '")
            continue
        out.append(ch)
        i += 1
    return "".join(out)


def test_julia_blocks_and_brackets_balance():
    """No julia binary here: at least every function / if / for / while / do / struct / module / try block is closed by
    its `end` (an `end` inside [] or () is an index, a for / if inside them a comprehension) and brackets pair up."""
    code = _julia_code_only(open(JULIA, encoding="utf-8").read())
    stack, br, line = [], [], 1
    for m in re.finditer(r"[\[\](){}]|\b(function|if|for|while|do|struct|module|begin|let|try|quote|end)\b|\n", code):
        t = m.group(0)
        if t == "\n":
            line += 1
        elif t in "[({":
            br.append((t, line))
        elif t in "])}":
            assert br and "[({".index(br[-1][0]) == "])}".index(t), f"line {line}: unmatched {t}"
            br.pop()
        elif t == "end":
            if not br:
                assert stack, f"line {line}: `end` without an open block"
                stack.pop()
        elif not (br and t in ("for", "if")):
            stack.append((t, line))
    assert not stack and not br, f"unclosed: {stack} {br}"


def test_julia_host_covers_the_reference_surface():
    """names a NoName user calls (src/NoName.jl:8, src/ParticleMesh.jl operator list) are defined in the module"""
    code = _julia_code_only(open(JULIA, encoding="utf-8").read())
    for name in ("run_noname", "noname", "restart", "evolvePM", "evolvePMCosmology", "InitializeParticleSimulation", "analyzePM",
                 "readyForOutput", "write_all", "deposit", "compute_potential", "compute_acceleration", "interpVecToPoints",
                 "drift", "kick", "make_periodic", "powerSpectrum", "UniformParticles", "UniformParticleWind",
                 "ZeldovichPancakeParticles", "PowerSpectrumParticles", "a_from_t_internal", "dadt_from_t_internal",
                 "z_from_t_internal", "hdf5_writer", "device_for", "sync_host!"):
        assert re.search(rf"(^|\n)\s*(function\s+)?{re.escape(name)}\(", code), name


def test_julia_file_lexes_cleanly():
    """Pygments' Julia lexer finds no malformed token (it does not know regex literals: backslashes inside r"..." are
    the only tokens it may flag)"""
    pygments = __import__("pytest").importorskip("pygments")
    from pygments.lexers import JuliaLexer
    from pygments.token import Error
    src = open(JULIA, encoding="utf-8").read()
    lines, line = src.splitlines(), 1
    for tok, val in JuliaLexer().get_tokens(src):
        if tok in Error:
            assert val == "\\" and 'r"' in lines[line - 1], (line, val, lines[line - 1])
        line += val.count("\n")


def test_julia_calls_resolve_to_definitions_or_base():
    """every name called as f(...) is defined in the file or is one of the Base / stdlib functions listed here (catches
    typos that only a Julia run would otherwise find)"""
    base = set("""Dict Float64 IOBuffer Int Int32 Int64 String Tuple UInt64 UInt8 abs asinh copy count div dropdims eachline
        endswith enumerate fill first floor get get! haskey joinpath match max maximum min mkpath mod occursin ones parse prod
        replace round serialize deserialize similar sinh size sizeof split sprintf sqrt startswith string strip sum take!
        tryparse unsafe_string zeros zip error isdir isfile isempty length merge! open println push! write finalizer ccall
        Ref ndims collect findall sin cos exp log""".split())
    code = _julia_code_only(open(JULIA, encoding="utf-8").read())
    defs = set(re.findall(r"(?:^|\n)\s*(?:function\s+)?([A-Za-z_][\w!]*)\(", code))
    defs |= set(re.findall(r"\bstruct\s+(\w+)", code)) | set(re.findall(r"\bconst\s+(\w+)", code))
    calls = set(re.findall(r"(?<![\w.:@])([A-Za-z_][\w!]*)\(", code))
    unknown = sorted(c for c in calls if c not in defs and c not in base)
    assert not unknown, unknown


def test_julia_defaults_equal_the_python_defaults():
    """both hosts restate src/default.conf: the Julia Dict literal must hold the same keys and values"""
    import sys
    sys.path.insert(0, ROOT)
    from noname_b200 import conf as cf
    src = open(JULIA, encoding="utf-8").read()
    body = src[src.index("const DEFAULTS = Dict{String,Any}("):]
    body = body[:body.index("\nfunction load_conf")]
    pairs = re.findall(r'"([^"]+)"\s*=>\s*(\[[^\]]*\]|"[^"]*"|[^,\)\n]+)', body)
    got = {k: cf.parse_value(v.strip()) for k, v in pairs}
    assert got == cf.DEFAULTS
