"""The iso_c_binding layer (pixelflow_b200/fortran/pixelflow_gpu_mod.f90) against the C ABI (include/pixelflow_gpu.h).

No Fortran compiler exists in the image, so the module cannot be compiled; what CAN be done is to parse it — numpy's
f2py ships a Fortran 90 parser (crackfortran) — and hold every `bind(C)` interface to the C prototype it names:
argument count and order, by-value vs by-reference, C type of every argument and of the result, and the layout of
`type, bind(C) :: pf_config` field by field against `struct pf_config`.  A mismatch here would corrupt the call
stack of a real Fortran driver without any compiler noticing (interfaces are trusted, not checked, by the linker).
Also checked: the complete driver `ibm3_uniform_gpu.f90` calls only entry points the module declares, with the right
number of arguments.
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOD = os.path.join(ROOT, "pixelflow_b200", "fortran", "pixelflow_gpu_mod.f90")
DRV = os.path.join(ROOT, "pixelflow_b200", "fortran", "ibm3_uniform_gpu.f90")
HDR = os.path.join(ROOT, "include", "pixelflow_gpu.h")

cf = pytest.importorskip("numpy.f2py.crackfortran")


def _strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


@pytest.fixture(scope="module")
def header():
    """{name: (return type, [(type, name), ...])} and the struct's [(type, name, array length)]"""
    text = _strip_c_comments(open(HDR).read())
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[ \*]+)(pf_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = re.sub(r"\s+", " ", ret).replace(" *", "*")
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\s+", " ", a.strip())
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
                alist.append((mm.group(1).strip().replace(" *", "*").replace("* *", "**"), mm.group(2)))
        protos[name] = (ret, alist)
    body = re.search(r"typedef struct pf_config \{(.*?)\} pf_config;", text, re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = re.sub(r"\s+", " ", decl.strip())
        if not decl:
            continue
        mm = re.match(r"(const void \*|int|double)\s*(.*)$", decl)
        typ = mm.group(1).strip()
        for nm in mm.group(2).split(","):
            nm = nm.strip()
            arr = re.match(r"([A-Za-z_0-9]+)\[(\d+)\]$", nm)
            fields.append((typ, arr.group(1), int(arr.group(2))) if arr else (typ, nm, 0))
    return protos, fields


@pytest.fixture(scope="module")
def module():
    cf.verbose = 0
    cf.quiet = 1
    cwd = os.getcwd()
    try:
        blocks = cf.crackfortran([MOD])
    finally:
        os.chdir(cwd)
    mod = blocks[0]
    assert mod["block"] == "module" and mod["name"] == "pixelflow_gpu"
    return mod


def _bind_names():
    """Fortran procedure name -> C name from `bind(C, name="...")` (crackfortran drops the binding label)"""
    text = open(MOD).read()
    text = re.sub(r"&\s*\n\s*", " ", text)
    out = {}
    for m in re.finditer(r"(?:subroutine|function)\s+(\w+)\s*\([^)]*\)\s*bind\(C,\s*name=\"(\w+)\"\)", text, re.I):
        out[m.group(1).lower()] = m.group(2)
    return out


def _c_type_of(var, by_value_ok=True):
    """the C parameter type a Fortran dummy argument of an interoperable interface corresponds to"""
    ts = var["typespec"]
    value = "value" in var.get("attrspec", [])
    array = "dimension" in var
    if ts == "type":
        base = {"c_ptr": "void*", "pf_config": "pf_config"}[var["typename"]]
        if var["typename"] == "c_ptr":
            return "void*" if value else "void**"
        assert not value
        return base + "*"
    kind = var.get("kindselector", {}).get("kind")
    if ts == "character":
        kind = var.get("kindselector", {}).get("kind") or var.get("charselector", {}).get("kind")
        assert kind == "c_char" and array
        return "char*"
    base = {("integer", "c_int"): "int", ("integer", "c_size_t"): "size_t", ("integer", "c_long_long"): "long long",
            ("real", "c_double"): "double"}[(ts, kind)]
    if value:
        assert not array
        return base
    return base + "*"


def _normalise_c(t):
    """drop const and the opaque handle's name: `const pf_solver*` ~ `void*` (type(c_ptr), value)"""
    t = t.replace("const ", "").strip()
    t = t.replace("pf_solver", "void")
    return t.replace(" *", "*")


def test_every_interface_matches_its_c_prototype(header, module):
    protos, _ = header
    names = _bind_names()
    iface = [b for b in module["body"] if b["block"] == "interface"]
    routines = [r for i in iface for r in i["body"]]
    assert len(routines) >= 31
    for r in routines:
        cname = names.get(r["name"])
        assert cname == r["name"], f"{r['name']}: bind(C) label {cname}"
        assert cname in protos, f"{cname} is not declared in include/pixelflow_gpu.h"
        ret, cargs = protos[cname]
        assert len(r["args"]) == len(cargs), f"{cname}: {len(r['args'])} Fortran vs {len(cargs)} C arguments"
        for fa, (ctype, cn) in zip(r["args"], cargs):
            got = _c_type_of(r["vars"][fa])
            want = _normalise_c(ctype)
            assert got == want, f"{cname}({cn}): Fortran `{fa}` is {got}, C says {ctype}"
        if r["block"] == "subroutine":
            assert ret == "void", cname
        else:
            res = r["vars"][r.get("result", r["name"])]
            if res["typespec"] == "type":
                assert res["typename"] == "c_ptr" and ret.endswith("*"), cname        # const char* as type(c_ptr)
            else:
                want = {"c_int": "int", "c_size_t": "size_t"}[res["kindselector"]["kind"]]
                assert ret == want, f"{cname}: result {want} vs {ret}"


def test_pf_config_type_mirrors_the_struct(header, module):
    _, fields = header
    typ = [b for b in module["body"] if b["block"] == "type" and b["name"] == "pf_config"][0]
    names = [n.replace("rank_bn", "rank") for n in typ["varnames"]]       # f2py renames `rank` (an intrinsic)
    assert names == [f[1].lower() for f in fields], "field order"
    for (ctype, cname, carr), fname in zip(fields, typ["varnames"]):
        v = typ["vars"][fname]
        if ctype == "const void *":
            assert v["typespec"] == "type" and v["typename"] == "c_ptr", cname
            continue
        want = {"int": ("integer", "c_int"), "double": ("real", "c_double")}[ctype]
        assert (v["typespec"], v["kindselector"]["kind"]) == want, cname
        dims = [int(d) for d in v.get("dimension", [])]
        assert dims == ([carr] if carr else []), cname
    src = open(MOD).read()
    assert re.search(r"type,\s*bind\(C\),\s*public\s*::\s*pf_config", src), "the type must be bind(C)"


def test_driver_calls_only_declared_entry_points(header, module):
    protos, _ = header
    iface = {r["name"]: r for i in module["body"] if i["block"] == "interface" for r in i["body"]}
    helpers = {b["name"]: b for b in module["body"] if b["block"] in ("function", "subroutine")}
    text = re.sub(r"!.*", "", open(DRV).read())
    text = re.sub(r"&\s*\n\s*", " ", text)
    used = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text, re.I))
    assert {"pf_create", "pf_set_porosity", "pf_upload", "pf_initial_conditions", "pf_step", "pf_gather",
            "pf_destroy", "pf_ranks_launch", "pf_ranks_count", "pf_ranks_unique_id", "pf_ranks_finish"} <= {
                u.lower() for u in used}
    for name in used:
        name = name.lower()
        assert name in iface or name in helpers, f"ibm3_uniform_gpu.f90 calls {name}, which the module does not declare"
    # argument counts of the calls
    for m in re.finditer(r"\b(pf_[a-z0-9_]+)\s*\(", text, re.I):
        name = m.group(1).lower()
        depth, i, nargs, cur = 1, m.end(), 0, False
        while depth:
            ch = text[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            if depth and not ch.isspace():
                cur = True
            i += 1
        nargs = nargs + 1 if cur else 0
        want = len((iface.get(name) or helpers[name])["args"])
        assert nargs == want, f"{name}: called with {nargs} arguments, declared with {want}"


def test_fortran_sources_respect_the_free_form_line_length():
    """132 characters (F2008 3.3.2.1): gfortran rejects longer code lines without -ffree-line-length-none"""
    d = os.path.dirname(MOD)
    for name in sorted(os.listdir(d)):
        if name.endswith(".f90"):
            for no, line in enumerate(open(os.path.join(d, name)), 1):
                assert len(line.rstrip("\n")) <= 132, f"{name}:{no}"


def test_every_driver_calls_only_declared_entry_points(module):
    import glob
    iface = {r["name"] for i in module["body"] if i["block"] == "interface" for r in i["body"]}
    helpers = {b["name"] for b in module["body"] if b["block"] in ("function", "subroutine")}
    drivers = sorted(glob.glob(os.path.join(os.path.dirname(MOD), "ibm*_gpu.f90")))
    assert len(drivers) == 5
    for path in drivers:
        text = re.sub(r"!.*", "", open(path).read())
        for name in set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text, re.I)):
            assert name.lower() in iface | helpers, f"{os.path.basename(path)} calls {name}"
