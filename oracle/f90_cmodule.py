"""Description of an iso_c_binding Fortran module for oracle/f90toc.py.  TEST INFRASTRUCTURE.

`pixelflow_b200/fortran/pixelflow_gpu_mod.f90` (the product's Fortran binding layer) is parsed with numpy's f2py
parser (crackfortran); out come its interoperable derived types (field order, C types), its integer named constants
and its `bind(C, name=...)` interfaces (argument by value / by reference, C types, result type) — exactly the facts a
Fortran compiler takes from the module when it compiles a `use pixelflow_gpu` program.  f90toc then translates such
a program (pixelflow_b200/fortran/ibm3_uniform_gpu.f90) to C that calls the C ABI the way the Fortran would.

The two contained helper procedures `pf_check` and `pf_error_message` use allocatable deferred-length strings,
c_f_pointer and transfer(); they are outside the translator's subset and are supplied as C by hand
(oracle/fortran_helpers.c) — error path only, no numerics.
"""
from __future__ import annotations

import os
import re


def _ctype(var):
    ts = var["typespec"]
    if ts == "type":
        tn = var["typename"]
        return "cptr" if tn == "c_ptr" else "type:" + tn
    if ts == "character":
        return "char"
    kind = var.get("kindselector", {}).get("kind")
    return {("integer", "c_int"): "int", ("integer", "c_size_t"): "size_t", ("integer", "c_long_long"): "longlong",
            ("real", "c_double"): "real"}[(ts, kind)]


def describe(path: str) -> dict:
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    cf.quiet = 1
    cwd = os.getcwd()
    import contextlib
    import io
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mod = cf.crackfortran([path])[0]
    finally:
        os.chdir(cwd)
    text = re.sub(r"&\s*\n\s*", " ", open(path).read())
    labels = {m.group(1).lower(): m.group(2) for m in
              re.finditer(r"(?:subroutine|function)\s+(\w+)\s*\([^)]*\)\s*bind\(C,\s*name=\"(\w+)\"\)", text, re.I)}
    desc = {"name": mod["name"], "src": path, "types": {}, "params": {}, "functions": {}}
    for b in mod["body"]:
        if b["block"] == "type":
            fields = []
            for fname in b["varnames"]:
                v = b["vars"][fname]
                real_name = "rank" if fname == "rank_bn" else fname      # f2py renames `rank` (an intrinsic's name)
                dims = [int(d) for d in v.get("dimension", [])]
                fields.append((real_name, _ctype(v), dims[0] if dims else 0))
            desc["types"][b["name"]] = fields
        elif b["block"] == "interface":
            for r in b["body"]:
                args = []
                for a in r["args"]:
                    v = r["vars"][a]
                    how = "value" if "value" in v.get("attrspec", []) else ("array" if "dimension" in v else "ref")
                    args.append((_ctype(v), how))
                if r["block"] == "subroutine":
                    ret = "void"
                else:
                    ret = _ctype(r["vars"][r.get("result", r["name"])])
                desc["functions"][r["name"]] = {"ret": ret, "args": args, "cname": labels[r["name"]]}
    # named integer constants of the module specification part
    for name, v in mod.get("vars", {}).items():
        if "parameter" in v.get("attrspec", []) and v.get("typespec") == "integer" and "=" in v:
            try:
                desc["params"][name.lower()] = int(v["="])
            except ValueError:
                pass
    # hand-written helpers (oracle/fortran_helpers.c)
    desc["functions"]["pf_check"] = {"ret": "void", "cname": "ft_pf_check",
                                     "args": [("int", "value"), ("cptr", "value"), ("cstr", "value")]}
    desc["functions"]["pf_error_message"] = {"ret": "cstr", "cname": "ft_pf_error_message", "args": [("cptr", "value")]}
    return desc
