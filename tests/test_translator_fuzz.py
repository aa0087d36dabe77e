"""Random expressions through oracle/f90toc.py against an independent evaluator.

The translator's expression handling (precedence, left-to-right association, unary minus vs `**`, integer division
and mod truncating toward zero, mixed integer/real arithmetic, `real()`, `max/min/abs`) is what the golden vectors of
the translated reference rest on.  Here random expression TREES are generated, rendered as Fortran text with the
minimum of parentheses Fortran's grammar requires, translated to C, compiled and run; the same trees are evaluated
directly in Python (IEEE doubles, explicit truncating integer division) — no parser involved on that side.
"""
import ctypes as C
import math
import os
import random
import subprocess

import numpy as np

from oracle import build_ref, f90toc

# precedence levels of the Fortran grammar (higher binds tighter)
P_ADD, P_MUL, P_POW, P_ATOM = 1, 2, 3, 4
REALS = {"a": 1.7, "b": -0.3, "c": 2.25, "d": 1.0e-3, "e": 7.5}
INTS = {"i": 7, "j": -3, "k": 2, "n": 11}


class T:
    def __init__(self, op, kids=(), typ="real", val=None):
        self.op, self.kids, self.typ, self.val = op, kids, typ, val


def gen(rng, depth, typ):
    if depth == 0 or rng.random() < 0.2:
        if typ == "real":
            if rng.random() < 0.6:
                return T("var", typ="real", val=rng.choice(list(REALS)))
            if rng.random() < 0.5:
                return T("lit", typ="real", val=rng.choice(["0.5", "2.", "1.e-1", "3.25", "1.5d0"]))
            return T("realfn", (gen(rng, 0, "int"),), "real")
        if rng.random() < 0.7:
            return T("var", typ="int", val=rng.choice(list(INTS)))
        return T("lit", typ="int", val=rng.choice(["1", "2", "3", "5"]))
    r = rng.random()
    if typ == "real":
        if r < 0.55:
            op = rng.choice("+-*/")
            lt = "real" if rng.random() < 0.8 else "int"
            rt_ = "real" if (lt == "int" or rng.random() < 0.8) else "int"
            return T(op, (gen(rng, depth - 1, lt), gen(rng, depth - 1, rt_)), "real")
        if r < 0.65:
            return T("neg", (gen(rng, depth - 1, "real"),), "real")
        if r < 0.75:
            return T("pow2", (gen(rng, depth - 1, "real"),), "real")
        if r < 0.85:
            return T(rng.choice(["max", "min"]), (gen(rng, depth - 1, "real"), gen(rng, depth - 1, "real")), "real")
        if r < 0.92:
            return T("abs", (gen(rng, depth - 1, "real"),), "real")
        return T("paren", (gen(rng, depth - 1, "real"),), "real")
    if r < 0.6:
        op = rng.choice("+-*/")
        return T(op, (gen(rng, depth - 1, "int"), gen(rng, depth - 1, "int")), "int")
    if r < 0.75:
        return T("mod", (gen(rng, depth - 1, "int"), gen(rng, depth - 1, "int")), "int")
    if r < 0.85:
        return T("neg", (gen(rng, depth - 1, "int"),), "int")
    return T("paren", (gen(rng, depth - 1, "int"),), "int")


def render(t, first_in_term=True):
    """-> (text, precedence level of the text's outermost operator)"""
    if t.op in ("var", "lit"):
        return t.val, P_ATOM
    if t.op == "realfn":
        return f"real({render(t.kids[0])[0]})", P_ATOM
    if t.op in ("max", "min", "mod"):
        return f"{t.op}({render(t.kids[0])[0]}, {render(t.kids[1])[0]})", P_ATOM
    if t.op == "abs":
        return f"abs({render(t.kids[0])[0]})", P_ATOM
    if t.op == "paren":
        return f"({render(t.kids[0])[0]})", P_ATOM
    if t.op == "pow2":
        s, p = render(t.kids[0])
        return (s if p == P_ATOM else f"({s})") + "**2", P_POW
    if t.op == "neg":
        # unary minus has the precedence of binary minus and applies to the following TERM
        s, p = render(t.kids[0])
        return "-" + (s if p >= P_MUL else f"({s})"), P_ADD
    lv = P_ADD if t.op in "+-" else P_MUL
    ls, lp = render(t.kids[0])
    rs, rp = render(t.kids[1])
    # left operand: same level is fine (left association); a unary minus (level ADD) inside a product needs ()
    if lp < lv:
        ls = f"({ls})"
    # right operand: must bind tighter than this operator (a - (b - c), a / (b * c)); a leading sign needs () too
    if rp <= lv or rs.startswith("-"):
        rs = f"({rs})"
    return f"{ls} {t.op} {rs}", lv


class Bad(Exception):
    pass


def ev(t):
    if t.op == "var":
        return REALS[t.val] if t.typ == "real" else INTS[t.val]
    if t.op == "lit":
        return float(t.val.replace("d", "e")) if t.typ == "real" else int(t.val)
    k = [ev(x) for x in t.kids]
    if t.op == "realfn":
        return float(k[0])
    if t.op == "paren":
        return k[0]
    if t.op == "neg":
        return -k[0]
    if t.op == "pow2":
        return k[0] * k[0]
    if t.op == "abs":
        return abs(k[0])
    if t.op == "max":
        return k[0] if k[0] > k[1] else k[1]
    if t.op == "min":
        return k[0] if k[0] < k[1] else k[1]
    if t.typ == "int":
        a, b = k
        if t.op == "+":
            r = a + b
        elif t.op == "-":
            r = a - b
        elif t.op == "*":
            r = a * b
        else:
            if b == 0:
                raise Bad()
            q = abs(a) // abs(b)
            q = q if (a >= 0) == (b >= 0) else -q          # truncation toward zero
            r = q if t.op == "/" else a - q * b            # mod: sign of a
        if abs(r) > 2 ** 30:
            raise Bad()
        return r
    a, b = float(k[0]), float(k[1])
    if t.op == "+":
        r = a + b
    elif t.op == "-":
        r = a - b
    elif t.op == "*":
        r = a * b
    else:
        if b == 0.0:
            raise Bad()
        r = a / b
    if not math.isfinite(r) or abs(r) > 1e150:
        raise Bad()
    return r


def test_random_expressions(tmp_path):
    rng = random.Random(20240607)
    cases = []
    while len(cases) < 400:
        typ = "real" if rng.random() < 0.75 else "int"
        t = gen(rng, rng.randint(2, 5), typ)
        try:
            v = ev(t)
        except (Bad, OverflowError):
            continue
        cases.append((typ, render(t)[0], v))
    nr = sum(1 for c in cases if c[0] == "real")
    ni = len(cases) - nr
    lines = ["program main", "implicit none", f"real, dimension({nr}) :: r", f"integer, dimension({ni}) :: q",
             "real :: " + ", ".join(REALS), "integer :: " + ", ".join(INTS)]
    lines += [f"{k} = {v!r}" for k, v in REALS.items()] + [f"{k} = {v}" for k, v in INTS.items()]
    ir = iq = 0
    for typ, text, _ in cases:
        if typ == "real":
            ir += 1
            lines.append(f"r({ir}) = {text}")
        else:
            iq += 1
            lines.append(f"q({iq}) = {text}")
    lines.append("end program main")
    tr = f90toc.Translator()
    tr.add_source("\n".join(lines) + "\n", "fuzz.f90")
    (tmp_path / "fuzz.c").write_text(tr.emit())
    so = tmp_path / "fuzz.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", build_ref.HERE,
                           str(tmp_path / "fuzz.c"), os.path.join(build_ref.HERE, "ref_runtime.c"), "-o", str(so),
                           "-lm", "-ldl"])
    from oracle.ref_translated import RtVar
    L = C.CDLL(str(so))
    L.ref_run.argtypes = [C.c_char_p]
    L.ref_lookup.argtypes = [C.c_char_p]
    L.ref_lookup.restype = C.POINTER(RtVar)
    assert L.ref_run(str(tmp_path).encode()) == 0
    r = np.ctypeslib.as_array(C.cast(L.ref_lookup(b"r").contents.ptr, C.POINTER(C.c_double)), (nr,))
    q = np.ctypeslib.as_array(C.cast(L.ref_lookup(b"q").contents.ptr, C.POINTER(C.c_int)), (ni,))
    ir = iq = 0
    for typ, text, v in cases:
        if typ == "real":
            assert r[ir] == v, f"{text}: translated {r[ir]!r}, expected {v!r}"
            ir += 1
        else:
            assert q[iq] == v, f"{text}: translated {q[iq]}, expected {v}"
            iq += 1
