"""The hand-written oracle (oracle/pf_oracle.c) against the REFERENCE ITSELF.

The reference is Fortran and no Fortran compiler exists in this image; oracle/f90toc.py translates the reference's
own source files mechanically into C (oracle/build_ref.py -> oracle/_ref/*.so, git-ignored), and the translated
`program main` is run on project directories like the reference is.  Three layers, CPU only:

  1. the translator on small Fortran snippets with hand-computed answers (precedence, integer division, do/if
     forms, array bounds, by-reference arguments, static locals) — the translator is the new trusted component;
  2. the oracle against the golden vectors the translated reference produced (tests/golden/ref_translated.npz,
     made by tests/golden/make_ref_translated.py) — runs anywhere, also without oracle/_ref;
  3. live, where oracle/_ref exists (here; on the GPU box the prebuilt libraries travel with the snapshot): the
     five translated programs re-run on seeded random decks and on the shipped decks, serial and OpenMP flavours,
     bit-compared with the oracle.
Bar: bit-exact (np.array_equal) on u, v, w, p, porosity incl. halos and on the logged 'p error'.
"""
import hashlib
import json
import os
import re
import subprocess
import tempfile
import zlib

import numpy as np
import pytest

from oracle import build_ref, f90toc
from oracle import ref_translated as rt

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref_translated.npz")

needs_ref = pytest.mark.skipif(not (build_ref.available() or os.path.isdir(build_ref.OUT)),
                               reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if build_ref.available():
        build_ref.build()


# ------------------------------------------------------------------------------------------------ 1. translator

SNIPPET = """
module consts
  implicit none
  integer, parameter:: md = 6, nd = 4
end module consts

program main
  use consts
  implicit none
  real,dimension(0:md,0:nd):: a
  real,dimension(3):: r
  integer,dimension(8):: iv
  real:: x, y, s
  integer:: i, j, k, cnt
  logical:: flag
  x = 7.
  y = 2.
  ! precedence and association: left-to-right * /, unary minus below **, parentheses kept
  r(1) = x/y*4. - 3.*y**2
  r(2) = -x**2 + (x - y)/y/y
  r(3) = 1.e-1 + 2.5d0*real(7/2) + mod(x, y)
  ! integer arithmetic truncates toward zero
  iv(1) = 7/2
  iv(2) = (-7)/2
  iv(3) = mod(-7, 3)
  iv(4) = 2**5 + max(3, 9, 4) - min(3, 9) + abs(-4)
  ! do with a step, loop variable after the loop, one-line if, else-if chain
  cnt = 0
  do k = 2, 11, 3
    cnt = cnt + k
  end do
  iv(5) = cnt
  iv(6) = k
  do j = 0, nd
    do i = 0, md
      a(i,j) = real(i) + 10.*real(j)
      if (mod(i,2) == 0 .and. j /= 1) a(i,j) = -a(i,j)
    end do
  end do
  flag = .not. (x < y) .and. (iv(1) == 3 .or. .false.)
  if (flag .and. a(2,3) < 0.) then
    iv(7) = 1
  else if (a(2,3) > 0.) then
    iv(7) = 2
  else
    iv(7) = 3
  end if
  s = 0.
  call accumulate(a, s, 3, 2)
  call accumulate(a, s, 3, 2)
  iv(8) = int(s)
  write(*,*) 'p error:', s
end program main

subroutine accumulate(a, s, i0, j0)
  use consts
  implicit none
  real,intent(in),dimension(0:md,0:nd):: a
  real,intent(inout):: s
  integer,intent(in):: i0, j0
  integer:: calls
  real, parameter:: half = 0.5
  calls = calls + 1
  s = s + a(i0,j0)*half + a(i0+1,j0-1) + real(calls)
  return
end subroutine accumulate
"""


def test_translator_on_snippet():
    """every expected number below is computed by hand from the Fortran text"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "snip.f90")
        with open(src, "w") as f:
            f.write(SNIPPET)
        csrc = os.path.join(d, "snip.c")
        with open(csrc, "w") as f:
            f.write(f90toc.translate([(src, None, None)]))
        so = os.path.join(d, "snip.so")
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", build_ref.HERE, csrc,
                               os.path.join(build_ref.HERE, "ref_runtime.c"), "-o", so, "-lm"])
        import ctypes as C
        L = C.CDLL(so)
        L.ref_run.argtypes = [C.c_char_p]
        L.ref_lookup.argtypes = [C.c_char_p]
        L.ref_lookup.restype = C.POINTER(rt.RtVar)
        L.ref_perr.argtypes = [C.c_int]
        L.ref_perr.restype = C.c_double
        for run in range(2):       # the second run must start from zeroed statics again (calls = 0)
            assert L.ref_run(d.encode()) == 0
            r = np.ctypeslib.as_array(C.cast(L.ref_lookup(b"r").contents.ptr, C.POINTER(C.c_double)), (3,))
            iv = np.ctypeslib.as_array(C.cast(L.ref_lookup(b"iv").contents.ptr, C.POINTER(C.c_int)), (8,))
            a = np.ctypeslib.as_array(C.cast(L.ref_lookup(b"a").contents.ptr, C.POINTER(C.c_double)), (5, 7))
            assert r[0] == 7. / 2. * 4. - 3. * (2. * 2.)            # 2.0
            assert r[1] == -(7. * 7.) + (7. - 2.) / 2. / 2.          # -47.75
            assert r[2] == 1.e-1 + 2.5 * 3.0 + 1.0                   # 7/2 = 3 (integer), mod(7.,2.) = 1.
            assert list(iv[:4]) == [3, -3, -1, 32 + 9 - 3 + 4]
            assert iv[4] == 2 + 5 + 8 + 11 and iv[5] == 14           # Fortran: k = 14 after the loop
            # a[j][i] column-major with lower bounds 0; sign flipped where i even and j /= 1
            assert a[3, 2] == -(2. + 30.) and a[1, 2] == 12. and a[2, 3] == 23.
            assert iv[6] == 1                                        # flag true and a(2,3) = -32 < 0
            # accumulate: a(3,2)*0.5 + a(4,1) + calls, twice; a(3,2) = 23, a(4,1) = 14 (j = 1: not flipped)
            assert L.ref_perr(0) == (23. * 0.5 + 14. + 1.) + (23. * 0.5 + 14. + 2.)
            assert iv[7] == 54


def test_translator_expression_text():
    """the emitted C is fully parenthesised in Fortran's evaluation order"""
    tr = f90toc.Translator()
    tr.add_source("program main\nimplicit none\nreal:: a, b, c, d\ninteger:: i, m\n"
                  "a = b/c*d\na = -b*c + d\na = b - c - d\na = b**2*c\ni = (i - 1)/m + 1\nend program main\n", "t.f90")
    text = tr.emit()
    main = text[text.index("void f_MAIN(void)"):text.index("const rt_var rt_registry")]
    body = [ln.strip() for ln in main.splitlines() if ln.strip().startswith("f_") and "=" in ln]
    assert body == ["f_a = ((f_b/f_c)*f_d);", "f_a = ((-((f_b*f_c)))+f_d);", "f_a = ((f_b-f_c)-f_d);",
                    "f_a = (rt_sq(f_b)*f_c);", "f_i = ((((f_i-1))/f_m)+1);"]


def test_translator_rejects_what_it_does_not_know():
    tr = f90toc.Translator()
    tr.add_source("program main\nimplicit none\nreal:: a\na = undeclared + 1.\nend program main\n", "t.f90")
    with pytest.raises(SyntaxError):
        tr.emit()
    tr = f90toc.Translator()
    tr.add_source("program main\nimplicit none\nreal:: a\nwhere (a > 0.) a = 1.\nend program main\n", "t.f90")
    with pytest.raises(SyntaxError):
        tr.emit()


# ------------------------------------------------------------------------------------------------ helpers

def _oracle_for(oracle, case, dims, st, spacing, eps_in):
    m, n, l = (int(x) for x in dims)
    kw = dict(xnue=st["xnue"], xlambda=st["xlambda"], density=st["density"], thickness=st["thickness"],
              nonslip=st["nonslip"], iter_max=st["iter_max"], relux_factor=st["relux_factor"],
              inlet_velocity=st["inlet_velocity"], outlet_pressure=st["outlet_pressure"], AoA=st["AoA"])
    eps = np.maximum(eps_in, st["threshold"])      # lib/grid.f90:50 / :289
    if case.startswith("ibm3"):
        dx, dy, dz, dt = spacing
        P = oracle.make_params(m=m, n=n, l=l, dx=dx, dy=dy, dz=dz, dt=dt, **kw)
        oc = oracle.Oracle3D(P, case == "ibm3_air_condition", eps)
    else:
        dx, dy, dt = spacing
        P = oracle.make_params(m=m, n=n, dx=dx, dy=dy, dt=dt, **kw)
        oc = oracle.Oracle2D(P, case == "ibm2_backstep", eps)
    oc.initialise()
    return oc


def _restated_spacing(case, dims, st):
    """dx = width/real(m-1) ... dt = time/real(istep_max) (lib/grid.f90:54-56, :297-300), restated"""
    from pixelflow_b200 import workloads as wl
    m, n, l = (int(x) for x in dims)
    dx, dy, dz, dt = wl.grid_spacing(st["width"], st["height"], st["depth"], st["time"], st["istep_max"], m, n, l)
    return (dx, dy, dz, dt) if case.startswith("ibm3") else (dx, dy, dt)


def _fields(oc, case):
    names = ("u", "v", "w", "p") if case.startswith("ibm3") else ("u", "v", "p")
    d = {k: getattr(oc, k) for k in names}
    d["porosity"] = oc.e
    return d


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SMALL = ["u3_even", "u3_odd", "u3_mixed", "u3_pout", "a3_even", "a3_odd", "u2_even", "u2_odd", "b2_mixed", "d2_mixed"]


# ------------------------------------------------------------------------------------------------ 2. golden

@pytest.mark.parametrize("name", SMALL)
def test_oracle_equals_golden_reference_outputs(oracle, gold, name):
    case = str(gold[f"{name}/case"])
    dims = gold[f"{name}/dims"]
    st = json.loads(str(gold[f"{name}/settings"]))
    spacing = tuple(gold[f"{name}/spacing"])
    assert spacing == _restated_spacing(case, dims, st), "dx, dy, dz, dt as the reference computed them"
    oc = _oracle_for(oracle, case, dims, st, spacing, gold[f"{name}/porosity_in"])
    steps = st["istep_max"]
    errs, forces = [], []
    for _ in range(steps):
        errs.append(oc.step(1)[0])
        if case == "ibm2_drag":
            forces.append(oc.force_log(st["radius"]))
    for k, a in _fields(oc, case).items():
        assert np.array_equal(a, gold[f"{name}/{k}"]), f"{name}: {k} differs from the reference's"
    assert np.array_equal(np.array(errs), gold[f"{name}/perr"]), "logged p error per step"
    if case == "ibm2_drag":
        # serial sums in the reference's loop order: equal to the last bit
        assert np.array_equal(np.array(forces), gold[f"{name}/force"])


@pytest.mark.parametrize("deck,golden_deck", [("cylinder", "cylinder"), ("cylinder", "cylinder_drag"),
                                              ("backstep", "backstep"), ("room", "room"), ("room", "room_long")])
def test_oracle_equals_golden_shipped_decks(oracle, gold, deck, golden_deck):
    """the reference's three shipped decks, run unmodified by the translated programs for 3 steps; the oracle is fed
    from the independent fixture tests/golden/decks/*.npz (make_decks.py) and must reproduce the field hashes"""
    from pixelflow_b200.controldict import parse_controldict
    z = np.load(os.path.join(HERE, "golden", "decks", deck + ".npz"))
    cd = parse_controldict(str(z["controldict"]))
    case = str(gold[f"deck_{golden_deck}/case"])
    steps = int(gold[f"deck_{golden_deck}/steps"])
    m, n, l = (int(x) for x in z["dims"])
    st = dict(xnue=cd.xnue, xlambda=cd.xlambda, density=cd.density, thickness=cd.thickness, nonslip=cd.nonslip,
              iter_max=cd.iter_max, relux_factor=cd.relux_factor, inlet_velocity=cd.inlet_velocity,
              outlet_pressure=cd.outlet_pressure, AoA=cd.AoA, threshold=cd.threshold, width=cd.width,
              height=cd.height, depth=cd.depth, time=cd.time, istep_max=cd.istep_max, radius=cd.radius)
    spacing = tuple(gold[f"deck_{golden_deck}/spacing"])
    assert spacing == _restated_spacing(case, (m, n, l), st)
    eps = z["porosity"] if case.startswith("ibm3") else z["porosity"][0]
    oc = _oracle_for(oracle, case, (m, n, l), st, spacing, eps)
    errs, forces = [], []
    for _ in range(steps):
        errs.append(oc.step(1)[0])
        if case == "ibm2_drag":
            forces.append(oc.force_log(cd.radius))
    sha = json.loads(str(gold[f"deck_{golden_deck}/sha"]))
    assert np.array_equal(oc.p.ravel()[::37], gold[f"deck_{golden_deck}/p_sample"])
    for k, a in _fields(oc, case).items():
        assert _sha(a) == sha[k], f"{golden_deck}: {k} differs from the reference's"
    assert np.array_equal(np.array(errs), gold[f"deck_{golden_deck}/perr"])
    if case == "ibm2_drag":
        assert np.array_equal(np.array(forces), gold[f"deck_{golden_deck}/force"])


@pytest.mark.parametrize("name", SMALL)
def test_host_porosity_halos_equal_reference(gold, name):
    """pixelflow_b200.workloads.with_halos (what the Python host side feeds pf_set_porosity) == the padded porosity
    array the reference's grid routine left behind, corners included"""
    from pixelflow_b200 import workloads as wl
    case = str(gold[f"{name}/case"])
    st = json.loads(str(gold[f"{name}/settings"]))
    e = wl.with_halos(np.maximum(gold[f"{name}/porosity_in"], st["threshold"]), case)
    assert np.array_equal(e, gold[f"{name}/porosity"])


# ------------------------------------------------------------------------------------------------ 3. live

def _force_lines(log):
    num = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"
    vals = []
    for line in log.splitlines():
        t = line.strip()
        if t.startswith(("Fp =", "Fv =", "F  =", "Cd =")):
            vals += [float(x) for x in re.findall(num, t.split("=", 1)[1].replace("Cl =", " "))]
    return np.array(vals).reshape(-1, 8)


LIVE = [
    # case, dims, extra settings
    ("ibm3_uniform", (16, 12, 10), dict()),
    ("ibm3_uniform", (15, 11, 9), dict(AoA=20.0, xlambda=1e-4)),
    ("ibm3_uniform", (16, 11, 10), dict(nonslip=False)),
    ("ibm3_uniform", (15, 12, 9), dict(outlet_pressure=0.2)),
    ("ibm3_air_condition", (14, 12, 10), dict()),
    ("ibm3_air_condition", (13, 11, 9), dict(outlet_pressure=0.2)),
    ("ibm2_uniform", (40, 24, 1), dict()),
    ("ibm2_uniform", (39, 23, 1), dict(AoA=8.0)),
    ("ibm2_backstep", (40, 23, 1), dict()),
    ("ibm2_drag", (39, 24, 1), dict(radius=0.05)),
]


def _live_deck(case, dims, extra, seed, steps=4, iter_max=12):
    m, n, l = dims
    rng = np.random.default_rng(seed)
    kk, jj, ii = np.meshgrid(np.arange(l), np.arange(n), np.arange(m), indexing="ij")
    r = np.sqrt(((ii - 0.35 * m) / (0.22 * m)) ** 2 + ((jj - 0.5 * n) / (0.3 * n)) ** 2
                + (((kk - 0.5 * l) / (0.3 * l)) ** 2 if l > 1 else 0))
    eps = np.clip(0.5 * np.tanh((r - 1.0) * 2.0) + 0.5 + 0.06 * (rng.random((l, n, m)) - 0.5), 0.02, 1.0)
    st = dict(rt.DEFAULTS)
    st.update(width=0.1 * (m - 1) / 16, height=0.1 * (n - 1) / 16, depth=0.1 * max(l - 1, 1) / 16, time=0.0005 * steps,
              istep_max=steps, iter_max=iter_max, inlet_velocity=0.9)
    st.update(extra)
    return (eps if l > 1 else eps[0]), st


@needs_ref
@pytest.mark.parametrize("flavour", ["serial", "omp"])
@pytest.mark.parametrize("case,dims,extra", LIVE)
def test_oracle_equals_translated_reference_live(oracle, case, dims, extra, flavour, monkeypatch):
    """`program main` of the translated reference, run on a freshly written project directory (controlDict namelists
    + CSV read by the translated read_settings / grid routine), against the oracle.  The OpenMP flavour (directives
    translated to #pragma omp, 3 threads) must give the same bits: the path has max reductions only."""
    if case == "ibm2_drag" and flavour == "omp":
        pytest.skip("output_force_log_2d's OpenMP sum is order-dependent (and racy in the reference, SURVEY 5)")
    monkeypatch.setenv("OMP_NUM_THREADS", "3")
    eps, st = _live_deck(case, dims, extra, seed=zlib.crc32(repr((case, dims)).encode()))
    R = rt.RefProgram(case, flavour, "s")
    with tempfile.TemporaryDirectory() as d:
        rt.write_deck(d, eps, **st)
        perr = R.run(d)
    ref = R.fields()
    assert len(perr) == st["istep_max"]
    spacing = tuple(R.scalar(k) for k in (("dx", "dy", "dz", "dt") if R.d3 else ("dx", "dy", "dt")))
    assert spacing == _restated_spacing(case, dims, st)
    oc = _oracle_for(oracle, case, dims, st, spacing, eps)
    errs, forces = [], []
    for _ in range(st["istep_max"]):
        errs.append(oc.step(1)[0])
        if case == "ibm2_drag":
            forces.append(oc.force_log(st["radius"]))
    for k, a in _fields(oc, case).items():
        assert np.array_equal(a, ref[k]), f"{case} {dims} {flavour}: {k}"
    assert np.array_equal(np.array(errs), perr)
    if case == "ibm2_drag":
        assert np.array_equal(np.array(forces), _force_lines(R.log()))
    # the stubs are the output routines only, called as often as the program text says
    assert R.stub_count("get_now_time") == 4
    out0 = "output_paraview_temp_3d" if R.d3 else "output_paraview_temp_2d"
    assert R.stub_count(out0) >= 1


@needs_ref
def test_translated_reference_reproduces_golden(gold):
    """the committed golden file is what the translated reference produces today (guards against a stale fixture)"""
    for name in ("u3_odd", "a3_odd", "b2_mixed"):
        case = str(gold[f"{name}/case"])
        st = json.loads(str(gold[f"{name}/settings"]))
        R = rt.RefProgram(case, "serial", "s")
        with tempfile.TemporaryDirectory() as d:
            rt.write_deck(d, gold[f"{name}/porosity_in"], **st)
            perr = R.run(d)
        assert np.array_equal(perr, gold[f"{name}/perr"])
        for k, a in R.fields().items():
            assert np.array_equal(a, gold[f"{name}/{k}"]), (name, k)


@needs_ref
def test_step_limit_equals_short_run(gold):
    """leaving the time loop after n steps (how the shipped decks are run) == running a deck whose time loop has the
    same dt and n steps"""
    name = "u2_even"
    st = json.loads(str(gold[f"{name}/settings"]))
    eps = gold[f"{name}/porosity_in"]
    R = rt.RefProgram("ibm2_uniform", "serial", "s")
    long = dict(st, istep_max=st["istep_max"] * 4, time=st["time"] * 4)
    with tempfile.TemporaryDirectory() as d:
        rt.write_deck(d, eps, **long)
        perr = R.run(d, step_limit=st["istep_max"])
    if R.scalar("dt") != gold[f"{name}/spacing"][-1]:
        pytest.skip("4*time/(4*steps) rounds differently from time/steps")
    assert np.array_equal(perr, gold[f"{name}/perr"])
    assert np.array_equal(R.array("p"), gold[f"{name}/p"])


@needs_ref
@pytest.mark.parametrize("name", ["u3_even", "a3_odd", "u2_even", "b2_mixed"])
def test_reference_as_shipped_fp32_is_close_to_the_fp64_evaluation(gold, name):
    """The reference's build files set no real kind: as shipped it computes in 32-bit reals, while the
    north star fixes fp64 (= `-fdefault-real-8`, what every parity test here compares with).  The "r4" flavour is the
    same translation with `real` = float and libm's float functions; after the golden cases' 3 steps it sits at
    float rounding distance from the fp64 fields — the size of the change a user of the shipped build sees."""
    case = str(gold[f"{name}/case"])
    st = json.loads(str(gold[f"{name}/settings"]))
    R = rt.RefProgram(case, "r4", "s")
    with tempfile.TemporaryDirectory() as d:
        rt.write_deck(d, gold[f"{name}/porosity_in"], **st)
        perr = R.run(d)
    f = R.fields()
    for k in ("u", "v", "p"):
        assert f[k].dtype == np.float32
        ref = gold[f"{name}/{k}"]
        rel = np.linalg.norm((f[k].astype(np.float64) - ref).ravel()) / np.linalg.norm(ref.ravel())
        assert 0 < rel < 2e-5, (k, rel)
    assert np.allclose(perr, gold[f"{name}/perr"], rtol=1e-4)


# module wall_conditions (ibm_3d_air_condition_omp_cpu.f90:4-16) holds compile-time parameters; the ABI takes them at
# run time (pf_config.wall).  Order here and there: top, bottom, east, west, south, north.
WALLS = [(0, 0, 0, 0, 0, 0), (2, 1, 2, 1, 1, 2), (1, 2, 1, 2, 2, 1), (0, 2, 2, 0, 1, 1), (2, 2, 2, 2, 2, 2),
         (1, 1, 1, 1, 1, 1), (0, 1, 2, 0, 1, 2), (2, 0, 1, 2, 0, 1)]


@needs_ref
@pytest.mark.parametrize("wall", WALLS)
@pytest.mark.parametrize("dims", [(10, 8, 6), (9, 7, 8)])
def test_air_condition_wall_codes_equal_translated_reference(oracle, wall, dims):
    """every face as wall / inlet / outlet: the reference rebuilt (translated) with those wall_conditions parameters
    against the oracle given the same codes at run time — boundary_matrix (:665-867) and boundary (:873-1170) are
    300 lines of per-face if ladders, and their order of application matters at the edges"""
    if not build_ref.available():
        pytest.skip("wall-code variants are translated on demand from /root/reference")
    names = ("top_wall", "bottom_wall", "east_wall", "west_wall", "south_wall", "north_wall")
    lib = build_ref.build_variant("ibm_3d_air_condition_omp_cpu", "w" + "".join(str(w) for w in wall),
                                  dict(zip(names, wall)))
    m, n, l = dims
    rng = np.random.default_rng(sum(wall) * 10 + m)
    # porosity: both >= 0.9 and < 0.9 on every face (the inlet / outlet patches are where porosity >= 0.9)
    eps = np.clip(0.55 + 0.5 * rng.random((l, n, m)), 0.05, 1.0)
    st = dict(rt.DEFAULTS)
    st.update(width=0.1 * (m - 1) / 16, height=0.1 * (n - 1) / 16, depth=0.1 * (l - 1) / 16, time=0.0005 * 3,
              istep_max=3, iter_max=8, inlet_velocity=0.7, outlet_pressure=0.05)
    R = rt.RefProgram("ibm3_air_condition", lib=lib)
    with tempfile.TemporaryDirectory() as d:
        rt.write_deck(d, eps, **st)
        perr = R.run(d)
    ref = R.fields()
    spacing = tuple(R.scalar(k) for k in ("dx", "dy", "dz", "dt"))
    kw = dict(xnue=st["xnue"], xlambda=st["xlambda"], density=st["density"], thickness=st["thickness"],
              nonslip=st["nonslip"], iter_max=st["iter_max"], relux_factor=st["relux_factor"],
              inlet_velocity=st["inlet_velocity"], outlet_pressure=st["outlet_pressure"], AoA=st["AoA"])
    P = oracle.make_params(m=m, n=n, l=l, dx=spacing[0], dy=spacing[1], dz=spacing[2], dt=spacing[3], wall=wall, **kw)
    oc = oracle.Oracle3D(P, True, np.maximum(eps, st["threshold"]))
    oc.initialise()
    errs = oc.step(3)
    for k, a in (("u", oc.u), ("v", oc.v), ("w", oc.w), ("p", oc.p), ("porosity", oc.e)):
        assert np.array_equal(a, ref[k]), f"wall={wall} {dims}: {k}"
    assert np.array_equal(errs, perr)


@needs_ref
def test_force_log_3d_equals_translated_reference(oracle, gold):
    """output_force_log_3d (lib/output.f90:1090-1165) is called by none of the programs; here it is called on the
    fields a translated ibm3 run left behind, and the oracle's restatement must print the same 12 numbers"""
    name = "u3_odd"
    st = json.loads(str(gold[f"{name}/settings"]))
    dims = gold[f"{name}/dims"]
    R = rt.RefProgram("ibm3_uniform", "serial", "s")
    with tempfile.TemporaryDirectory() as d:
        rt.write_deck(d, gold[f"{name}/porosity_in"], **st)
        R.run(d)
    radius = 0.0123
    n0 = len(R.log().splitlines())
    R.call("output_force_log_3d", "p", "u", "v", "w", "dx", "dy", "dz", "porosity", "m", "n", "l",
           "xnue", "density", "thickness", radius, "inlet_velocity")
    lines = R.log().splitlines()[n0:]
    assert [ln.split("=")[0].strip() for ln in lines] == ["Fp", "Fv", "F", "Cd(x)"]
    num = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"
    vals = []
    for ln in lines:
        body = ln.split("=", 1)[1].replace("Cl =", " ").replace("Cd(z) =", " ")
        vals += [float(x) for x in re.findall(num, body)]
    assert len(vals) == 12
    oc = _oracle_for(oracle, "ibm3_uniform", dims, st, tuple(gold[f"{name}/spacing"]), gold[f"{name}/porosity_in"])
    oc.step(st["istep_max"])
    assert np.array_equal(oc.force_log(radius), np.array(vals))


@needs_ref
def test_translated_reference_does_not_depend_on_the_optimiser(gold):
    """-O0 and -O3 builds of the translated program leave the same bits (the C is fully parenthesised, no fast-math,
    no contraction): the golden vectors are a property of the source, not of gcc's optimiser"""
    if not build_ref.available():
        pytest.skip("variants are translated on demand from /root/reference")
    name = "u3_mixed"
    st = json.loads(str(gold[f"{name}/settings"]))
    base = ["-ffp-contract=off", "-fPIC", "-shared", "-mcmodel=medium", "-fno-strict-aliasing"]
    for tag, flags in (("O0", ["-O0"] + base), ("O2native", ["-O2", "-march=native"] + base)):
        lib = build_ref.build_variant("ibm_3d_uniform_omp_cpu", tag, {}, cflags=flags)
        R = rt.RefProgram("ibm3_uniform", lib=lib)
        with tempfile.TemporaryDirectory() as d:
            rt.write_deck(d, gold[f"{name}/porosity_in"], **st)
            perr = R.run(d)
        assert np.array_equal(perr, gold[f"{name}/perr"]), tag
        for k, a in R.fields().items():
            assert np.array_equal(a, gold[f"{name}/{k}"]), (tag, k)


ISO_SNIPPET = """
program main
  use iso_c_binding
  use fake_binding
  implicit none
  real, dimension(0:3, 0:2) :: a
  real :: x, r(2)
  integer :: n, k
  logical :: flag
  type(box_t) :: b
  type(c_ptr) :: h
  a = 2.5; n = 3; flag = .true.
  b%count = n; b%scale = 0.5; b%codes(K_SECOND) = 7
  b%count = merge(b%count + 1, 0, flag)
  x = 0.
  if (box_open(h, b) /= 0) then
    stop 1
  end if
  call box_fill(h, a, n, x)
  r(1) = x
  r(2) = real(box_count(h)) + a(3, 2)
  k = box_close(h)
  write(*,*) 'p error:', r(1)
  write(*,*) 'p error:', r(2)
end program main
"""

ISO_C = """
#include <stdlib.h>
typedef struct { int count; double scale; int codes[3]; } box_t;
static box_t *the_box;
int box_open(void **h, box_t *b) { the_box = malloc(sizeof *the_box); *the_box = *b; *h = the_box; return b->codes[1] == 7 ? 0 : 1; }
void box_fill(void *h, double *a, int n, double *x) { box_t *b = h; *x = a[5] * b->scale * n + b->count; }
int box_count(void *h) { return ((box_t *)h)->count; }
int box_close(void *h) { free(h); return 0; }
"""


def test_translator_iso_c_binding_subset(tmp_path):
    """derived-type components, `;`, merge, whole-array assignment, bind(C) procedures by value and by reference —
    what the product's Fortran drivers use on top of the reference's subset"""
    desc = {"name": "fake_binding", "types": {"box_t": [("count", "int", 0), ("scale", "real", 0), ("codes", "int", 3)]},
            "params": {"k_second": 2},
            "functions": {"box_open": {"ret": "int", "args": [("cptr", "ref"), ("type:box_t", "ref")]},
                          "box_fill": {"ret": "void", "args": [("cptr", "value"), ("real", "array"), ("int", "value"),
                                                               ("real", "ref")]},
                          "box_count": {"ret": "int", "args": [("cptr", "value")]},
                          "box_close": {"ret": "int", "args": [("cptr", "value")]}}}
    tr = f90toc.Translator()
    tr.add_c_module(desc)
    tr.add_source(ISO_SNIPPET, "iso.f90")
    (tmp_path / "iso.c").write_text(tr.emit())
    (tmp_path / "fake.c").write_text(ISO_C)
    so = tmp_path / "iso.so"
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-I", build_ref.HERE, str(tmp_path / "iso.c"),
                           str(tmp_path / "fake.c"), os.path.join(build_ref.HERE, "ref_runtime.c"), "-o", str(so),
                           "-lm", "-ldl"])
    import ctypes as C
    L = C.CDLL(str(so))
    L.ref_run.argtypes = [C.c_char_p]
    L.ref_perr.argtypes = [C.c_int]
    L.ref_perr.restype = C.c_double
    assert L.ref_run(str(tmp_path).encode()) == 0
    # count = merge(3 + 1, 0, .true.) = 4; x = a(1,1) * 0.5 * 3 + 4 = 2.5 * 1.5 + 4; r(2) = 4 + 2.5
    assert L.ref_perr(0) == 2.5 * 0.5 * 3 + 4 and L.ref_perr(1) == 6.5


@needs_ref
@pytest.mark.parametrize("threads", [1, 2, 5, 8])
def test_openmp_flavour_is_thread_count_independent(gold, threads):
    """the translated `!$omp` program gives the golden bits for any team size: the path's only reduction is a max"""
    for name in ("u3_odd", "a3_even", "u2_odd"):
        case = str(gold[f"{name}/case"])
        st = json.loads(str(gold[f"{name}/settings"]))
        R = rt.RefProgram(case, "omp", "s")
        assert R.set_threads(threads) == threads
        with tempfile.TemporaryDirectory() as d:
            rt.write_deck(d, gold[f"{name}/porosity_in"], **st)
            perr = R.run(d)
        assert np.array_equal(perr, gold[f"{name}/perr"]), (name, threads)
        for k, a in R.fields().items():
            assert np.array_equal(a, gold[f"{name}/{k}"]), (name, threads, k)
