"""The drop-in driver's log and output files against the REFERENCE's own, byte for byte — on CPU.

Reference side: the translated programs in the "gf" flavour (oracle/build_ref.py): lib/output.f90 is translated
as well and every WRITE is executed by libgfortran.so.5, the runtime library of a gfortran build.  A run directory
then holds what the reference leaves behind: stdout.log (unit *), etc/grid.dat, etc/solution_uvp.dat,
etc/divergent.dat, etc/surface_profile.dat, <output_folder>/output_NNNNN.vtk and output_paraview.vtk.

Driver side: pixelflow_b200/driver/pixelflow_driver `--replay RECORD` — the driver's own log and file writers fed with
a recorded run (p errors, force log, final fields) instead of the device; no computation happens in that mode.  The
VTK bodies are formatted on the GPU in a real run (pf_vtk_section); here the driver writes the header lines only and
the bodies are checked through oracle_np.vtk_section, the checker tests/test_gpu_output.py holds pf_vtk_section to.

Only the `# --- TIME:` stamps are left out (wall-clock; get_now_time is a stub in the translated programs).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import build_ref, gfortran_rt
from oracle import oracle_np as onp
from oracle import ref_translated as rt

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DRIVER = os.path.join(ROOT, "pixelflow_b200", "driver", "pixelflow_driver")

pytestmark = [
    pytest.mark.skipif(not (build_ref.available() or os.path.isdir(build_ref.OUT)),
                       reason="oracle/_ref not built and /root/reference absent"),
    pytest.mark.skipif(gfortran_rt.find_libgfortran() is None, reason="libgfortran.so.5 not found"),
]

CASES = [
    # case, dims, settings
    ("ibm3_uniform", (12, 10, 8), dict(istep_out=2, AoA=4.0)),
    ("ibm3_air_condition", (9, 8, 7), dict(istep_out=1)),
    ("ibm2_uniform", (20, 12, 1), dict(istep_out=2)),
    ("ibm2_backstep", (19, 11, 1), dict(istep_out=3)),
    ("ibm2_drag", (20, 11, 1), dict(istep_out=100, radius=0.05)),
]
STEPS = 3


def _deck(case, dims, extra):
    m, n, l = dims
    rng = np.random.default_rng(m * 100 + n)
    kk, jj, ii = np.meshgrid(np.arange(l), np.arange(n), np.arange(m), indexing="ij")
    r = np.sqrt(((ii - 0.35 * m) / (0.22 * m)) ** 2 + ((jj - 0.5 * n) / (0.3 * n)) ** 2
                + (((kk - 0.5 * l) / (0.3 * l)) ** 2 if l > 1 else 0))
    eps = np.clip(0.5 * np.tanh((r - 1.0) * 2.0) + 0.5 + 0.06 * (rng.random((l, n, m)) - 0.5), 0.0, 1.0)
    eps[r < 0.3] = 0.0       # exact zeros: clamped to the threshold on read, and p_fluid's `porosity > small` switch
    st = dict(rt.DEFAULTS)
    st.update(width=0.1 * (m - 1) / 16, height=0.1 * (n - 1) / 16, depth=0.1 * max(l - 1, 1) / 16,
              time=0.0002 * STEPS, istep_max=STEPS, iter_max=6, inlet_velocity=0.9, output_folder="out",
              csv_file="data/poro.csv")
    st.update(extra)
    return (eps if l > 1 else eps[0]), st


def _write_pfbin(csv, eps_clamped, d3, threshold):
    """the driver's binary cache of a parsed CSV (pixelflow_driver.cpp: CacheHeader): lets it start without the GPU
    CSV parser"""
    sc = os.stat(csv)
    if d3:
        l, n, m = eps_clamped.shape
        a = np.zeros((l + 2, n + 2, m + 2))
        a[1:-1, 1:-1, 1:-1] = eps_clamped
    else:
        n, m = eps_clamped.shape
        l = 1
        a = np.zeros((n + 2, m + 2))
        a[1:-1, 1:-1] = eps_clamped
    with open(csv + ".pfbin", "wb") as f:
        f.write(struct.pack("<8siiiidqqq", b"PFBIN02\0", m, n, l, int(d3), float(threshold), sc.st_size,
                            sc.st_mtime_ns // 10 ** 9, sc.st_mtime_ns % 10 ** 9))
        f.write(a.tobytes())


def _write_replay(path, perr, force, fields, d3):
    with open(path, "wb") as f:
        f.write(struct.pack("<8sii", b"PFREPLAY", len(perr), 1 if force is not None else 0))
        f.write(np.asarray(perr, dtype=np.float64).tobytes())
        if force is not None:
            f.write(np.asarray(force, dtype=np.float64).tobytes())
        for k in (("u", "v", "w", "p") if d3 else ("u", "v", "p")):
            f.write(np.ascontiguousarray(fields[k]).tobytes())


def _no_time(text):
    return [ln for ln in text.splitlines() if not ln.startswith(" # --- TIME:")]


def _is_record(line):
    """a body record of a VTK file: one or three f16.4 items"""
    toks = line.split()
    if len(toks) not in (1, 3):
        return False
    for t in toks:
        if set(t) == {"*"} or t in ("NaN", "Infinity", "-Infinity"):
            continue
        try:
            float(t)
        except ValueError:
            return False
        if "." not in t:
            return False
    return True


@pytest.fixture(scope="module", autouse=True)
def _built():
    if build_ref.available():
        build_ref.build()
    from pixelflow_b200 import build
    build.build_library()
    build.build_drivers()


@pytest.mark.parametrize("case,dims,extra", CASES)
def test_driver_log_and_files_equal_the_reference(case, dims, extra, tmp_path):
    d3 = case.startswith("ibm3")
    eps, st = _deck(case, dims, extra)
    # ---------------- the reference, run as a user runs it
    ref_dir = tmp_path / "ref"
    (ref_dir / "etc").mkdir(parents=True)          # the program's `call system('mkdir -p ...')` is a stub
    (ref_dir / st["output_folder"]).mkdir()
    rt.write_deck(str(ref_dir), eps, **st)
    R = rt.RefProgram(case, "gf", "s")
    perr = R.run(str(ref_dir))
    fields = R.fields()
    assert len(perr) == STEPS
    ref_log = (ref_dir / "stdout.log").read_text()
    force = None
    if case == "ibm2_drag":
        from tests.test_ref_translation import _force_lines
        force = _force_lines(R.log())
        assert force.shape == (STEPS, 8)
    # ---------------- the driver, replaying that run
    drv_dir = tmp_path / "drv"
    drv_dir.mkdir()
    rt.write_deck(str(drv_dir), eps, **st)
    _write_pfbin(str(drv_dir / st["csv_file"]), np.maximum(eps, st["threshold"]), d3, st["threshold"])
    _write_replay(str(drv_dir / "run.replay"), perr, force, fields, d3)
    r = subprocess.run([DRIVER, "--case", case, "--replay", "run.replay", "--cache"], cwd=drv_dir, capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    # ---------------- the log
    assert _no_time(r.stdout) == _no_time(ref_log)
    # ---------------- list-directed files: byte for byte
    for name in ("etc/grid.dat", "etc/solution_uvp.dat", "etc/divergent.dat", "etc/surface_profile.dat"):
        a, b = (drv_dir / name).read_bytes(), (ref_dir / name).read_bytes()
        if a != b:
            la, lb = a.decode().splitlines(), b.decode().splitlines()
            first = next((i for i, (x, y) in enumerate(zip(la, lb)) if x != y), min(len(la), len(lb)))
            raise AssertionError(f"{name}: line {first + 1} of {len(la)}/{len(lb)}:\n driver    {la[first][:200] if first < len(la) else '<eof>'}\n"
                                 f" reference {lb[first][:200] if first < len(lb) else '<eof>'}")
    # ---------------- VTK files: the same set of files; header lines from the driver, bodies through the checker
    ref_vtk = sorted(os.listdir(ref_dir / st["output_folder"]))
    assert sorted(os.listdir(drv_dir / st["output_folder"])) == ref_vtk
    snaps = [0] + [i for i in range(1, STEPS + 1) if i % st["istep_out"] == 0]
    assert ref_vtk == sorted([f"output_{i:05d}.vtk" for i in snaps] + ["output_paraview.vtk"])
    for name in ref_vtk:
        ref_lines = (ref_dir / st["output_folder"] / name).read_text().splitlines()
        drv_lines = (drv_dir / st["output_folder"] / name).read_text().splitlines()
        assert drv_lines == [ln for ln in ref_lines if not _is_record(ln)], name
    # bodies of the final file == the restated section formatter on the final fields
    xp, yp = R.array("xp")[:dims[0] + 2], R.array("yp")[:dims[1] + 2]
    zp = R.array("zp")[:dims[2] + 2] if d3 else None
    ref_final = (ref_dir / st["output_folder"] / "output_paraview.vtk").read_bytes()
    u, v, p, e = fields["u"], fields["v"], fields["p"], fields["porosity"]
    w = fields.get("w")
    pos = 0
    order = (["points", "velocity", "velocityInFluid", "pressure", "VelocityDivergent", "porosity"] if d3 else
             ["points", "velocity", "velocityInFluid", "dimless_v", "porosity", "pressure", "VelocityDivergent",
              "abs_dimless_v"])
    for section in order:
        body = onp.vtk_section(section, 3 if d3 else 2, u, v, w, p, e, xp, yp, zp, inlet_velocity=st["inlet_velocity"])
        at = ref_final.find(body, pos)
        assert at >= 0, f"section {section}: the restated body is not in the reference's file"
        gap = ref_final[pos:at].decode()
        assert all(not _is_record(ln) for ln in gap.splitlines()), f"records between sections before {section}"
        pos = at + len(body)
    assert pos == len(ref_final)


def test_gf_flavour_reads_its_inputs_with_libgfortran(tmp_path):
    """in the "gf" flavour OPEN / READ / namelist READ are libgfortran's as well: a list-directed record in the repeat
    form `r*c` and a namelist with Fortran's upper-case / `d`-exponent spellings — which the plain runtime's reader
    does not know — are read like any other"""
    eps, st = _deck("ibm2_uniform", (20, 12, 1), dict(istep_out=100))
    for flavour in ("gf", "serial"):
        d = tmp_path / flavour
        (d / "etc").mkdir(parents=True)
        (d / st["output_folder"]).mkdir()
        rt.write_deck(str(d), eps, **st)
        csv = d / st["csv_file"]
        lines = csv.read_text().splitlines()
        assert lines[1].startswith("1, 1, 1, ")
        lines[1] = "3*1, " + lines[1].split(", ", 3)[3]          # x = y = z = 1
        csv.write_text("\n".join(lines) + "\n")
        cd = d / "config" / "controlDict.txt"
        cd.write_text(cd.read_text().replace("xnue = 0.001", "XNUE = 1.0D-3"))
        R = rt.RefProgram("ibm2_uniform", flavour, "s")
        if flavour == "gf":
            perr_gf = R.run(str(d))
            p_gf = R.array("p")
        else:
            with pytest.raises(RuntimeError, match="bad integer"):
                R.run(str(d))
    # and the result is the one of the plainly written deck
    d = tmp_path / "plain"
    (d / "etc").mkdir(parents=True)
    (d / st["output_folder"]).mkdir()
    rt.write_deck(str(d), eps, **st)
    R = rt.RefProgram("ibm2_uniform", "serial", "s")
    perr = R.run(str(d))
    assert np.array_equal(perr, perr_gf) and np.array_equal(R.array("p"), p_gf)
