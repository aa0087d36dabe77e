"""The I/O on either side of the hot path, pinned by the reference's own Fortran RUNTIME.

libgfortran.so.5 -- what executes the reference's formatted and list-directed statements -- is in the image although no
Fortran compiler is (oracle/gfortran_rt.py).  tests/golden/gfortran_io.npz holds its answers
(tests/golden/make_gfortran_io.py); here:
  * the live runtime still gives those answers (when the library is present),
  * the restated f16.4 editing of oracle/oracle_np.py -- the checker of the GPU snapshot formatter -- equals them on the
    very values tests/test_gpu_output.py::test_f16_4_torture feeds the GPU,
  * the C++ driver's list-directed imitation (`--format-selftest`) reproduces the runtime's records byte for byte,
  * the expectations of tests/test_gpu_ingest.py (Python's float() on the value text) equal what
    `read(52,*) x, y, z, poro_val` returns for every record form the reference's tools write.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import gfortran_rt as gf
from oracle import oracle_np as onp
from tests.gfortran_cases import F16_VALUES, LIST_RECORDS, READ_RECORDS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(HERE, "golden", "gfortran_io.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return {"values": z["f16_values"], "records": bytes(z["f16_records"]).decode().split("\n"),
            "three": bytes(z["f16_three"]).decode().split("\n"), **meta}


def test_fixture_inputs_are_the_shared_cases(golden):
    vals = np.array(F16_VALUES())
    assert np.array_equal(np.isnan(vals), np.isnan(golden["values"]))
    ok = ~np.isnan(vals)
    assert np.array_equal(vals[ok], golden["values"][ok]) and len(golden["records"]) == len(vals)
    assert len(golden["list_write"]) == len(LIST_RECORDS()) and [r["line"] for r in golden["list_read"]] == READ_RECORDS()


@pytest.mark.skipif(not gf.available(), reason="no libgfortran.so.5 in this environment")
def test_live_runtime_reproduces_the_fixture(golden):
    vals = F16_VALUES()
    assert gf.formatted_write("(3(f16.4,1x))", vals[:3000], 1).decode().split("\n")[:-1] == golden["records"][:3000]
    assert gf.formatted_write("(3(f16.4,1x))", vals[:3000], 3).decode().split("\n")[:-1] == golden["three"][:1000]
    assert [gf.list_write(*r).decode() for r in LIST_RECORDS()] == golden["list_write"]
    for rec in golden["list_read"]:
        ios, x, y, z, v = gf.list_read_record(rec["line"])
        assert ios == rec["iostat"], rec
        if ios == 0:
            assert [x, y, z] == rec["xyz"] and float(v).hex() == rec["value"], rec


def test_restated_f16_4_equals_libgfortran(golden):
    """every torture value: ties, signed zero, tiny, huge, asterisks, NaN / Infinity"""
    vals = golden["values"]
    mine = [onp.f16_4(float(v)) for v in vals]
    bad = [(float(v).hex(), a, b) for v, a, b in zip(vals, mine, golden["records"]) if a != b]
    assert not bad, bad[:10]
    assert golden["records"][1] == "         -0.0000" and golden["records"][2] == "          0.0312"
    assert golden["records"][19] == "*" * 16 and golden["records"][26] == "             NaN"


def test_restated_records_equal_libgfortran(golden):
    """(3(f16.4,1x)): three items -> 50 columns, one item -> 16; the trailing 1x leaves no blank (as pf_output.cu writes)"""
    vals = golden["values"]
    n3 = len(vals) // 3 * 3
    cols = [np.ascontiguousarray(vals[q:n3:3]) for q in range(3)]
    assert onp._records(cols).decode().split("\n")[:-1] == golden["three"]
    assert onp._records([vals]).decode().split("\n")[:-1] == golden["records"]
    assert all(len(r) == 50 for r in golden["three"]) and all(len(r) == 16 for r in golden["records"])


def test_driver_list_directed_output_equals_libgfortran(golden):
    """log lines and etc/*.dat rows of the C++ twin driver: leading blank, I11 / 25-column items, separators"""
    exe = os.path.join(ROOT, "pixelflow_b200", "driver", "pixelflow_driver")
    if not os.path.exists(exe):
        from pixelflow_b200 import build
        build.build_library()
        build.build_drivers()
    r = subprocess.run([exe, "--format-selftest"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    got = r.stdout.split("\n")[:-1]
    want = [s.rstrip("\n") for s in golden["list_write"]]
    assert len(got) == len(want)
    bad = [(a, b) for a, b in zip(got, want) if a != b]
    assert not bad, bad[:5]


def test_csv_record_semantics_equal_list_directed_read(golden):
    """tests/test_gpu_ingest.py expects max(float(text), threshold) per record: the same numbers the runtime reads.
    Records the runtime rejects are the ones the GPU parser must reject too -- except index columns written as reals
    (`1.0,1,1,0.5`), which list-directed input refuses for an integer item and pf_parse_porosity_csv accepts."""
    accepted = rejected = 0
    for rec in golden["list_read"]:
        line = rec["line"]
        if rec["iostat"] == 0:
            parts = line.replace(",", " ").split()
            assert [int(float(t)) for t in parts[:3]] == rec["xyz"], rec
            assert float(parts[3].lower().replace("d", "e")).hex() == rec["value"], rec
            accepted += 1
        else:
            rejected += 1
    assert accepted >= 95 and rejected == 3
    by_line = {r["line"]: r for r in golden["list_read"]}
    assert by_line["1,1,1,abc"]["iostat"] != 0 and by_line["1.0,1,1,0.5"]["iostat"] != 0
    # a record with too few items reads on into the next record in Fortran (here: end of the internal unit)
    assert by_line["1,1,1"]["iostat"] != 0
    # trailing extra items are ignored by list-directed input -- and by the GPU parser (test_gpu_ingest.py)
    assert by_line["1,1,1,0.5,7"]["iostat"] == 0


# ---- namelist input (config/controlDict.txt): the runtime's own reader against the two parsers of this repository --------
def _decks():
    out = {}
    for name in ("cylinder", "backstep", "room"):
        out[name] = str(np.load(os.path.join(HERE, "golden", "decks", name + ".npz"))["controldict"])
    from tests.gfortran_cases import ODD_CONTROLDICT
    out["odd"] = ODD_CONTROLDICT
    return out


@pytest.mark.skipif(not gf.available(), reason="no libgfortran.so.5 in this environment")
@pytest.mark.parametrize("deck", ["cylinder", "backstep", "room", "odd"])
def test_controldict_parser_equals_the_runtime_namelist_reader(deck, tmp_path):
    """read_settings (lib/global.f90:47-62) executed by libgfortran vs pixelflow_b200.controldict.parse_controldict"""
    from pixelflow_b200.controldict import parse_controldict
    text = _decks()[deck]
    path = tmp_path / "controlDict.txt"
    path.write_text(text)
    r = gf.read_settings(str(path))
    assert r["rc"] == 0, r
    cd = parse_controldict(text)
    for n in gf.REAL_NAMES + gf.INT_NAMES + ("output_folder", "csv_file"):
        assert getattr(cd, n) == r[n], (n, getattr(cd, n), r[n])


@pytest.mark.skipif(not gf.available(), reason="no libgfortran.so.5 in this environment")
@pytest.mark.parametrize("deck", ["cylinder", "backstep", "room", "odd"])
def test_driver_reads_and_echoes_the_controldict_like_the_reference(deck, tmp_path):
    """the C++ twin driver's namelist reader + header echo (`--echo-settings`, no GPU) against the runtime reading the
    same file and writing the same `write(*,*)` statements (lib/global.f90:66-90)"""
    from tests.gfortran_cases import echo_records
    exe = os.path.join(ROOT, "pixelflow_b200", "driver", "pixelflow_driver")
    if not os.path.exists(exe):
        from pixelflow_b200 import build
        build.build_library()
        build.build_drivers()
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "controlDict.txt").write_text(_decks()[deck])
    run = subprocess.run([exe, "--echo-settings", "--project", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, run.stderr
    r = gf.read_settings(str(tmp_path / "config" / "controlDict.txt"))
    want = [gf.list_write(*rec).decode().rstrip("\n") for rec in echo_records(r)]
    got = run.stdout.split("\n")[:-1]
    assert got == want, [(a, b) for a, b in zip(got, want) if a != b][:5]


@pytest.mark.skipif(not gf.available(), reason="no libgfortran.so.5 in this environment")
def test_runtime_rejects_what_the_parser_rejects(tmp_path):
    """groups out of order are not found by the sequential READs; an unknown object is an error"""
    from pixelflow_b200.controldict import parse_controldict
    good = _decks()["room"]
    swapped = good.replace("&file_control", "&TMP").replace("&grid_control", "&file_control").replace("&TMP", "&grid_control") \
                  .replace("istep_out = 1001", "istep_TMP").replace("istep_max = 2000", "istep_out = 1001").replace("istep_TMP", "istep_max = 2000")
    unknown = good.replace("xlambda = 0.000000", "xlambda = 0.000000\nviscosity = 1.0")
    for text in (swapped, unknown):
        path = tmp_path / "c.txt"
        path.write_text(text)
        assert gf.read_settings(str(path))["rc"] != 0
        with pytest.raises((ValueError, KeyError)):
            parse_controldict(text)
