"""Porosity CSV records parsed on the GPU (pf_parse_porosity_csv, csrc/pf_ingest.cu) against Python's correctly
rounded float() on the same text: the formats the reference's tools write (stl2poro: csv.writer + '.6E',
voxel2poro: 'i, j, k, %.10f'), list-directed oddities, and the records that need the extended-precision path."""
import csv
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected(records, m, n, l, threshold):
    shape = (l + 2, n + 2, m + 2) if l else (n + 2, m + 2)
    e = np.zeros(shape)
    for x, y, z, txt in records:
        v = max(float(txt.lower().replace("d", "e")), threshold)
        if l:
            e[z, y, x] = v
        else:
            e[y, x] = v
    return e


def test_stl2poro_and_voxel2poro_formats():
    from pixelflow_b200 import parse_porosity_csv
    m, n, l = 13, 7, 5
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.random(m * n * l - 6), [0.0, 1.0, 1e-9, 0.5, 1e-6, 0.9999999]])
    idx = [(i, j, k) for k in range(1, l + 1) for j in range(1, n + 1) for i in range(1, m + 1)]
    # tools/stl2poro/stl2poro.py:75-88: csv.writer rows [ix+1, iy+1, iz+1, format(v, '.6E')] (CRLF line ends)
    buf = io.StringIO()
    w = csv.writer(buf)
    rec = []
    for (i, j, k), v in zip(idx, vals):
        w.writerow([i, j, k, format(v, ".6E")])
        rec.append((i, j, k, format(v, ".6E")))
    got, cnt = parse_porosity_csv(buf.getvalue().encode(), m, n, l, 1e-6)
    assert cnt == m * n * l and np.array_equal(got, _expected(rec, m, n, l, 1e-6))
    # tools/voxel2poro/voxel2poro.py:200-210: "i, j, k, %.10f", shuffled (records carry their indices)
    order = rng.permutation(len(idx))
    rec = [(*idx[q], f"{np.float32(vals[q]):.10f}") for q in order]
    text = "".join(f"{i}, {j}, {k}, {t}\n" for i, j, k, t in rec)
    got, cnt = parse_porosity_csv(text.encode(), m, n, l, 1e-6)
    assert cnt == m * n * l and np.array_equal(got, _expected(rec, m, n, l, 1e-6))


def test_list_directed_oddities_and_long_numbers():
    from pixelflow_b200 import parse_porosity_csv
    m, n, l = 4, 3, 2
    texts = ["1.0d0", "5.000000D-01", "+0.25", ".125", "1.", "3.0E-2", "1e-30", "0.12345678901234567890", "7",
             "0.1", "0.30000000000000004", "2.5e-1", "1.7976931348623157e+308", "4.9e-324", "123456789012345678",
             "0.000001", "9.999999E-01", "1.000000E+00", "0.333333333333333314829616256247", "1.0E+0", "6.02214076e23",
             "1e22", "1e23", "8.5e-23"]
    idx = [(i, j, k) for k in range(1, l + 1) for j in range(1, n + 1) for i in range(1, m + 1)]
    rec = [(*idx[q], t) for q, t in enumerate(texts)]
    lines = [f"{i},{j},{k},{t}" for i, j, k, t in rec]
    lines[3] = f"  {rec[3][0]} ,\t{rec[3][1]}   {rec[3][2]} , {rec[3][3]}  "     # blanks, tabs, mixed separators
    lines[5] = f"{rec[5][0]}.0, {rec[5][1]}.000, {rec[5][2]}., {rec[5][3]}\r"    # indices written as reals, CR
    text = "\n".join(lines[:10]) + "\n\n   \n" + "\n".join(lines[10:])           # blank lines, no final newline
    got, cnt = parse_porosity_csv(text.encode(), m, n, l, 0.0)
    assert cnt == len(texts)
    want = _expected(rec, m, n, l, 0.0)
    assert np.array_equal(got, want), [(t, a, b) for (_, _, _, t), a, b in
                                       zip(rec, got[1:-1, 1:-1, 1:-1].ravel(), want[1:-1, 1:-1, 1:-1].ravel()) if a != b]


def test_2d_file_and_preserved_cells():
    from pixelflow_b200 import parse_porosity_csv
    m, n = 6, 4
    rec = [(i, j, 1, f"{(i * 7 + j) / 31:.6E}") for j in range(1, n + 1) for i in range(1, m + 1) if (i + j) % 3]
    text = "".join(f"{i},{j},{k},{t}\n" for i, j, k, t in rec)
    out = np.full((n + 2, m + 2), -1.0)
    got, cnt = parse_porosity_csv(text.encode(), m, n, 0, 1e-6, out=out)
    want = _expected(rec, m, n, 0, 1e-6)
    mask = np.zeros_like(out, bool)
    for i, j, _, _ in rec:
        mask[j, i] = True
    assert cnt == len(rec) and np.array_equal(got[mask], want[mask]) and (got[~mask] == -1.0).all()


@pytest.mark.parametrize("bad", ["1,1,1,abc\n", "1,1,9,0.5\n", "-1,1,1,0.5\n", "4,1,1,0.5\n", "1,1,1\n"])
def test_bad_records_fail(bad):
    from pixelflow_b200 import PixelFlowError, parse_porosity_csv
    with pytest.raises(PixelFlowError, match="record"):
        parse_porosity_csv(("1,1,1,0.5\n" + bad).encode(), 2, 2, 2)


def test_trailing_items_are_ignored_like_list_directed_input():
    """`read(52,*) x, y, z, poro_val` stops after four items (libgfortran: tests/test_gfortran_io.py)"""
    from pixelflow_b200 import parse_porosity_csv
    got, cnt = parse_porosity_csv(b"1,1,1,0.5,7\n2,1,1,0.25 ! comment\n", 2, 2, 2, 1e-6)
    assert cnt == 2 and got[1, 1, 1] == 0.5 and got[1, 1, 2] == 0.25


def test_large_file_roundtrip():
    """a 96x64x48 file (295 k records, ~7 MB of text) through the parser: every value exact"""
    from pixelflow_b200 import parse_porosity_csv
    m, n, l = 96, 64, 48
    rng = np.random.default_rng(5)
    vals = rng.random((l, n, m))
    k, j, i = np.meshgrid(np.arange(1, l + 1), np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    txt = np.char.mod("%.6E", vals.ravel())
    lines = np.char.add(np.char.add(np.char.add(np.char.add(np.char.add(np.char.add(
        np.char.mod("%d", i.ravel()), ","), np.char.mod("%d", j.ravel())), ","), np.char.mod("%d", k.ravel())), ","), txt)
    text = ("\n".join(lines.tolist()) + "\n").encode()
    got, cnt = parse_porosity_csv(text, m, n, l, 1e-6)
    assert cnt == m * n * l
    assert np.array_equal(got[1:-1, 1:-1, 1:-1], np.maximum(txt.astype(np.float64).reshape(l, n, m), 1e-6))


def test_the_reference_loop_semantics():
    """lib/grid.f90:281-294 executes exactly m*n*l READs into porosity(0:md,0:nd,0:ld): records after the first m*n*l
    non-blank lines are never read (whatever they contain), index 0 and m+1 address halo cells, and of several records
    for one cell the last one read wins"""
    from pixelflow_b200 import parse_porosity_csv
    m, n, l = 3, 2, 2
    N = m * n * l
    idx = [(i, j, k) for k in range(1, l + 1) for j in range(1, n + 1) for i in range(1, m + 1)]
    lines = [f"{i},{j},{k},{0.01 * q + 0.1:.6E}" for q, (i, j, k) in enumerate(idx)]
    # trailing records: a valid one for a cell already set, and garbage -- both beyond the m*n*l READs
    text = "\n".join(lines) + "\n1,1,1,0.999\nthis is not a record\n"
    got, cnt = parse_porosity_csv(text.encode(), m, n, l, 1e-6)
    assert cnt == N and got[1, 1, 1] == float(f"{0.1:.6E}")
    # blank lines do not count as records
    text2 = "\n\n" + "\n   \n".join(lines) + "\n\n9,9,9,x\n"
    got2, cnt2 = parse_porosity_csv(text2.encode(), m, n, l, 1e-6)
    assert cnt2 == N and np.array_equal(got2, got)
    # halo indices are legal storage
    lines3 = list(lines)
    lines3[0] = "0,1,1,0.25"
    lines3[1] = f"{m + 1},{n + 1},{l + 1},0.75"
    got3, cnt3 = parse_porosity_csv(("\n".join(lines3) + "\n").encode(), m, n, l, 1e-6)
    assert cnt3 == N and got3[1, 1, 0] == 0.25 and got3[l + 1, n + 1, m + 1] == 0.75 and got3[1, 1, 1] == 0.0


def test_duplicate_records_the_last_one_read_wins():
    from pixelflow_b200 import parse_porosity_csv
    m, n, l = 40, 30, 20
    N = m * n * l
    rng = np.random.default_rng(11)
    # N records over only N/4 distinct cells, in random order: every cell is written about four times
    cells = rng.integers(0, N // 4, N)
    vals = rng.random(N)
    k, rem = np.divmod(cells, m * n)
    j, i = np.divmod(rem, m)
    text = "".join(f"{a + 1},{b + 1},{c + 1},{v:.6E}\n" for a, b, c, v in zip(i, j, k, vals))
    want = np.zeros((l + 2, n + 2, m + 2))
    for a, b, c, v in zip(i, j, k, vals):                      # sequential reads: later records overwrite earlier ones
        want[c + 1, b + 1, a + 1] = max(float(f"{v:.6E}"), 1e-6)
    for _ in range(3):                                          # (an unordered scatter would differ from run to run)
        got, cnt = parse_porosity_csv(text.encode(), m, n, l, 1e-6)
        assert cnt == N and np.array_equal(got, want)
    # ... also when the winning (or a losing) record needs the extended-precision path
    text2 = "1,1,1,0.5\n1,1,1,0.12345678901234567890123\n2,1,1,0.12345678901234567890123\n2,1,1,0.25\n"
    got, cnt = parse_porosity_csv(text2.encode(), 2, 2, 1, 1e-6)
    assert got[1, 1, 1] == float("0.12345678901234567890123") and got[1, 1, 2] == 0.25
