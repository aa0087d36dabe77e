"""STL -> porosity (tools/stl2poro/stl2poro.py): the CPU checker and the host-side mirror.

The reference's tool needs vtk, which the image does not have, so no output of the tool itself exists to compare with
(PARITY UNPINNED for this row; oracle/stl_oracle.c says so in its header).  What pins the restated signed distance:
the analytic distance to the sphere behind stl_files/sphere.stl, inside/outside parity against ray casting on the
dragon, and invariances of a distance field.  The CUDA kernel is then held to the checker bit for bit
(tests/test_gpu_stl2poro.py).
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def meshes():
    return np.load(os.path.join(HERE, "golden", "stl_meshes.npz"))


def oracle_sdf(tri, pts):
    from oracle import oracle_c
    oracle_c.build()
    L = C.CDLL(os.path.join(os.path.dirname(HERE), "oracle", "liboracle.so"))
    L.pfo_stl_signed_distance.argtypes = [C.POINTER(C.c_float), C.c_longlong, C.POINTER(C.c_double), C.c_longlong,
                                          C.POINTER(C.c_double)]
    tri = np.ascontiguousarray(tri, dtype=np.float32).reshape(-1, 9)
    p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    out = np.empty(len(p))
    assert L.pfo_stl_signed_distance(tri.ctypes.data_as(C.POINTER(C.c_float)), len(tri),
                                     p.ctypes.data_as(C.POINTER(C.c_double)), len(p),
                                     out.ctypes.data_as(C.POINTER(C.c_double))) == 0
    return out.reshape(np.shape(pts)[:-1])


def test_oracle_distance_to_the_sphere_is_the_analytic_one(meshes):
    """sphere.stl is an icosphere of 5120 facets on the unit sphere: |d - (|x| - 1)| <= the sagitta of a facet"""
    tri = meshes["sphere"]
    r = np.sqrt((tri.reshape(-1, 3).astype(np.float64) ** 2).sum(1))
    assert abs(r.min() - 1) < 1e-6 and abs(r.max() - 1) < 1e-6 and len(tri) == 5120
    rng = np.random.default_rng(7)
    pts = rng.uniform(-2.5, 2.5, (6000, 3))
    d = oracle_sdf(tri, pts)
    ana = np.sqrt((pts ** 2).sum(1)) - 1.0
    assert np.abs(d - ana).max() < 1.3e-3          # facet edge ~0.06 -> sagitta ~ 0.06^2/8 ... 1.2e-3 at facet centres
    far = np.abs(ana) > 2e-3
    assert ((d < 0) == (ana < 0))[far].all()
    assert (d[np.sqrt((pts ** 2).sum(1)) < 0.99] < 0).all()


def test_oracle_distance_field_properties(meshes):
    """exact on the surface, 1-Lipschitz, symmetric under the mesh's own symmetry, indifferent to the triangle order"""
    tri = meshes["sphere"]
    rng = np.random.default_rng(3)
    # points ON the surface (random barycentric combinations of facets): distance ~ 0
    idx = rng.integers(0, len(tri), 500)
    w = rng.dirichlet([1, 1, 1], 500)
    on = (tri[idx].astype(np.float64) * w[:, :, None]).sum(1)
    assert np.abs(oracle_sdf(tri, on)).max() < 1e-7
    pts = rng.uniform(-2, 2, (800, 3))
    d = oracle_sdf(tri, pts)
    q = pts + rng.normal(0, 0.05, pts.shape)
    assert (np.abs(oracle_sdf(tri, q) - d) <= np.sqrt(((q - pts) ** 2).sum(1)) + 1e-12).all()
    perm = rng.permutation(len(tri))
    assert np.allclose(oracle_sdf(tri[perm], pts), d, rtol=0, atol=1e-15)      # ties aside, the same closest point
    # a cube from 12 triangles: exact distances and signs, vertices and edges included
    c = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]   # outward
    cube = np.array([[c[a], c[b], c[cc]] for a, b, cc, dd in quads] + [[c[a], c[cc], c[dd]] for a, b, cc, dd in quads])
    probe = np.array([[0.5, 0.5, 0.5], [0.5, 0.5, 0.9], [2, 0.5, 0.5], [2, 2, 0.5], [2, 2, 2], [-1, -1, -1],
                      [0.5, 0.5, 1.25], [1.0, 1.0, 1.0], [0.25, 0.5, 0.5]])
    want = np.array([-0.5, -0.1, 1.0, math.sqrt(2), math.sqrt(3), math.sqrt(3), 0.25, 0.0, -0.25])
    got = oracle_sdf(cube, probe)
    assert np.allclose(got, want, rtol=0, atol=1e-15), got


def test_oracle_sign_agrees_with_ray_casting_on_the_dragon(meshes):
    """inside <=> an odd number of crossings of a +z ray (Moller-Trumbore, fp64), away from the surface"""
    tri = meshes["dragon"].astype(np.float64)
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    rng = np.random.default_rng(11)
    pts = rng.uniform(lo, hi, (400, 3))
    d = oracle_sdf(meshes["dragon"], pts)
    a, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    dirn = np.array([0.0, 0.0, 1.0])
    pv = np.cross(dirn, e2)
    det = (e1 * pv).sum(1)
    ok = np.abs(det) > 1e-18
    inside = np.zeros(len(pts), dtype=bool)
    for k, p in enumerate(pts):
        tv = p - a
        u = (tv * pv).sum(1) / np.where(ok, det, 1)
        qv = np.cross(tv, e1)
        v = (qv @ dirn) / np.where(ok, det, 1)
        t = (e2 * qv).sum(1) / np.where(ok, det, 1)
        inside[k] = (ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0)).sum() % 2 == 1
    clear = np.abs(d) > 2e-4 * (hi - lo).max()
    assert clear.sum() > 350 and ((d < 0) == inside)[clear].all()
    assert 0.02 < (d < 0).mean() < 0.5


def test_host_mirror_grid_and_csv(tmp_path, meshes, monkeypatch):
    """the reference's main(): sphere.stl, bounds_factor [2, 4, 1.5, 1.5, 1.5, 1.5], grid 60 along x, thickness 1.5
    (stl2poro.py:8-14) -> bounds (-4, 8) x (-3, 3)^2 (the `* 2` of :34), pitch 0.2, 60 x 30 x 30 cells, centres at
    min + (i + 1/2) * pitch - pitch/2; porosity 0.5*tanh(d/(1.5*0.2)) + 0.5 -- here with the checker's distance"""
    from pixelflow_b200 import stl2poro as S
    tri = meshes["sphere"]
    b = S.ratio_margin_to_bounds_for_three_axis(np.array(S.get_bounds(tri)), [2.0, 4.0, 1.5, 1.5, 1.5, 1.5])
    assert np.allclose(b, [-4, 8, -3, 3, -3, 3], atol=1e-6)
    pitch, mesh_pitch, mins = S.calculate_pitch_and_mins(b, 60, 0)
    assert abs(pitch - 0.2) < 1e-7 and mesh_pitch == [pitch] * 3
    dims = [math.ceil((b[i * 2 + 1] - b[i * 2]) / mesh_pitch[i // 2]) for i in range(3)]
    # (`ceil` of a quotient that is 30 up to rounding: the vertices are float32, z spans -1 .. 1 but y only
    # -0.99999994 .. 0.99999994 -- the reference's formula, rounding included, decides between 30 and 31)
    assert dims[0] == 60 and dims[1] in (30, 31) and dims[2] in (30, 31)
    c = S.cell_centers(dims, mesh_pitch, mins)
    assert c.shape == (dims[2], dims[1], 60, 3)
    assert np.abs(c[0, 0, :, 0] - (b[0] + np.arange(60) * pitch)).max() < 1e-6       # float32 storage of VTK's points
    assert (c[..., 0].astype(np.float32) == c[..., 0]).all()
    monkeypatch.setattr(S, "calculate_sdf", lambda t, p, device=-1: oracle_sdf(t, p))
    stl = tmp_path / "sphere.stl"
    with open(stl, "wb") as f:      # a binary STL round trip of the fixture
        f.write(b"\0" * 80 + np.uint32(len(tri)).tobytes())
        rec = np.zeros(len(tri), dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
        rec["v"] = tri
        f.write(rec.tobytes())
    assert np.array_equal(S.read_stl_file(str(stl)), tri)
    poro = S.process_stl_file_three_axis(str(stl), [2.0, 4.0, 1.5, 1.5, 1.5, 1.5], 60, 0, 1.5)
    assert poro.shape == tuple(dims)
    x, y, z = c[..., 0].transpose(2, 1, 0), c[..., 1].transpose(2, 1, 0), c[..., 2].transpose(2, 1, 0)
    ana = 0.5 * np.tanh((np.sqrt(x * x + y * y + z * z) - 1.0) / (1.5 * pitch)) + 0.5
    assert np.abs(poro - ana).max() < 3e-3 and poro.min() < 2e-3 and poro.max() > 0.999
    # the CSV the solver reads (lib/grid.f90:281-294): header, then ix fastest, `.6E`, CRLF like csv.writer
    small = np.ascontiguousarray(poro[:3, :2, :2])
    out = tmp_path / "output.csv"
    S.save_3d_array_to_csv(str(out), small.shape, small)
    import csv
    import io
    ref = io.StringIO(newline="")
    w = csv.writer(ref)
    w.writerow(small.shape)
    for iz in range(2):
        for iy in range(2):
            for ix in range(3):
                w.writerow([ix + 1, iy + 1, iz + 1, format(small[ix, iy, iz], ".6E")])
    assert out.read_bytes() == ref.getvalue().encode()
