"""STL -> porosity on the GPU: pf_stl_signed_distance against the CPU checker, bit for bit, and the
whole tool (pixelflow_b200/stl2poro.py) on the reference's own sample job (stl2poro.py:8-14)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def meshes():
    return np.load(os.path.join(HERE, "golden", "stl_meshes.npz"))


def test_signed_distance_equals_the_checker_bit_for_bit(meshes):
    from pixelflow_b200 import stl2poro as S
    from tests.test_stl2poro import oracle_sdf
    rng = np.random.default_rng(5)
    # the sphere: grid points of the reference's sample job (float32-rounded centres), plus points on and near facets
    tri = meshes["sphere"]
    pts = np.concatenate([rng.uniform(-2.5, 2.5, (20000, 3)),
                          (tri[rng.integers(0, len(tri), 3000)].astype(np.float64) * rng.dirichlet([1, 1, 1], 3000)[:, :, None]).sum(1),
                          tri.reshape(-1, 3)[::7].astype(np.float64) * 1.001])
    a, b = S.calculate_sdf(tri, pts), oracle_sdf(tri, pts)
    assert np.array_equal(a, b)
    # random triangle soups (open, self-intersecting, with a degenerate and a duplicated triangle): the same bits
    for seed in range(4):
        r = np.random.default_rng(100 + seed)
        soup = r.normal(0, 1, (300, 3, 3)).astype(np.float32)
        soup[5] = soup[4]
        soup[9, 2] = soup[9, 1]                     # zero area
        p = r.normal(0, 1.5, (5000, 3))
        assert np.array_equal(S.calculate_sdf(soup, p), oracle_sdf(soup, p))
    # the dragon (67,116 triangles): a few thousand points through the checker
    tri = meshes["dragon"]
    lo, hi = tri.reshape(-1, 3).min(0).astype(np.float64), tri.reshape(-1, 3).max(0).astype(np.float64)
    p = rng.uniform(lo - 0.02, hi + 0.02, (3000, 3))
    assert np.array_equal(S.calculate_sdf(tri, p), oracle_sdf(tri, p))


def test_reference_sample_job_and_csv(meshes, tmp_path):
    """main() of the reference's tool: sphere.stl, bounds_factor [2, 4, 1.5, 1.5, 1.5, 1.5], grid 60, axis 0,
    thickness 1.5 -> output.csv.  Against the checker's distances (same porosity bits) and the analytic sphere."""
    from pixelflow_b200 import stl2poro as S
    from tests.test_stl2poro import oracle_sdf
    tri = meshes["sphere"]
    stl = tmp_path / "sphere.stl"
    with open(stl, "wb") as f:
        f.write(b"\0" * 80 + np.uint32(len(tri)).tobytes())
        rec = np.zeros(len(tri), dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
        rec["v"] = tri
        f.write(rec.tobytes())
    poro = S.process_stl_file_three_axis(str(stl), [2.0, 4.0, 1.5, 1.5, 1.5, 1.5], 60, 0, 1.5)
    nx, ny, nz = poro.shape
    assert nx == 60 and ny in (30, 31) and nz in (30, 31)
    b = S.ratio_margin_to_bounds_for_three_axis(np.array(S.get_bounds(tri)), [2.0, 4.0, 1.5, 1.5, 1.5, 1.5])
    pitch, mp, mins = S.calculate_pitch_and_mins(b, 60, 0)
    c = S.cell_centers([nx, ny, nz], mp, mins)
    want = 0.5 * np.array([np.tanh(0)]) + 0.5     # placeholder to keep numpy's tanh out of the comparison below
    del want
    import math
    d = oracle_sdf(tri, c)
    ref = np.array([0.5 * math.tanh(v / (1.5 * pitch)) + 0.5 for v in d.ravel()]).reshape(d.shape).transpose(2, 1, 0)
    assert np.array_equal(poro, ref)
    r = np.sqrt((c ** 2).sum(-1)).transpose(2, 1, 0)
    assert np.abs(poro - (0.5 * np.tanh((r - 1.0) / (1.5 * pitch)) + 0.5)).max() < 3e-3
    out = tmp_path / "output.csv"
    S.save_3d_array_to_csv(str(out), poro.shape, poro)
    lines = out.read_bytes().split(b"\r\n")
    assert lines[0] == f"{nx},{ny},{nz}".encode() and len(lines) == nx * ny * nz + 2 and lines[-1] == b""
    assert lines[1] == f"1,1,1,{poro[0, 0, 0]:.6E}".encode() and lines[2].startswith(b"2,1,1,")
    # ... which is a porosity file the solver's own reader takes (lib/grid.f90:281-294: list-directed read)
    from pixelflow_b200.api import parse_porosity_csv
    body = out.read_bytes().split(b"\r\n", 1)[1]
    eps, nrec = parse_porosity_csv(body, nx, ny, nz, threshold=1e-6)
    assert nrec == nx * ny * nz
    rounded = np.array([float(f"{v:.6E}") for v in poro.ravel()]).reshape(poro.shape)
    assert np.array_equal(eps[1:-1, 1:-1, 1:-1], np.maximum(rounded, 1e-6).transpose(2, 1, 0))


def test_dragon_stl_config(meshes):
    """(also: the voxel fixture of configs[3], tests/golden/dragon_voxels_256.npz, made by z-ray parity in
    tests/golden/make_dragon.py, is the solid region of this distance everywhere beyond a cell from the surface)
    the dragon of BASELINE configs[3] through the STL route: porosity = tanh profile of the exact signed distance to
    the reference's 67,116 triangles at every cell centre.  64^3 against the checker bit for bit and two steps against
    the oracle; 256^3 (16.8 M points x 67 k triangles on the GPU): the body is the voxel fixture's, the slab split of the
    bench gives the planes of the whole grid, and two steps match the oracle bit for bit."""
    import time
    from oracle import oracle_c
    from pixelflow_b200 import Solver, stl2poro as S, workloads as wl
    from tests.test_gpu_voxel2poro import _dragon_case
    from tests.test_stl2poro import oracle_sdf
    # 64^3: distances against the CPU checker
    N = 64
    tri = wl.dragon_triangles_in_cells(meshes["dragon"], N)
    eps = wl.porosity_from_stl(tri, N)
    c = np.arange(N) + 0.5
    rng = np.random.default_rng(3)
    idx = rng.integers(0, N, (4000, 3))
    pts = np.stack([c[idx[:, 2]], c[idx[:, 1]], c[idx[:, 0]]], axis=1)
    d = oracle_sdf(tri, pts)
    want = np.maximum(0.5 * np.tanh(d / 1.5) + 0.5, 1e-6)
    assert np.array_equal(eps[idx[:, 0] + 1, idx[:, 1] + 1, idx[:, 2] + 1], want)
    assert eps.shape == (N + 2, N + 2, N + 2) and eps.min() >= 1e-6 and (eps < 0.3).sum() > 100   # a thin body at 64^3
    assert np.array_equal(eps[0], eps[N]) and np.array_equal(eps[:, 0], eps[:, N]) and np.array_equal(eps[:, :, 0], eps[:, :, 1])
    # a z-slab of it is the planes of the whole grid, ghost planes periodic
    sl = wl.porosity_from_stl(tri, N, k_first=1, k_count=16)
    assert np.array_equal(sl[1:17], eps[1:17]) and np.array_equal(sl[0], eps[N]) and np.array_equal(sl[17], eps[17])
    # 256^3
    N = 256
    tri = wl.dragon_triangles_in_cells(meshes["dragon"], N)
    t0 = time.perf_counter()
    eps = wl.porosity_from_stl(tri, N)
    print(f"dragon 256^3 through the STL route: {time.perf_counter() - t0:.1f} s")
    occ = wl.load_occupancy(os.path.join(HERE, "golden", "dragon_voxels_256.npz"))     # [x][y][z], 1 = fluid
    solid_vox = occ.transpose(2, 1, 0) < 0.5
    inner = eps[1:-1, 1:-1, 1:-1]
    clear = np.abs(inner - 0.5) > 0.3            # more than ~half a cell away from the surface
    differ = ((inner < 0.5) != solid_vox)[clear]
    print(f"cells more than a cell from the surface where the z-ray voxel model and the distance disagree: {int(differ.sum())}")
    assert clear.mean() > 0.95 and differ.mean() < 1e-4
    kw = _dragon_case(256, 10)
    P = oracle_c.make_params(m=N, n=N, l=N, **kw)
    oc = oracle_c.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    oc.initialise()
    err_o = oc.step(2)
    s = Solver("ibm3_uniform", N, N, N, **kw)
    s.set_porosity(eps)
    s.initial_conditions()
    err_g = s.step(2)
    u, v, w, p = s.download()
    s.close()
    assert np.array_equal(err_o, err_g)
    for a, b in ((u, oc.u), (v, oc.v), (w, oc.w), (p, oc.p)):
        assert np.array_equal(a, b)
