"""voxel -> porosity on CPU: the restated scipy.ndimage.convolve (oracle/pf_oracle.c) against golden
vectors produced by the REFERENCE's own code (tests/golden/make_voxel2poro.py runs tools/voxel2poro/voxel2poro.py
functions + scipy in the build container).  This row of the oracle is pinned."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "voxel2poro.npz"))


def test_kernel_matches_the_reference_kernel(golden):
    """create_tanh_kernel (voxel2poro.py:189-197): same bytes as the reference's kernel for thickness 1.5"""
    from pixelflow_b200.voxel2poro import create_tanh_kernel
    k = create_tanh_kernel(thickness=1.5)
    assert k.shape == (43, 43, 43) and k.dtype == np.float64
    assert np.array_equal(k[21, 21, :], golden["sphere32_kernel_centre_row"])
    digest = np.frombuffer(hashlib.sha256(k.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, golden["sphere32_kernel_sha256"])
    assert create_tanh_kernel(0.5).shape == (15, 15, 15) and create_tanh_kernel(2.0).shape == (57, 57, 57)


@pytest.mark.parametrize("case", ["grey", "box", "sphere32"])
def test_oracle_convolution_equals_the_reference_output(oracle, golden, case):
    """bit-exact float32: the reference's sample (32^3 sphere, 43^3 kernel wider than the box), a non-cubic
    binary box with a thin interface, grey voxels"""
    from pixelflow_b200.voxel2poro import create_tanh_kernel
    a = golden[case + "_in"].astype(np.float32)
    out = oracle.convolve3d_nearest(a, create_tanh_kernel(float(golden[case + "_thickness"])))
    assert out.dtype == np.float32
    assert np.array_equal(out, golden[case + "_out"])
    if case == "sphere32":   # a porosity field: 1 far from the solid, ~0 deep inside, 0.5-ish at the surface
        ref = golden[case + "_out"]
        assert ref.max() == 1.0 and ref.min() < 1e-4 and a.sum() == 32 ** 3 - 7208


def test_edge_mode_is_nearest_not_zero(oracle):
    """a constant field stays constant under a normalised kernel only if the edges replicate"""
    from pixelflow_b200.voxel2poro import create_tanh_kernel
    out = oracle.convolve3d_nearest(np.full((5, 4, 6), 0.75, np.float32), create_tanh_kernel(0.5))
    assert np.allclose(out, 0.75, rtol=0, atol=1e-7)


def test_write_porosity_format(tmp_path):
    """voxel2poro.py:200-210: `m,n,l` header, `i, j, k, %.10f` rows, i fastest, k outermost"""
    from pixelflow_b200.voxel2poro import write_porosity
    d = np.arange(24, dtype=np.float32).reshape(2, 3, 4) / 7
    f = tmp_path / "p.csv"
    write_porosity(d, str(f))
    lines = f.read_text().splitlines()
    assert lines[0] == "2,3,4" and len(lines) == 25
    assert lines[1] == f"1, 1, 1, {d[0, 0, 0]:.10f}" and lines[2] == f"2, 1, 1, {d[1, 0, 0]:.10f}"
    assert lines[3] == f"1, 2, 1, {d[0, 1, 0]:.10f}" and lines[-1] == f"2, 3, 4, {d[1, 2, 3]:.10f}"


def test_bitmap_stack_mapping(tmp_path):
    """load_bitmap_image (voxel2poro.py:56-65): 0 -> 1.0 (fluid), 128 and 255 -> 0.0 (solid); slice i -> [:, :, i]"""
    from PIL import Image
    from pixelflow_b200.voxel2poro import load_bitmap_stack
    dim = 4
    rng = np.random.default_rng(3)
    vox = rng.choice(np.array([0, 128, 255], dtype=np.uint8), size=(dim, dim, dim))
    for i in range(dim):
        Image.fromarray(vox[i], "L").save(tmp_path / f"img_{i:05d}.bmp")
    arr = load_bitmap_stack(str(tmp_path), dim)
    assert arr.dtype == np.float32 and arr.shape == (dim, dim, dim)
    for i in range(dim):
        assert np.array_equal(arr[:, :, i], (vox[i] == 0).astype(np.float32))


def test_dragon_fixture_is_a_closed_body():
    """tests/golden/dragon_voxels_*.npz (make_dragon.py): a compact solid away from the domain faces"""
    from pixelflow_b200 import workloads as wl
    for n, cells in ((64, 791), (256, 50181)):
        occ = wl.load_occupancy(os.path.join(HERE, "golden", f"dragon_voxels_{n}.npz"))
        assert occ.shape == (n, n, n) and occ.dtype == np.float32
        solid = occ == 0
        assert int(solid.sum()) == cells
        idx = np.nonzero(solid)
        assert all(a.min() > n // 8 and a.max() < n - n // 8 for a in idx)
