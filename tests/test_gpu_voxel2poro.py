"""voxel -> porosity on the GPU (pf_convolve3d_nearest, csrc/pf_voxel.cu) against the reference's own outputs
(tests/golden/voxel2poro.npz), against the restated scipy convolution, and as the input stage of BASELINE configs[3]
(Stanford dragon, 256^3)."""
import os

import numpy as np
import pytest

from pixelflow_b200 import workloads as wl

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "voxel2poro.npz"))


@pytest.mark.parametrize("case", ["grey", "box", "sphere32"])
def test_gpu_equals_the_reference_output(golden, case):
    """bit-exact float32 against what tools/voxel2poro/voxel2poro.py + scipy produced"""
    from pixelflow_b200.voxel2poro import voxel2poro
    out = voxel2poro(golden[case + "_in"].astype(np.float32), thickness=float(golden[case + "_thickness"]))
    assert out.dtype == np.float32 and out.shape == golden[case + "_out"].shape
    assert np.array_equal(out, golden[case + "_out"]), int((out != golden[case + "_out"]).sum())


@pytest.mark.parametrize("shape,thickness", [((5, 7, 300), 0.3), ((3, 130, 129), 0.5), ((40, 33, 31), 1.0), ((1, 1, 1), 0.5)])
def test_gpu_equals_oracle_on_ragged_shapes(oracle, shape, thickness):
    """shapes that are not multiples of the 128x4 tile, rows longer than one tile, kernels wider than the box"""
    from pixelflow_b200.voxel2poro import convolve_nearest, create_tanh_kernel
    rng = np.random.default_rng(sum(shape))
    a = rng.random(shape).astype(np.float32)
    k = create_tanh_kernel(thickness)
    assert np.array_equal(convolve_nearest(a, k), oracle.convolve3d_nearest(a, k))


def test_asymmetric_kernel_is_flipped_like_convolve(oracle):
    """convolve, not correlate: an asymmetric kernel must be applied reversed (scipy semantics)"""
    from pixelflow_b200.voxel2poro import convolve_nearest
    rng = np.random.default_rng(9)
    a = rng.random((9, 8, 11)).astype(np.float32)
    k = rng.random((3, 5, 7))
    out = convolve_nearest(a, k)
    assert np.array_equal(out, oracle.convolve3d_nearest(a, k))
    pad = np.pad(a.astype(np.float64), ((1, 1), (2, 2), (3, 3)), mode="edge")
    i, j, l = 4, 3, 5
    direct = sum(k[p, q, r] * pad[i + 1 + (1 - p), j + 2 + (2 - q), l + 3 + (3 - r)]
                 for p in range(3) for q in range(5) for r in range(7))
    assert abs(out[i, j, l] - direct) < 1e-5


def test_bad_arguments_fail_loudly():
    from pixelflow_b200 import PixelFlowError
    from pixelflow_b200.voxel2poro import convolve_nearest
    with pytest.raises(PixelFlowError, match="odd"):
        convolve_nearest(np.zeros((4, 4, 4), np.float32), np.ones((2, 3, 3)))


def _dragon_case(n, iter_max):
    m = l = n
    width = {64: 0.063, 256: 0.255}[n]
    dx, dy, dz, dt = wl.grid_spacing(width, width, width, 0.02, 400, m, n, l)
    return dict(dx=dx, dy=dy, dz=dz, dt=dt, xnue=1e-3, iter_max=iter_max, inlet_velocity=1.0, outlet_pressure=0.0, AoA=0.0)


def test_dragon_64_pipeline_bit_exact(oracle):
    """voxels -> tanh filter (GPU == restated scipy, every voxel) -> 3 time steps (GPU == oracle)"""
    from pixelflow_b200 import Solver
    from pixelflow_b200.voxel2poro import create_tanh_kernel, voxel2poro
    occ = wl.load_occupancy(os.path.join(HERE, "golden", "dragon_voxels_64.npz"))
    por = voxel2poro(occ, thickness=1.5)
    assert np.array_equal(por, oracle.convolve3d_nearest(occ, create_tanh_kernel(1.5)))
    assert por.min() < 0.9 and por.max() == 1.0   # a 24-cell dragon: thinner than the 1.5-cell interface in places
    eps = wl.porosity_from_occupancy(occ)
    kw = _dragon_case(64, 50)
    P = oracle.make_params(m=64, n=64, l=64, **kw)
    oc = oracle.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    assert np.array_equal(oc.e, eps)
    oc.initialise()
    err_o = oc.step(3)
    s = Solver("ibm3_uniform", 64, 64, 64, **kw)
    s.set_porosity(eps)
    s.initial_conditions()
    err_g = s.step(3)
    u, v, w, p = s.download()
    s.close()
    assert np.array_equal(err_o, err_g) and err_o[-1] > 0
    for a, b in ((u, oc.u), (v, oc.v), (w, oc.w), (p, oc.p)):
        assert np.array_equal(a, b)
    assert np.abs(v).max() > 0   # the body deflects the flow


def test_dragon_256_config(oracle):
    """BASELINE configs[3] at full size: the 256^3 tanh filter on the GPU (spot-checked against a sequential
    numpy evaluation of 300 voxels around the surface), then two steps bit-identical to the oracle"""
    from pixelflow_b200 import Solver
    from pixelflow_b200.voxel2poro import create_tanh_kernel, voxel2poro
    occ = wl.load_occupancy(os.path.join(HERE, "golden", "dragon_voxels_256.npz"))
    por = voxel2poro(occ, thickness=1.5)
    k = create_tanh_kernel(1.5)
    k = np.where(np.abs(k) > np.finfo(np.float64).eps, k, 0.0)[::-1, ::-1, ::-1]
    rng = np.random.default_rng(5)
    near = np.argwhere((por > 0.05) & (por < 0.95))
    pts = near[rng.choice(len(near), 300, replace=False)]
    N = occ.shape[0]
    acc = np.zeros(len(pts))
    for a0 in range(43):
        z = np.clip(pts[:, 0] + a0 - 21, 0, N - 1)
        for a1 in range(43):
            y = np.clip(pts[:, 1] + a1 - 21, 0, N - 1)
            for a2 in range(43):
                x = np.clip(pts[:, 2] + a2 - 21, 0, N - 1)
                acc = acc + occ[z, y, x].astype(np.float64) * k[a0, a1, a2]
    assert np.array_equal(acc.astype(np.float32), por[pts[:, 0], pts[:, 1], pts[:, 2]])
    eps = wl.porosity_from_occupancy(occ)
    assert eps.shape == (258, 258, 258) and eps.min() >= 1e-6
    kw = _dragon_case(256, 10)
    P = oracle.make_params(m=256, n=256, l=256, **kw)
    oc = oracle.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    oc.initialise()
    err_o = oc.step(2)
    s = Solver("ibm3_uniform", 256, 256, 256, **kw)
    assert s.sor_variant == 6
    s.set_porosity(eps)
    s.initial_conditions()
    err_g = s.step(2)
    u, v, w, p = s.download()
    s.close()
    assert np.array_equal(err_o, err_g)
    for a, b in ((u, oc.u), (v, oc.v), (w, oc.w), (p, oc.p)):
        assert np.array_equal(a, b)


def test_force_log_3d_on_the_dragon(oracle):
    """output_force_log_3d (lib/output.f90:1090-1165) on the 64^3 dragon after two steps: per-cell terms exact,
    the six sums equal the serial reference sums to rounding (different summation order)"""
    from pixelflow_b200 import Solver
    occ = wl.load_occupancy(os.path.join(HERE, "golden", "dragon_voxels_64.npz"))
    eps = wl.porosity_from_occupancy(occ)
    kw = _dragon_case(64, 30)
    P = oracle.make_params(m=64, n=64, l=64, **kw)
    oc = oracle.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    oc.initialise()
    oc.step(2)
    s = Solver("ibm3_uniform", 64, 64, 64, **kw)
    s.set_porosity(eps)
    s.initial_conditions()
    s.step(2)
    radius = 0.012
    fo, fg = oc.force_log(radius), s.force_log_3d(radius)["raw"]
    s.close()
    scale = np.abs(fo[:6]).max()
    assert scale > 0 and fo[9] > 0                      # the dragon feels a drag along +x
    assert np.allclose(fg[:9], fo[:9], rtol=1e-11, atol=1e-12 * scale), (fg, fo)
    assert np.allclose(fg[9:], fo[9:], rtol=1e-11, atol=1e-12 * abs(fo[9])), (fg, fo)
    with pytest.raises(Exception, match="3D"):
        s2 = Solver("ibm2_uniform", 16, 8, dx=0.1, dy=0.1, dt=1e-3, xnue=1e-3)
        try:
            s2.force_log_3d(1.0)
        finally:
            s2.close()
