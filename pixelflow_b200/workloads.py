"""Deterministic synthetic inputs and the reference's grid arithmetic (host side, numpy only).

* `grid_spacing` restates lib/grid.f90:297-300 (dx = width/real(m-1) ..., dt = time/real(istep_max)).
* `porosity_halo_*` restate the porosity halo rules of lib/grid.f90 (:92-106 2D, :215-243 wall,
  :349-378 y/z-periodic) -- input preparation, not part of the per-step hot path.
* `porous_channel` is the synthetic S1/S2/S3 workload of SURVEY.md 8(d): a periodic lattice of rods
  along x with a tanh porosity profile of thickness 1.5 cells, radius modulated along x so that
  all three porosity-gradient terms of the predictor are exercised.  No RNG.
"""
from __future__ import annotations

import math

import numpy as np


def grid_spacing(width, height, depth, time, istep_max, m, n, l=1):
    dx = width / float(m - 1)
    dy = height / float(n - 1)
    dz = depth / float(l - 1) if l > 1 else 1.0
    dt = time / float(istep_max)
    return dx, dy, dz, dt


def porosity_halo_3d_periodic(e):
    """lib/grid.f90:349-378 on an (l+2, n+2, m+2) array whose interior is filled (in place)."""
    l, n, m = (s - 2 for s in e.shape)
    e[1:l + 2, 1:n + 2, 0] = e[1:l + 2, 1:n + 2, 1]
    e[1:l + 2, 1:n + 2, m + 1] = e[1:l + 2, 1:n + 2, m]
    e[:, 0, :] = e[:, n, :]
    e[:, n + 1, :] = e[:, 1, :]
    e[0, :, :] = e[l, :, :]
    e[l + 1, :, :] = e[1, :, :]
    return e


def porosity_halo_3d_wall(e):
    """lib/grid.f90:215-243"""
    l, n, m = (s - 2 for s in e.shape)
    e[:, :, 0] = e[:, :, 1]
    e[:, :, m + 1] = e[:, :, m]
    e[:, 0, :] = e[:, 1, :]
    e[:, n + 1, :] = e[:, n, :]
    e[0, :, :] = e[1, :, :]
    e[l + 1, :, :] = e[l, :, :]
    return e


def porosity_halo_2d(e):
    """lib/grid.f90:92-106"""
    n, m = (s - 2 for s in e.shape)
    e[1:n + 2, 0] = e[1:n + 2, 1]
    e[1:n + 2, m + 1] = e[1:n + 2, m]
    e[0, :] = e[n, :]
    e[n + 1, :] = e[1, :]
    return e


def with_halos(eps_interior, case):
    """interior porosity [l,n,m] (or [n,m]) -> the padded array the reference's grid routine builds for `case`
    (zero-initialised storage, interior, then the halo rules of lib/grid.f90)."""
    a = np.asarray(eps_interior, dtype=np.float64)
    e = np.zeros(tuple(s + 2 for s in a.shape))
    if a.ndim == 2:
        e[1:-1, 1:-1] = a
        return porosity_halo_2d(e)
    e[1:-1, 1:-1, 1:-1] = a
    return porosity_halo_3d_wall(e) if case == "ibm3_air_condition" else porosity_halo_3d_periodic(e)


def _lattice_distance(nj, nk, pitch):
    """distance (in cells) from cell (j,k) to the nearest centre of a square lattice of pitch `pitch`
    whose centres sit at (pitch/2 + a*pitch, pitch/2 + b*pitch); periodic by construction."""
    j = np.arange(1, nj + 1, dtype=np.float64)
    k = np.arange(1, nk + 1, dtype=np.float64)
    dj = np.mod(j - 0.5, pitch) - pitch / 2.0
    dk = np.mod(k - 0.5, pitch) - pitch / 2.0
    return np.sqrt(dk[:, None] ** 2 + dj[None, :] ** 2)  # [k, j]


def porous_channel(m, n, l, *, pitch=64, radius=16.0, wobble=4.0, threshold=1.0e-6, thickness=1.5,
                   k_first=1, k_count=None, with_halo=True, out=None):
    """Porosity of the synthetic porous channel, shape (k_count+2, n+2, m+2), halos per
    lib/grid.f90:349-378.  `k_first`/`k_count` select a z-slab (1-based global planes) so that a
    rank can build only what it owns (+1 ghost plane each side, periodic in z).

    eps(i,j,k) = max(threshold, 0.5*tanh(d/thickness)+0.5),  d = dist_to_nearest_rod_axis(j,k) - R(i),
    R(i) = radius + wobble*sin(2*pi*(i-1)/256).
    """
    pitch = min(pitch, n, l)
    k_count = l if k_count is None else k_count
    dist = _lattice_distance(n, l, pitch)  # [l, n], global
    ks = (np.arange(k_first - 1, k_first + k_count + 1) - 1) % l  # global k-1 incl. ghosts, periodic
    dist = dist[ks]  # [k_count+2, n]
    i = np.arange(1, m + 1, dtype=np.float64)
    R = radius * min(1.0, pitch / 64.0) + wobble * min(1.0, pitch / 64.0) * np.sin(2.0 * math.pi * (i - 1.0) / 256.0)
    e = np.zeros((k_count + 2, n + 2, m + 2)) if out is None else out
    # plane by plane to bound temporaries
    for kk in range(k_count + 2):
        d = dist[kk][:, None] - R[None, :]
        e[kk, 1:n + 1, 1:m + 1] = np.maximum(threshold, 0.5 * np.tanh(d / thickness) + 0.5)
    if with_halo:
        # x zero-gradient then periodic y (z ghosts were generated periodically above)
        e[:, 1:n + 1, 0] = e[:, 1:n + 1, 1]
        e[:, 1:n + 1, m + 1] = e[:, 1:n + 1, m]
        e[:, 0, :] = e[:, n, :]
        e[:, n + 1, :] = e[:, 1, :]
        if k_count == l and k_first == 1:
            # exact lib/grid.f90 corner semantics on a full array
            inner = e[1:-1, 1:-1, 1:-1].copy()
            e[...] = 0.0
            e[1:-1, 1:-1, 1:-1] = inner
            porosity_halo_3d_periodic(e)
    return e


def room_like(m, n, l, *, threshold=1.0e-6, thickness=1.5):
    """Small air-condition style porosity: a box-shaped solid in the middle of the room, open
    (eps ~ 1) patches on the top (inlet) and south (outlet) faces; zero-gradient halos."""
    k, j, i = np.meshgrid(np.arange(1, l + 1), np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    # signed distance (cells) to a centred box of half-sizes (m/6, n/6, l/6)
    qx = np.abs(i - (m + 1) / 2.0) - m / 6.0
    qy = np.abs(j - (n + 1) / 2.0) - n / 6.0
    qz = np.abs(k - (l + 1) / 2.0) - l / 6.0
    outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
    inside = np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0.0)
    d = outside + inside
    # walls: distance to the nearest face, solid outside except the two openings
    wall_d = np.minimum.reduce([i - 0.5, m + 0.5 - i, j - 0.5, n + 0.5 - j, k - 0.5, l + 0.5 - k]).astype(float)
    open_top = (k > l - 3) & (np.abs(i - (m + 1) / 2.0) < m / 8.0 + 0.5) & (np.abs(j - (n + 1) / 2.0) < n / 8.0 + 0.5)
    open_south = (j < 4) & (np.abs(i - (m + 1) / 2.0) < m / 8.0 + 0.5) & (np.abs(k - (l + 1) / 4.0) < l / 8.0 + 0.5)
    wall_d = np.where(open_top | open_south, 10.0, wall_d - 1.0)
    d = np.minimum(d, wall_d)
    e = np.zeros((l + 2, n + 2, m + 2))
    e[1:-1, 1:-1, 1:-1] = np.maximum(threshold, 0.5 * np.tanh(d / thickness) + 0.5)
    return porosity_halo_3d_wall(e)


def cylinder_2d(m, n, *, cx=0.25, cy=0.5, radius_cells=None, threshold=1.0e-6, thickness=1.5):
    """2D cylinder porosity in the style of tools/cylinder (tanh profile of the signed distance)."""
    radius_cells = n / 16.0 if radius_cells is None else radius_cells
    j, i = np.meshgrid(np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    d = np.sqrt((i - 1 - cx * (m - 1)) ** 2 + (j - 1 - cy * (n - 1)) ** 2) - radius_cells
    e = np.zeros((n + 2, m + 2))
    e[1:-1, 1:-1] = np.maximum(threshold, 0.5 * np.tanh(d / thickness) + 0.5)
    return porosity_halo_2d(e)


# the BASELINE.json configurations that are synthetic
WORKLOADS = {
    # name: (m, n, l, width, height, depth)
    "s1_1024x512x512": (1024, 512, 512, 1.023, 0.511, 0.511),
    "s2_256": (256, 256, 256, 0.255, 0.255, 0.255),
    "s3_64": (64, 64, 64, 0.063, 0.063, 0.063),
}
# time/istep_max -> dt = 5e-5.  SURVEY.md 8(d) proposed istep_max=100 (dt = 2e-4), but with dx = 1e-3 and
# nu = 1e-3 that gives a diffusion number nu*dt/dx^2 = 0.2 > 1/6: the explicit predictor is unstable in
# 3D and the run blows up after ~30 steps (checked with the CPU restatement).  dt = 5e-5 (0.05) is stable
# for hundreds of steps; the arithmetic per step is identical, so throughput is unaffected.
CHANNEL_PHYSICS = dict(xnue=1.0e-3, xlambda=0.0, density=1.0, time=0.02, istep_max=400, inlet_velocity=1.0,
                       outlet_pressure=0.0, AoA=0.0, thickness=1.5, threshold=1.0e-6, nonslip=True,
                       iter_max=100, relux_factor=1.7)

# BASELINE configs[3]: "stanford-dragon 3D ibm3 porosity from voxel2poro, 256^3 grid".  The voxel model comes from
# tests/golden/dragon_voxels_256.npz (made from the reference's dragon.stl by tests/golden/make_dragon.py); the
# porosity is computed from it by the GPU tanh filter, as the reference's pipeline does with scipy.
WORKLOADS["dragon_256"] = (256, 256, 256, 0.255, 0.255, 0.255)
WORKLOADS["dragon_64"] = (64, 64, 64, 0.063, 0.063, 0.063)


# The same dragon through the STL route: signed distance to the surface at the cell centres, then the
# reference's profile 0.5*tanh(d / (thickness*pitch)) + 0.5 (tools/stl2poro/stl2poro.py:87-97, :202).  Placement as in
# tests/golden/make_dragon.py (longest extent along x over 0.375 N cells, centred at (0.375, 0.5, 0.5) N, the same
# sub-cell shift), so both dragons describe one body; lengths in cells, i.e. pitch = 1.
WORKLOADS["dragon_stl_256"] = (256, 256, 256, 0.255, 0.255, 0.255)
WORKLOADS["dragon_stl_64"] = (64, 64, 64, 0.063, 0.063, 0.063)


def dragon_triangles_in_cells(triangles, N, extent=None):
    """float32 triangles [ntri][3][3] of the reference's dragon.stl -> the same mesh in grid units (float32)"""
    tri = np.asarray(triangles, dtype=np.float64)
    extent = 0.375 * N if extent is None else extent
    v = tri.reshape(-1, 3)
    lo, hi = v.min(0), v.max(0)
    order = np.argsort(-(hi - lo))
    scale = extent / (hi - lo).max()
    centre = np.array([0.375 * N, 0.5 * N, 0.5 * N])
    t = (tri[:, :, order] - 0.5 * (lo + hi)[order]) * scale + centre + np.array([0.1234567, 0.2345678, 0.3456789])
    return np.ascontiguousarray(t, dtype=np.float32)


def porosity_from_stl(triangles, N, *, thickness=1.5, threshold=1.0e-6, k_first=1, k_count=None, device=-1):
    """mesh in grid units -> porosity [k][j][i] with halos for the y/z-periodic ibm3 solver (one z-slab with its two
    ghost planes if k_first / k_count are given): pf_stl_signed_distance at the cell centres (i-0.5, j-0.5, k-0.5),
    tanh profile, max(porosity, threshold) and the halos of lib/grid.f90:349-378"""
    from .stl2poro import calculate_sdf
    k_count = N if k_count is None else k_count
    ks = (np.arange(k_first - 1, k_first + k_count + 1) - 1) % N + 1          # ghost planes: periodic images
    uniq = np.unique(ks)
    c = np.arange(N) + 0.5
    pts = np.empty((len(uniq), N, N, 3))
    pts[..., 0] = c[None, None, :]
    pts[..., 1] = c[None, :, None]
    pts[..., 2] = (uniq - 0.5)[:, None, None]
    sdf = calculate_sdf(triangles, pts, device)
    por = np.maximum(0.5 * np.tanh(sdf / thickness) + 0.5, threshold)          # [k][j][i]
    # halos of lib/grid.f90:349-378, plane by plane (they never mix planes: zero-gradient x halos, periodic y rows;
    # the z ghost planes of a slab are the periodic images of whole planes)
    planes = np.zeros((len(uniq), N + 2, N + 2))
    planes[:, 1:-1, 1:-1] = por
    planes[:, 1:N + 1, 0] = planes[:, 1:N + 1, 1]
    planes[:, 1:N + 1, N + 1] = planes[:, 1:N + 1, N]
    planes[:, 0, :] = planes[:, N, :]
    planes[:, N + 1, :] = planes[:, 1, :]
    index = {int(k): idx for idx, k in enumerate(uniq)}
    return np.ascontiguousarray(planes[[index[int(k)] for k in ks]])


def load_occupancy(path):
    """bit-packed voxel fixture -> float32 array [x][y][z], 1 = fluid, 0 = solid (voxel2poro.py:56-65)"""
    z = np.load(path)
    shape = tuple(int(v) for v in z["shape"])
    return np.unpackbits(z["packed"])[:int(np.prod(shape))].reshape(shape).astype(np.float32)


def porosity_from_occupancy(occ, *, thickness=1.5, threshold=1.0e-6, k_first=1, k_count=None, device=-1):
    """occupancy [x][y][z] -> porosity [k][j][i] with halos, for the y/z-periodic ibm3 solver:
    voxel2poro.py:31-35 on the GPU (pixelflow_b200.voxel2poro), then what lib/grid.f90 does on input:
    max(porosity, threshold) (:289) and the halos of :349-378.  `k_first`/`k_count` return one z-slab
    (+1 ghost plane each side, periodic)."""
    from .voxel2poro import voxel2poro
    por = voxel2poro(occ, thickness=thickness, device=device)          # float32 [x][y][z]
    m, n, l = por.shape
    e = np.zeros((l + 2, n + 2, m + 2))
    e[1:-1, 1:-1, 1:-1] = np.maximum(por.astype(np.float64), threshold).transpose(2, 1, 0)
    porosity_halo_3d_periodic(e)
    if k_count is None or (k_first == 1 and k_count == l):
        return e
    ks = (np.arange(k_first - 1, k_first + k_count + 1) - 1) % l + 1
    return np.ascontiguousarray(e[ks])
