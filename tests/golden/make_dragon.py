#!/usr/bin/env python
"""Voxel model of the Stanford dragon for BASELINE configs[3] ("stanford-dragon 3D ibm3 porosity from
voxel2poro, 256^3 grid").

The reference ships the surface (tools/stl2poro/stl_files/dragon.stl, binary STL) but neither the
bitmap stack that tools/voxel2poro/voxel2poro.py reads nor a voxeliser.  This script (build container
only: it reads /root/reference) voxelises the surface by z-ray parity at the cell centres of an N^3 grid and
stores the occupancy array (1 = fluid, 0 = solid: what load_bitmap_image produces, voxel2poro.py:56-65)
bit-packed in tests/golden/dragon_voxels_<N>.npz.  The porosity itself is then computed at test / bench time
by the GPU tanh filter (pixelflow_b200.voxel2poro), like the reference's pipeline does with scipy.

Placement: flow along x (array axis 0); the dragon's longest extent spans `extent` cells, centred at
(0.375 N, 0.5 N, 0.5 N).

    python tests/golden/make_dragon.py [N=256] [extent=0.375*N]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STL = "/root/reference/tools/stl2poro/stl_files/dragon.stl"


def read_binary_stl(path):
    raw = np.fromfile(path, dtype=np.uint8)
    ntri = int(raw[80:84].view("<u4")[0])
    assert raw.size == 84 + 50 * ntri, "not a binary STL"
    rec = raw[84:].reshape(ntri, 50)
    return rec[:, 12:48].copy().view("<f4").reshape(ntri, 3, 3).astype(np.float64)   # [triangle, vertex, xyz]


def voxelise(tri, N, extent):
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    scale = extent / (hi - lo).max()
    centre = np.array([0.375 * N, 0.5 * N, 0.5 * N])
    # an irrational-looking sub-cell shift keeps vertices and edges off the rays through the cell centres
    # order the axes so that the longest extent lies along x; rays run along z
    order = np.argsort(-(hi - lo))
    t = (tri[:, :, order] - 0.5 * (lo + hi)[order]) * scale + centre + np.array([0.1234567, 0.2345678, 0.3456789])
    x, y, z = t[:, :, 0], t[:, :, 1], t[:, :, 2]
    # candidate ray positions (cell centres ix+0.5, iy+0.5) inside each triangle's xy bounding box
    ix0 = np.ceil(x.min(1) - 0.5).astype(int)
    ix1 = np.floor(x.max(1) - 0.5).astype(int)
    iy0 = np.ceil(y.min(1) - 0.5).astype(int)
    iy1 = np.floor(y.max(1) - 0.5).astype(int)
    toggles = np.zeros((N, N, N + 1), dtype=np.int32)
    area = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    ok = area != 0
    for dx in range(int((ix1 - ix0).max()) + 1):
        for dy in range(int((iy1 - iy0).max()) + 1):
            ix, iy = ix0 + dx, iy0 + dy
            sel = ok & (ix <= ix1) & (iy <= iy1) & (ix >= 0) & (ix < N) & (iy >= 0) & (iy < N)
            if not sel.any():
                continue
            px, py = ix[sel] + 0.5, iy[sel] + 0.5
            xs, ys, zs, a = x[sel], y[sel], z[sel], area[sel]
            w0 = ((xs[:, 1] - px) * (ys[:, 2] - py) - (xs[:, 2] - px) * (ys[:, 1] - py)) / a
            w1 = ((xs[:, 2] - px) * (ys[:, 0] - py) - (xs[:, 0] - px) * (ys[:, 2] - py)) / a
            w2 = 1.0 - w0 - w1
            inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
            zc = (w0 * zs[:, 0] + w1 * zs[:, 1] + w2 * zs[:, 2])[inside]
            k = np.clip(np.ceil(zc - 0.5).astype(int), 0, N)    # first cell whose centre lies above the crossing
            np.add.at(toggles, (ix[sel][inside], iy[sel][inside], k), 1)
    crossings = toggles.sum(axis=2)
    bad = int((crossings % 2).sum())
    solid = (np.cumsum(toggles, axis=2)[:, :, :N] % 2).astype(bool)
    if bad:   # a leaky column would paint a streak to the top face: clear those columns' parity tails
        bx, by = np.nonzero(crossings % 2)
        for i, j in zip(bx, by):
            last = np.nonzero(toggles[i, j])[0].max()
            solid[i, j, last:] = False
    return solid, bad, int(ok.sum())


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    extent = float(sys.argv[2]) if len(sys.argv) > 2 else 0.375 * N
    tri = read_binary_stl(STL)
    solid, bad, ntri = voxelise(tri, N, extent)
    occ = (~solid).astype(np.uint8)                                   # 1 = fluid, 0 = solid
    out = os.path.join(HERE, f"dragon_voxels_{N}.npz")
    np.savez_compressed(out, packed=np.packbits(occ.reshape(-1)), shape=np.array(occ.shape),
                        triangles=np.int64(ntri), solid_cells=np.int64(solid.sum()))
    print(f"{out}: {ntri} triangles, {int(solid.sum())} solid cells ({solid.mean() * 100:.2f} %), "
          f"{bad} leaky columns, {os.path.getsize(out)} bytes")
    idx = np.nonzero(solid)
    print("solid bounding box:", [(int(a.min()), int(a.max())) for a in idx])


if __name__ == "__main__":
    main()
