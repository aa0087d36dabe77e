"""Regenerates tests/golden/decks/*.npz from the reference's shipped input decks.

Run in the build container only (reads /root/reference/test/*.zip, which does not exist on the
GPU box):   python tests/golden/make_decks.py

Each .npz holds the raw porosity values of the deck's CSV (before the `max(poro, threshold)` clamp of
lib/grid.f90:50/:289) as a float64 array in [k, j, i] order, plus the controlDict.txt text.  These are
INPUT fixtures (the reference ships no expected outputs).
"""
import io
import os
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/test"
DECKS = {
    "cylinder": ("cylinder-2d.zip", "cylinder-2d/data/porosity_cylinder.csv", "cylinder-2d/config/controlDict.txt"),
    "backstep": ("backstep.zip", "backstep/data/backstep.csv", "backstep/config/controlDict.txt"),
    "room": ("room.zip", "room/data/room.csv", "room/config/controlDict.txt"),
}


def main():
    out = os.path.join(HERE, "decks")
    os.makedirs(out, exist_ok=True)
    for name, (zf, csv, ctl) in DECKS.items():
        z = zipfile.ZipFile(os.path.join(REF, zf))
        text = z.read(csv).decode()
        header, body = text.split("\n", 1)
        m, n, l = (int(t) for t in header.strip().split(","))
        rec = np.loadtxt(io.StringIO(body), delimiter=",", dtype=np.float64)
        assert rec.shape == (m * n * l, 4), rec.shape
        e = np.zeros((l, n, m))
        # the reader honours the explicit indices (lib/grid.f90:288-289)
        ix, iy, iz = (rec[:, c].astype(np.int64) - 1 for c in range(3))
        e[iz, iy, ix] = rec[:, 3]
        np.savez_compressed(os.path.join(out, name + ".npz"), porosity=e, dims=np.array([m, n, l]),
                            controldict=np.array(z.read(ctl).decode()))
        print(name, (m, n, l), "min/max", e.min(), e.max(), os.path.getsize(os.path.join(out, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
