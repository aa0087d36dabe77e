#!/usr/bin/env python
"""Fixtures for the STL -> porosity row: the triangles of the reference's two sample surfaces
(/root/reference/tools/stl2poro/stl_files/{sphere,dragon}.stl) as float32 arrays -> tests/golden/stl_meshes.npz.
The reference's tool itself cannot be run to produce golden outputs (it needs vtk, absent from the image):
the row is pinned by the analytic distance to the sphere and by ray-casting parity instead (tests/test_stl2poro.py).
Run where /root/reference exists."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pixelflow_b200.stl2poro import read_stl_file  # noqa: E402

SRC = "/root/reference/tools/stl2poro/stl_files"
np.savez_compressed(os.path.join(HERE, "stl_meshes.npz"),
                    sphere=read_stl_file(os.path.join(SRC, "sphere.stl")),
                    dragon=read_stl_file(os.path.join(SRC, "dragon.stl")))
print(os.path.getsize(os.path.join(HERE, "stl_meshes.npz")))
