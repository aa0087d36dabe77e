#!/usr/bin/env python
"""Golden vectors for the voxel -> porosity path, produced by the REFERENCE ITSELF.

Runs in the build container only (needs /root/reference): imports the reference's
tools/voxel2poro/voxel2poro.py and tools/voxel2poro/make_bmp_sample.py unmodified (their two
visualisation imports, pyvista and pyevtk, are not installed and are stubbed -- the numerics use
numpy, PIL and scipy.ndimage only) and calls exactly what its main() calls:

    bitmap -> load_bitmap_image -> array_3d (float32) -> create_tanh_kernel(thickness)
           -> scipy.ndimage.convolve(array_3d, kernel, mode='nearest', cval=1.0)     (voxel2poro.py:19-35)

Output: tests/golden/voxel2poro.npz  (inputs, outputs, kernel digests).  The sample case takes ~1 minute
in scipy.

    python tests/golden/make_voxel2poro.py
"""
import hashlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tools/voxel2poro"


def load_reference():
    for name in ("pyvista", "pyevtk", "pyevtk.hl"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pyevtk.hl"].imageToVTK = lambda *a, **k: None
    sys.path.insert(0, REF)
    import make_bmp_sample
    import voxel2poro
    return voxel2poro, make_bmp_sample


def main():
    from PIL import Image
    from scipy.ndimage import convolve
    ref, mk = load_reference()
    out = {}

    # case 1: the reference's own sample (make_bmp_sample.py: 32^3, sphere radius 12; voxel2poro.py: thickness 1.5)
    dim = 32
    data = mk.create_voxel_data(dim, 12)
    arr = np.ones((dim, dim, dim), dtype=np.float32)
    for i in range(dim):
        buf = io.BytesIO()
        Image.fromarray(data[i, :, :], "L").save(buf, format="BMP")   # make_bmp_sample.save_bitmap, in memory
        buf.seek(0)
        arr[:, :, i] = ref.load_bitmap_image(buf)                     # voxel2poro.py:24-26
    kernel = ref.create_tanh_kernel(thickness=1.5)
    out["sphere32_in"] = arr.astype(np.uint8)
    out["sphere32_thickness"] = np.float64(1.5)
    out["sphere32_out"] = convolve(arr, kernel, mode="nearest", cval=1.0)
    out["sphere32_kernel_sha256"] = np.frombuffer(hashlib.sha256(kernel.tobytes()).digest(), dtype=np.uint8)
    out["sphere32_kernel_centre_row"] = kernel[21, 21, :].copy()

    # case 2: non-cubic box, thin interface (ksize = 7), binary voxels
    rng = np.random.default_rng(20240517)
    a2 = (rng.random((24, 20, 28)) < 0.35).astype(np.float32)
    k2 = ref.create_tanh_kernel(thickness=0.5)
    out["box_in"] = a2.astype(np.uint8)
    out["box_thickness"] = np.float64(0.5)
    out["box_out"] = convolve(a2, k2, mode="nearest", cval=1.0)

    # case 3: grey (non-binary float32) voxels, thickness 1.0 (ksize = 14: the kernel is wider than the box)
    a3 = rng.random((12, 9, 10)).astype(np.float32)
    k3 = ref.create_tanh_kernel(thickness=1.0)
    out["grey_in"] = a3
    out["grey_thickness"] = np.float64(1.0)
    out["grey_out"] = convolve(a3, k3, mode="nearest", cval=1.0)

    for k, v in out.items():
        if k.endswith("_out"):
            assert v.dtype == np.float32, (k, v.dtype)
    np.savez_compressed(os.path.join(HERE, "voxel2poro.npz"), **out)
    print({k: (getattr(v, "shape", None), str(getattr(v, "dtype", ""))) for k, v in out.items()})


if __name__ == "__main__":
    main()
