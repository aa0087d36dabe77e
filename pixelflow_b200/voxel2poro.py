"""Voxel model -> porosity field on the GPU: host-side mirror of the reference's tools/voxel2poro/voxel2poro.py.

The reference turns a stack of bitmaps (0 = fluid, 128/255 = solid) into a float32 occupancy array and
smooths it with a normalised `1 - tanh(r/thickness)` kernel through scipy.ndimage.convolve(mode='nearest')
(voxel2poro.py:19-35, :189-197) -- 79,507 taps per voxel at the shipped thickness 1.5, ~45 s for the 32^3
sample on a CPU, hours for the 256^3 grid of BASELINE configs[3].  Here the convolution runs in
libpixelflow_gpu.so (`pf_convolve3d_nearest`, csrc/pf_voxel.cu) and returns the same float32 bits; this
module keeps the reference's function names and argument meaning.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import PixelFlowError, load_library


def create_tanh_kernel(thickness: float = 2.0) -> np.ndarray:
    """voxel2poro.py:189-197: cube of half-width int(14*thickness), 1 - tanh(r/thickness), normalised to sum 1.
    Same numpy operations as the reference (np.tanh and the pairwise np.sum decide the last bit)."""
    half = int(thickness * 14)
    ax = np.linspace(-half, half, int(2 * half + 1))
    gx, gy, gz = np.meshgrid(ax, ax, ax)
    kern = 1 - np.tanh((np.sqrt(gx**2 + gy**2 + gz**2)) / thickness)
    kern /= kern.sum()
    return kern


def convolve_nearest(array_3d: np.ndarray, kernel: np.ndarray, device: int = -1) -> np.ndarray:
    """scipy.ndimage.convolve(array_3d, kernel, mode='nearest') for float32 input and odd-sized float64 kernels"""
    a = np.ascontiguousarray(array_3d, dtype=np.float32)
    w = np.ascontiguousarray(kernel, dtype=np.float64)
    if a.ndim != 3 or w.ndim != 3:
        raise ValueError("expected 3-D arrays")
    L = load_library()
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.pf_convolve3d_nearest.argtypes = [fp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, fp, C.c_int]
    out = np.empty_like(a)
    rc = L.pf_convolve3d_nearest(a.ctypes.data_as(fp), *a.shape, w.ctypes.data_as(dp), *w.shape,
                                 out.ctypes.data_as(fp), int(device))
    if rc:
        raise PixelFlowError("pf_convolve3d_nearest: " + L.pf_last_error(None).decode())
    return out


def voxel2poro(array_3d: np.ndarray, thickness: float = 1.5, device: int = -1) -> np.ndarray:
    """the numerical part of the reference's main() (voxel2poro.py:31-35): occupancy (1 = fluid) -> porosity"""
    return convolve_nearest(array_3d, create_tanh_kernel(thickness=thickness), device)


def load_bitmap_stack(folder: str, dim: int, buff: int = 0) -> np.ndarray:
    """voxel2poro.py:19-26 + load_bitmap_image (:56-65): img_00000.bmp ... -> float32 array, 0 -> 1.0 (fluid),
    128 / 255 -> 0.0 (solid), slice i stored at [:, :, buff + i]"""
    from PIL import Image
    arr = np.ones((dim + 2 * buff,) * 3, dtype=np.float32)
    for i in range(dim):
        img = np.array(Image.open(f"{folder}/img_{i:05d}.bmp"))
        occ = np.where(img == 0, 1, np.where((img == 128) | (img == 255), 0, img)).astype(img.dtype)
        arr[buff:buff + dim, buff:buff + dim, buff + i] = occ
    return arr


def write_porosity(data: np.ndarray, filename: str = "porosity.csv") -> None:
    """voxel2poro.py:200-210: header `m,n,l`, then `i, j, k, value` rows with k outermost, value as %.10f"""
    m, n, l = data.shape
    with open(filename, "w") as f:
        f.write(f"{m},{n},{l}\n")
        for k in range(l):
            plane = data[:, :, k]
            f.write("".join(f"{i + 1}, {j + 1}, {k + 1}, {plane[i, j]:.10f}\n" for j in range(n) for i in range(m)))
