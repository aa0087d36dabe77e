"""STL surface -> porosity field on the GPU: host-side mirror of the reference's tools/stl2poro/stl2poro.py.

The reference reads a binary STL with VTK, lays a uniform grid over the (scaled) bounding box, asks
vtkImplicitPolyDataDistance for the signed distance at every cell centre -- one Python call per point -- and turns
it into a porosity with 0.5*tanh(d/(thickness*pitch)) + 0.5 (stl2poro.py:71-84, :87-97, :153-206).  VTK is not
needed here: the file is read with numpy, the signed distance comes from libpixelflow_gpu.so
(`pf_stl_signed_distance`, csrc/pf_stl.cu: exact point-triangle distance over all triangles, pseudo-normal sign), and
everything around it keeps the reference's function names, arguments and arithmetic (bounds scaling incl. the
`* 2` of the three-axis variant, pitch from one axis, `ceil` cell counts, `.6E` CSV rows written by csv.writer).
No CPU fallback.

What VTK does implicitly and this module does explicitly: vtkPoints stores float32, so the grid corners
ix*pitch + min and the cell centres (the mean of two corners per axis) are rounded to float32 before the distance is
evaluated; vtkSTLReader merges coincident vertices (done inside the library, by exact equality).
"""
from __future__ import annotations

import csv
import ctypes as C
import math
import struct

import numpy as np

from .api import PixelFlowError, load_library


def read_stl_file(file_path: str) -> np.ndarray:
    """the triangles of an STL file as float32 [ntri][3][3] (vertex, coordinate); binary or ASCII"""
    raw = open(file_path, "rb").read()
    if len(raw) >= 84:
        n = struct.unpack_from("<I", raw, 80)[0]
        if 84 + 50 * n == len(raw):
            rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n,
                                offset=84)
            return np.ascontiguousarray(rec["v"], dtype=np.float32)
    verts = [ln.split()[1:4] for ln in raw.decode("ascii", "replace").splitlines() if ln.strip().startswith("vertex")]
    if not verts or len(verts) % 3:
        raise ValueError(f"{file_path}: neither a binary nor an ASCII STL")
    return np.asarray(verts, dtype=np.float32).reshape(-1, 3, 3)


def get_bounds(triangles: np.ndarray):
    """vtkPolyData.GetBounds(): (xmin, xmax, ymin, ymax, zmin, zmax) as Python floats"""
    v = triangles.reshape(-1, 3).astype(np.float64)
    lo, hi = v.min(axis=0), v.max(axis=0)
    return (float(lo[0]), float(hi[0]), float(lo[1]), float(hi[1]), float(lo[2]), float(hi[2]))


def ratio_margin_to_bounds(bounds, factor):
    return [bound * factor for bound in bounds]                                  # stl2poro.py:25-26


def ratio_margin_to_bounds_for_three_axis(bounds, bounds_factor):
    new_bounds = np.zeros_like(bounds)                                           # stl2poro.py:29-36 (sic: * 2)
    for i in range(6):
        new_bounds[i] = bounds[i] * bounds_factor[i] * 2
    return new_bounds


def calculate_pitch_and_mins(bounds, grid, axis):
    pitch = float((bounds[2 * axis + 1] - bounds[2 * axis]) / grid)              # stl2poro.py:46-50
    mesh_pitch = [pitch] * 3
    mins = [bound - pitch / 2 for bound in bounds[::2]]
    return pitch, mesh_pitch, mins


def cell_centers(cell_dims, mesh_pitch, mins) -> np.ndarray:
    """The points the reference evaluates the distance at, [nz][ny][nx][3] float64 (cell order of the structured
    grid: ix fastest).  create_mesh_grid_points (stl2poro.py:53-62) puts the corners ix*pitch + min into a vtkPoints
    (float32); vtkCellCenters takes the mean of the eight corners of a cell (exact in double) and stores float32."""
    axes = []
    for d in range(3):
        corners = np.array([ix * mesh_pitch[d] + mins[d] for ix in range(cell_dims[d] + 1)], dtype=np.float64)
        corners = corners.astype(np.float32).astype(np.float64)
        axes.append(((corners[:-1] + corners[1:]) / 2).astype(np.float32).astype(np.float64))
    pts = np.empty((cell_dims[2], cell_dims[1], cell_dims[0], 3))
    pts[..., 0] = axes[0][None, None, :]
    pts[..., 1] = axes[1][None, :, None]
    pts[..., 2] = axes[2][:, None, None]
    return pts


def calculate_sdf(triangles: np.ndarray, center_points: np.ndarray, device: int = -1) -> np.ndarray:
    """signed distance (negative inside) of every point to the mesh, on the GPU: stl2poro.py:71-84"""
    tri = np.ascontiguousarray(triangles, dtype=np.float32).reshape(-1, 9)
    pts = np.ascontiguousarray(center_points, dtype=np.float64).reshape(-1, 3)
    out = np.empty(len(pts))
    L = load_library()
    L.pf_stl_signed_distance.argtypes = [C.POINTER(C.c_float), C.c_longlong, C.POINTER(C.c_double), C.c_longlong,
                                         C.POINTER(C.c_double), C.c_int]
    rc = L.pf_stl_signed_distance(tri.ctypes.data_as(C.POINTER(C.c_float)), len(tri),
                                  pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts),
                                  out.ctypes.data_as(C.POINTER(C.c_double)), int(device))
    if rc:
        raise PixelFlowError(L.pf_last_error(None).decode())
    return out.reshape(center_points.shape[:-1])


_tanh = np.frompyfunc(math.tanh, 1, 1)      # the reference calls math.tanh (libm); np.tanh may differ in the last bit


def _process(triangles, bounds, grid, axis, thickness, device):
    pitch, mesh_pitch, mins = calculate_pitch_and_mins(bounds, grid, axis)
    cell_dims = [math.ceil((bounds[i * 2 + 1] - bounds[i * 2]) / mesh_pitch[i // 2]) for i in range(3)]   # :110, :169
    sdf = calculate_sdf(triangles, cell_centers(cell_dims, mesh_pitch, mins), device)       # [nz][ny][nx]
    X = sdf / (thickness * pitch)                                                           # :134, :202
    poro = 0.5 * _tanh(X).astype(np.float64) + 0.5
    return np.ascontiguousarray(poro.transpose(2, 1, 0))                                    # dist_3d_array[ix, iy, iz]


def process_stl_file(file_path, factor, grid, axis, thickness, device: int = -1):
    """stl2poro.py:99-140: porosity array [nx][ny][nz] of the STL with its bounds scaled by `factor`"""
    tri = read_stl_file(file_path)
    return _process(tri, ratio_margin_to_bounds(get_bounds(tri), factor), grid, axis, thickness, device)


def process_stl_file_three_axis(file_path, bounds_factor, grid, axis, thickness, device: int = -1):
    """stl2poro.py:143-206 (what its main() calls): six per-face factors"""
    tri = read_stl_file(file_path)
    bounds = ratio_margin_to_bounds_for_three_axis(np.array(get_bounds(tri)), bounds_factor)
    return _process(tri, bounds, grid, axis, thickness, device)


def save_3d_array_to_csv(csv_file_path, cell_dims, dist_3d_array):
    """stl2poro.py:85-96, byte for byte: csv.writer rows (CRLF), header = the dimensions, then `ix, iy, iz, %.6E` with
    ix fastest -- the porosity CSV lib/grid.f90:281-294 reads"""
    with open(csv_file_path, mode="w", newline="") as file:
        writer = csv.writer(file)
        writer.writerow(cell_dims)
        nx = dist_3d_array.shape[0]
        for iz in range(dist_3d_array.shape[2]):
            for iy in range(dist_3d_array.shape[1]):
                col = dist_3d_array[:, iy, iz]
                file.write("".join(f"{ix + 1},{iy + 1},{iz + 1},{format(col[ix], '.6E')}\r\n" for ix in range(nx)))
