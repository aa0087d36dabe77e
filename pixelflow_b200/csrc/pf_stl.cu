// pf_stl.cu -- signed distance from grid points to a triangle mesh, on the GPU: the numerical core of the reference's
// tools/stl2poro/stl2poro.py (calculate_sdf, :71-84), where it is vtkImplicitPolyDataDistance.FunctionValue called
// point by point from Python (hours for the 256^3 grid of BASELINE configs[3]).
//
// Distance = to the closest point of the mesh (exact point-triangle distance over ALL triangles: brute force, the
// triangles stream through shared memory in tiles, one grid point per thread); sign = that of (p - closest) . N, N
// the pseudo-normal of the feature the closest point lies on -- the face normal inside a triangle, the sum of the two
// face normals on an edge, the angle-weighted sum of the incident face normals at a vertex -- negative inside.
// Vertices are merged by exact coordinate equality (as vtkSTLReader does).  fp64 throughout, fixed evaluation order,
// no FMA contraction: the CPU checker of tests/ returns the same bits.  Host mirror of the tool:
// pixelflow_b200/stl2poro.py.  No CPU fallback.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <utility>
#include <vector>

#include "pf_internal.cuh"

namespace {

struct Tri {            // 34 doubles
  double a[3], ab[3], ac[3];
  double n[3];          // unit face normal
  double en[3][3];      // edge pseudo-normals: AB, BC, CA
  double vn[3][3];      // vertex pseudo-normals: A, B, C
  double c[3], rad;     // a sphere that contains the triangle (centroid, largest vertex distance rounded up)
};

struct Corner { float x[3]; long long idx; };

// merged vertex ids, face normals, pseudo-normals: on the host, once per mesh (libm acos / sqrt)
std::vector<Tri> build_triangles(const float *tri, long long ntri) {
  const long long nc = 3 * ntri;
  std::vector<Corner> keys((size_t)nc);
  for (long long c = 0; c < nc; ++c) {
    for (int d = 0; d < 3; ++d) {
      const float v = tri[3 * c + d];
      keys[c].x[d] = v == 0.0f ? 0.0f : v;   // -0 and +0 are one coordinate
    }
    keys[c].idx = c;
  }
  std::sort(keys.begin(), keys.end(), [](const Corner &p, const Corner &q) {
    const int c = memcmp(p.x, q.x, sizeof p.x);
    return c ? c < 0 : p.idx < q.idx;
  });
  std::vector<long long> vid((size_t)nc);
  for (long long c = 0, first = 0; c < nc; ++c) {
    if (c > 0 && memcmp(keys[c].x, keys[c - 1].x, sizeof keys[c].x) != 0) first = c;
    vid[keys[c].idx] = keys[first].idx;
  }
  struct V3 { double v[3] = {0., 0., 0.}; };
  std::vector<V3> vsum((size_t)nc), fn((size_t)ntri);
  std::vector<char> good((size_t)ntri, 0);
  std::map<std::pair<long long, long long>, V3> esum;   // keyed by the edge's two vertex ids; filled in triangle order
  auto dot = [](const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
  for (long long t = 0; t < ntri; ++t) {
    double p[3][3], ab[3], ac[3], n[3];
    for (int k = 0; k < 3; ++k)
      for (int d = 0; d < 3; ++d) p[k][d] = (double)tri[9 * t + 3 * k + d];
    for (int d = 0; d < 3; ++d) { ab[d] = p[1][d] - p[0][d]; ac[d] = p[2][d] - p[0][d]; }
    n[0] = ab[1] * ac[2] - ab[2] * ac[1];
    n[1] = ab[2] * ac[0] - ab[0] * ac[2];
    n[2] = ab[0] * ac[1] - ab[1] * ac[0];
    const double len = sqrt(dot(n, n));
    if (!(len > 0.0)) continue;   // a degenerate triangle has no surface and no normal
    good[t] = 1;
    for (int d = 0; d < 3; ++d) fn[t].v[d] = n[d] / len;
    for (int k = 0; k < 3; ++k) {
      double e1[3], e2[3];
      for (int d = 0; d < 3; ++d) { e1[d] = p[(k + 1) % 3][d] - p[k][d]; e2[d] = p[(k + 2) % 3][d] - p[k][d]; }
      double c = dot(e1, e2) / (sqrt(dot(e1, e1)) * sqrt(dot(e2, e2)));
      c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
      const double ang = acos(c);
      V3 &vs = vsum[vid[3 * t + k]];
      for (int d = 0; d < 3; ++d) vs.v[d] = vs.v[d] + ang * fn[t].v[d];
      const long long u0 = vid[3 * t + k], u1 = vid[3 * t + (k + 1) % 3];
      V3 &es = esum[std::make_pair(std::min(u0, u1), std::max(u0, u1))];
      for (int d = 0; d < 3; ++d) es.v[d] = es.v[d] + fn[t].v[d];
    }
  }
  std::vector<Tri> out;
  out.reserve((size_t)ntri);
  for (long long t = 0; t < ntri; ++t) {
    if (!good[t]) continue;
    Tri T;
    for (int d = 0; d < 3; ++d) {
      T.a[d] = (double)tri[9 * t + d];
      T.ab[d] = (double)tri[9 * t + 3 + d] - T.a[d];
      T.ac[d] = (double)tri[9 * t + 6 + d] - T.a[d];
      T.n[d] = fn[t].v[d];
    }
    for (int k = 0; k < 3; ++k) {
      const long long u0 = vid[3 * t + k], u1 = vid[3 * t + (k + 1) % 3];
      const V3 &es = esum[std::make_pair(std::min(u0, u1), std::max(u0, u1))];
      const V3 &vs = vsum[u0];
      for (int d = 0; d < 3; ++d) { T.en[k][d] = es.v[d]; T.vn[k][d] = vs.v[d]; }
    }
    double r2 = 0.0;
    for (int d = 0; d < 3; ++d) T.c[d] = T.a[d] + (T.ab[d] + T.ac[d]) / 3.0;
    for (int k = 0; k < 3; ++k) {
      double q2 = 0.0;
      for (int d = 0; d < 3; ++d) {
        const double v = (double)tri[9 * t + 3 * k + d] - T.c[d];
        q2 += v * v;
      }
      r2 = std::max(r2, q2);
    }
    T.rad = sqrt(r2) * (1.0 + 1e-12) + 1e-300;
    out.push_back(T);
  }
  return out;
}

constexpr int STL_THREADS = 256, STL_TILE = 64;

__device__ __forceinline__ double dot3(const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

// closest point of a triangle to p (Ericson, Real-Time Collision Detection 5.1.5): the offset p - closest point, the
// feature it lies on (0 face, 1..3 edge AB BC CA, 4..6 vertex A B C); returns the squared distance
__device__ __forceinline__ double closest(const Tri &T, const double *p, double *off, int &feature) {
  double ap[3], q[3];
  for (int d = 0; d < 3; ++d) ap[d] = p[d] - T.a[d];
  const double d1 = dot3(T.ab, ap), d2 = dot3(T.ac, ap);
  bool done = false;
  if (d1 <= 0.0 && d2 <= 0.0) {
    feature = 4;
    for (int d = 0; d < 3; ++d) q[d] = T.a[d];
    done = true;
  }
  double d3 = 0., d4 = 0., d5 = 0., d6 = 0.;
  if (!done) {
    double bp[3];
    for (int d = 0; d < 3; ++d) bp[d] = ap[d] - T.ab[d];
    d3 = dot3(T.ab, bp);
    d4 = dot3(T.ac, bp);
    if (d3 >= 0.0 && d4 <= d3) {
      feature = 5;
      for (int d = 0; d < 3; ++d) q[d] = T.a[d] + T.ab[d];
      done = true;
    }
  }
  double vc = 0.;
  if (!done) {
    vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
      const double v = d1 / (d1 - d3);
      feature = 1;
      for (int d = 0; d < 3; ++d) q[d] = T.a[d] + v * T.ab[d];
      done = true;
    }
  }
  if (!done) {
    double cp[3];
    for (int d = 0; d < 3; ++d) cp[d] = ap[d] - T.ac[d];
    d5 = dot3(T.ab, cp);
    d6 = dot3(T.ac, cp);
    if (d6 >= 0.0 && d5 <= d6) {
      feature = 6;
      for (int d = 0; d < 3; ++d) q[d] = T.a[d] + T.ac[d];
      done = true;
    }
  }
  double vb = 0.;
  if (!done) {
    vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
      const double w = d2 / (d2 - d6);
      feature = 3;
      for (int d = 0; d < 3; ++d) q[d] = T.a[d] + w * T.ac[d];
      done = true;
    }
  }
  if (!done) {
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
      const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
      feature = 2;
      for (int d = 0; d < 3; ++d) q[d] = (T.a[d] + T.ab[d]) + w * (T.ac[d] - T.ab[d]);
    } else {
      const double den = 1.0 / (va + vb + vc);
      const double v = vb * den, w = vc * den;
      feature = 0;
      for (int d = 0; d < 3; ++d) q[d] = (T.a[d] + T.ab[d] * v) + T.ac[d] * w;
    }
  }
  for (int d = 0; d < 3; ++d) off[d] = p[d] - q[d];
  return dot3(off, off);
}

__global__ void __launch_bounds__(STL_THREADS) stl_distance_kernel(const Tri *__restrict__ tris, long long ntri,
                                                                   const double *__restrict__ pts, long long npts,
                                                                   double *__restrict__ dist) {
  __shared__ Tri tile[STL_TILE];
  const long long i = blockIdx.x * (long long)STL_THREADS + threadIdx.x;
  double p[3] = {0., 0., 0.};
  if (i < npts) { p[0] = pts[3 * i]; p[1] = pts[3 * i + 1]; p[2] = pts[3 * i + 2]; }
  double best = INFINITY, sbest = INFINITY, boff[3] = {0., 0., 0.};
  long long bt = 0;
  int bf = 0;
  for (long long t0 = 0; t0 < ntri; t0 += STL_TILE) {
    const int nt = (int)min((long long)STL_TILE, ntri - t0);
    __syncthreads();
    {   // the tile as plain doubles, coalesced
      const double *src = reinterpret_cast<const double *>(tris + t0);
      double *dst = reinterpret_cast<double *>(tile);
      for (int q = threadIdx.x; q < nt * (int)(sizeof(Tri) / 8); q += STL_THREADS) dst[q] = src[q];
    }
    __syncthreads();
    if (i < npts)
      for (int t = 0; t < nt; ++t) {
        // a triangle whose bounding sphere lies farther away than the best distance so far cannot be closer: its
        // distance is at least |p - c| - rad.  The bound carries a relative margin far above rounding, so no triangle
        // that could win (or tie) is skipped and the result is that of the plain loop, bit for bit.
        const Tri &T = tile[t];
        const double cx = p[0] - T.c[0], cy = p[1] - T.c[1], cz = p[2] - T.c[2];
        const double dc2 = cx * cx + cy * cy + cz * cz;
        const double lim = best + (2.0 * sbest + T.rad) * T.rad;       // (sqrt(best) + rad)^2
        if (dc2 > lim * (1.0 + 1e-9)) continue;
        double off[3];
        int f;
        const double d2 = closest(T, p, off, f);
        if (d2 < best) {
          best = d2; bt = t0 + t; bf = f; boff[0] = off[0]; boff[1] = off[1]; boff[2] = off[2];
          sbest = sqrt(best) * (1.0 + 1e-12);
        }
      }
  }
  if (i >= npts) return;
  const Tri &B = tris[bt];
  const double *N = bf == 0 ? B.n : (bf <= 3 ? B.en[bf - 1] : B.vn[bf - 4]);
  const double s = dot3(boff, N), d = sqrt(best);
  dist[i] = s < 0.0 ? -d : d;
}

}  // namespace

extern "C" int pf_stl_signed_distance(const float *tri, long long ntri, const double *points, long long npoints,
                                      double *dist, int device) {
  Tri *d_tri = nullptr;
  double *d_pts = nullptr, *d_out = nullptr;
  int rc = 0;
  try {
    if (!tri || !points || !dist) throw std::string("pf_stl_signed_distance: null array");
    if (ntri < 1 || npoints < 0) throw std::string("pf_stl_signed_distance: empty mesh");
    const std::vector<Tri> T = build_triangles(tri, ntri);
    if (T.empty()) throw std::string("pf_stl_signed_distance: every triangle is degenerate");
    if (device >= 0) PF_CUDA_OK(cudaSetDevice(device));
    PF_CUDA_OK(cudaMalloc(&d_tri, T.size() * sizeof(Tri)));
    PF_CUDA_OK(cudaMemcpy(d_tri, T.data(), T.size() * sizeof(Tri), cudaMemcpyHostToDevice));
    // the points in slices: the grid of a 256^3 job is 400 MB of coordinates
    const long long slice = 1ll << 24;
    PF_CUDA_OK(cudaMalloc(&d_pts, (size_t)std::min(slice, std::max(npoints, 1ll)) * 3 * sizeof(double)));
    PF_CUDA_OK(cudaMalloc(&d_out, (size_t)std::min(slice, std::max(npoints, 1ll)) * sizeof(double)));
    for (long long p0 = 0; p0 < npoints; p0 += slice) {
      const long long np = std::min(slice, npoints - p0);
      PF_CUDA_OK(cudaMemcpy(d_pts, points + 3 * p0, (size_t)np * 3 * sizeof(double), cudaMemcpyHostToDevice));
      stl_distance_kernel<<<(unsigned)((np + STL_THREADS - 1) / STL_THREADS), STL_THREADS>>>(d_tri, (long long)T.size(), d_pts,
                                                                                           np, d_out);
      pf_count_launch();
      PF_CUDA_OK(cudaGetLastError());
      PF_CUDA_OK(cudaMemcpy(dist + p0, d_out, (size_t)np * sizeof(double), cudaMemcpyDeviceToHost));
    }
  } catch (const std::string &e) {
    pf_set_global_error(e);
    rc = 1;
  }
  cudaFree(d_tri);
  cudaFree(d_pts);
  cudaFree(d_out);
  return rc;
}
