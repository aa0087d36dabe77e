/* stl_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the signed distance field behind the reference's
 * tools/stl2poro/stl2poro.py (calculate_sdf, :71-84: vtkImplicitPolyDataDistance.FunctionValue at every cell centre).
 *
 * The reference delegates the geometry to VTK (`import vtk`, stl2poro.py:2; no version pinned anywhere in the
 * repository), which is absent from this image -- PARITY UNPINNED for this row: no output of the reference's tool can
 * be produced here.  What is restated is VTK's published algorithm: the distance from the point to the closest point
 * of the triangle mesh (exact point-triangle distance, over all triangles), signed by the normal at that closest
 * point -- the face normal inside a triangle, the sum of the two face normals on an edge, the angle-weighted sum of
 * the incident face normals at a vertex (Baerentzen & Aanaes' pseudo-normals; VTK averages the incident normals
 * without the angle weights, which gives the same sign wherever VTK's is right) -- negative inside.  Vertices are
 * merged by exact coordinate equality, as vtkSTLReader's point merging does.  The restatement is pinned by what CAN be
 * checked: the analytic distance to the sphere of stl_files/sphere.stl (tests/test_stl2poro.py) and inside/outside
 * parity against ray casting.  Only tests/ may call this.
 *
 * Brute force, O(points x triangles), OpenMP over the points; evaluation order fixed (no FMA: -ffp-contract=off), so
 * the CUDA kernel (pixelflow_b200/csrc/pf_stl.cu) can be compared bit for bit.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PUB __attribute__((visibility("default")))

typedef struct { double a[3], ab[3], ac[3], n[3], en[3][3], vn[3][3]; } tri_t;   /* en: AB, BC, CA; vn: A, B, C */

static double dot3(const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }

/* ---- vertex merging and pseudo-normals ------------------------------------------------------------------------- */
typedef struct { float x[3]; long long idx; } vkey_t;
static int vkey_cmp(const void *p, const void *q) {
  const vkey_t *a = p, *b = q;
  int c = memcmp(a->x, b->x, sizeof a->x);
  if (c) return c;
  return a->idx < b->idx ? -1 : (a->idx > b->idx);
}
typedef struct { long long u, v, t; int e; } ekey_t;
static int ekey_cmp(const void *p, const void *q) {
  const ekey_t *a = p, *b = q;
  if (a->u != b->u) return a->u < b->u ? -1 : 1;
  if (a->v != b->v) return a->v < b->v ? -1 : 1;
  return a->t < b->t ? -1 : (a->t > b->t);
}

/* tri[ntri][3][3] float32 as stored in a binary STL.  Returns the number of non-degenerate triangles written to out. */
static long long build_triangles(const float *tri, long long ntri, tri_t *out) {
  /* vertex ids: corners sorted by their 12 bytes; the id of a corner is the smallest corner index of its group */
  const long long nc = 3 * ntri;
  vkey_t *keys = malloc((size_t)nc * sizeof *keys);
  long long *vid = malloc((size_t)nc * sizeof *vid);
  for (long long c = 0; c < nc; c++) { memcpy(keys[c].x, tri + 3 * c, sizeof keys[c].x); keys[c].idx = c; }
  for (long long c = 0; c < nc; c++)       /* -0.0 and +0.0 are the same coordinate */
    for (int d = 0; d < 3; d++) if (keys[c].x[d] == 0.0f) keys[c].x[d] = 0.0f;
  qsort(keys, (size_t)nc, sizeof *keys, vkey_cmp);
  for (long long c = 0, first = 0; c < nc; c++) {
    if (c > 0 && memcmp(keys[c].x, keys[c - 1].x, sizeof keys[c].x) != 0) first = c;
    vid[keys[c].idx] = keys[first].idx;
  }
  free(keys);
  double (*vn)[3] = calloc((size_t)nc, sizeof *vn);      /* indexed by vertex id (a corner index) */
  double (*fn)[3] = calloc((size_t)ntri, sizeof *fn);
  char *good = calloc((size_t)ntri, 1);
  ekey_t *edges = malloc((size_t)nc * sizeof *edges);
  long long ne = 0;
  for (long long t = 0; t < ntri; t++) {
    double p[3][3];
    for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) p[k][d] = (double)tri[9 * t + 3 * k + d];
    double ab[3], ac[3], n[3];
    for (int d = 0; d < 3; d++) { ab[d] = p[1][d] - p[0][d]; ac[d] = p[2][d] - p[0][d]; }
    n[0] = ab[1] * ac[2] - ab[2] * ac[1];
    n[1] = ab[2] * ac[0] - ab[0] * ac[2];
    n[2] = ab[0] * ac[1] - ab[1] * ac[0];
    const double len = sqrt(dot3(n, n));
    if (!(len > 0.0)) continue;                            /* degenerate: no surface, no normal */
    good[t] = 1;
    for (int d = 0; d < 3; d++) fn[t][d] = n[d] / len;
    for (int k = 0; k < 3; k++) {                          /* angle at corner k, between the edges leaving it */
      double e1[3], e2[3];
      for (int d = 0; d < 3; d++) { e1[d] = p[(k + 1) % 3][d] - p[k][d]; e2[d] = p[(k + 2) % 3][d] - p[k][d]; }
      double c = dot3(e1, e2) / (sqrt(dot3(e1, e1)) * sqrt(dot3(e2, e2)));
      if (c > 1.0) c = 1.0;
      if (c < -1.0) c = -1.0;
      const double ang = acos(c);
      const long long v = vid[3 * t + k];
      for (int d = 0; d < 3; d++) vn[v][d] = vn[v][d] + ang * fn[t][d];
      const long long u0 = vid[3 * t + k], u1 = vid[3 * t + (k + 1) % 3];   /* edge k: corner k -> corner k+1 */
      edges[ne].u = u0 < u1 ? u0 : u1; edges[ne].v = u0 < u1 ? u1 : u0; edges[ne].t = t; edges[ne].e = k; ne++;
    }
  }
  /* edge pseudo-normals: the face normals of the triangles sharing the edge, summed in triangle order */
  double (*en)[3][3] = calloc((size_t)ntri, sizeof *en);
  qsort(edges, (size_t)ne, sizeof *edges, ekey_cmp);
  for (long long i = 0; i < ne;) {
    long long j = i;
    double s[3] = {0, 0, 0};
    while (j < ne && edges[j].u == edges[i].u && edges[j].v == edges[i].v) {
      for (int d = 0; d < 3; d++) s[d] = s[d] + fn[edges[j].t][d];
      j++;
    }
    for (long long q = i; q < j; q++) for (int d = 0; d < 3; d++) en[edges[q].t][edges[q].e][d] = s[d];
    i = j;
  }
  long long m = 0;
  for (long long t = 0; t < ntri; t++) {
    if (!good[t]) continue;
    tri_t *o = &out[m++];
    for (int d = 0; d < 3; d++) {
      o->a[d] = (double)tri[9 * t + d];
      o->ab[d] = (double)tri[9 * t + 3 + d] - o->a[d];
      o->ac[d] = (double)tri[9 * t + 6 + d] - o->a[d];
      o->n[d] = fn[t][d];
      for (int k = 0; k < 3; k++) { o->en[k][d] = en[t][k][d]; o->vn[k][d] = vn[vid[3 * t + k]][d]; }
    }
  }
  free(vid); free(vn); free(fn); free(good); free(edges); free(en);
  return m;
}

/* closest point of triangle T to p (Ericson, Real-Time Collision Detection 5.1.5): squared distance, the offset
 * p - closest point, and the feature the closest point lies on (0 face, 1..3 edge AB BC CA, 4..6 vertex A B C) */
static double closest(const tri_t *T, const double *p, double *off, int *feature) {
  double ap[3], q[3];
  for (int d = 0; d < 3; d++) ap[d] = p[d] - T->a[d];
  const double d1 = dot3(T->ab, ap), d2 = dot3(T->ac, ap);
  int f;
  if (d1 <= 0.0 && d2 <= 0.0) { f = 4; for (int d = 0; d < 3; d++) q[d] = T->a[d]; goto done; }
  double bp[3];
  for (int d = 0; d < 3; d++) bp[d] = ap[d] - T->ab[d];
  const double d3 = dot3(T->ab, bp), d4 = dot3(T->ac, bp);
  if (d3 >= 0.0 && d4 <= d3) { f = 5; for (int d = 0; d < 3; d++) q[d] = T->a[d] + T->ab[d]; goto done; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
    const double v = d1 / (d1 - d3);
    f = 1; for (int d = 0; d < 3; d++) q[d] = T->a[d] + v * T->ab[d]; goto done;
  }
  double cp[3];
  for (int d = 0; d < 3; d++) cp[d] = ap[d] - T->ac[d];
  const double d5 = dot3(T->ab, cp), d6 = dot3(T->ac, cp);
  if (d6 >= 0.0 && d5 <= d6) { f = 6; for (int d = 0; d < 3; d++) q[d] = T->a[d] + T->ac[d]; goto done; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    const double w = d2 / (d2 - d6);
    f = 3; for (int d = 0; d < 3; d++) q[d] = T->a[d] + w * T->ac[d]; goto done;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    f = 2; for (int d = 0; d < 3; d++) q[d] = (T->a[d] + T->ab[d]) + w * (T->ac[d] - T->ab[d]); goto done;
  }
  {
    const double den = 1.0 / (va + vb + vc);
    const double v = vb * den, w = vc * den;
    f = 0; for (int d = 0; d < 3; d++) q[d] = (T->a[d] + T->ab[d] * v) + T->ac[d] * w;
  }
done:
  for (int d = 0; d < 3; d++) off[d] = p[d] - q[d];
  *feature = f;
  return dot3(off, off);
}

/* signed distance at npts points (pts[npts][3], double) to the mesh tri[ntri][3][3] (float32, binary-STL layout) */
PUB int pfo_stl_signed_distance(const float *tri, long long ntri, const double *pts, long long npts, double *dist) {
  if (ntri < 1 || npts < 0) return 1;
  tri_t *T = malloc((size_t)ntri * sizeof *T);
  const long long m = build_triangles(tri, ntri, T);
  if (m < 1) { free(T); return 1; }
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < npts; i++) {
    double best = INFINITY, boff[3] = {0, 0, 0};
    long long bt = 0;
    int bf = 0;
    for (long long t = 0; t < m; t++) {
      double off[3];
      int f;
      const double d2 = closest(&T[t], pts + 3 * i, off, &f);
      if (d2 < best) { best = d2; bt = t; bf = f; boff[0] = off[0]; boff[1] = off[1]; boff[2] = off[2]; }
    }
    const double *N = bf == 0 ? T[bt].n : (bf <= 3 ? T[bt].en[bf - 1] : T[bt].vn[bf - 4]);
    const double s = dot3(boff, N), d = sqrt(best);
    dist[i] = s < 0.0 ? -d : d;
  }
  free(T);
  return 0;
}
