// pf_kernels.cu -- every phase of the PixelFlow time step except the SOR sweeps:
// divergence, momentum predictor, Poisson coefficients / right-hand side (with the boundary-matrix
// fold), projection, boundary conditions, initial conditions, natural<->checkerboard conversion.
//
// All kernels are fp64, one thread per cell along x (coalesced rows), compiled with -fmad=false so
// that every expression is evaluated exactly as typed: statement order and association follow the
// reference source line by line (citations at each device function).  These phases move
// 208 B/cell/step against 88 B/cell per SOR iteration, i.e. ~2 % of a step at
// iter_max=100; they are written for exactness first and rely on L1/L2 for stencil reuse.
#include <algorithm>

#include "pf_internal.cuh"

// per host thread: a solver is driven by one thread at a time (include/pixelflow_gpu.h), so two solvers on two threads
// never share a counter
static thread_local long long g_launches = 0;
long long pf_launch_count() { return g_launches; }
void pf_launch_count_reset() { g_launches = 0; }
void pf_count_launch() { ++g_launches; }
#define LAUNCHED() (++g_launches)

int pf_sm_count() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  PF_CUDA_OK(cudaGetDevice(&dev));
  if (dev != dev_cached) {
    PF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    dev_cached = dev;
  }
  return sms;
}

// (the register-prefetch fused kernels; the TMA kernel has its own schedule, pf_tma_schedule)
// Cost model fitted to B200 measurements (tools/chunk_sweep.py, profiles/r01_v5_chunk_sweep.txt): a block takes
// (cz + 2) z-steps, its start-up cost is below one step, and blocks are list-scheduled on `slots` resident blocks, so
// the makespan is the smaller of whole waves and (average load + a quarter block of tail), never less than one block.
int pf_chunk_planes(int lz, long long tiles, int slots) {
  int best = lz;
  double best_cost = 1e30;
  for (int cz = lz; cz >= 8; --cz) {
    const int nz = (lz + cz - 1) / cz;
    if (nz > 1 && (lz + nz - 1) / nz != cz) continue;   // only the even splits
    const double steps = cz + 2;
    const long long blocks = tiles * nz;
    const double waves = (double)((blocks + slots - 1) / slots);
    const double cost = std::max(steps, std::min(waves * steps, (double)blocks * steps / slots + 0.25 * steps));
    if (cost < best_cost) { best_cost = cost; best = cz; }
  }
  return best;
}

namespace {

constexpr double SMALL = 1.e-6;  // ibm_3d_uniform_omp_cpu.f90:169
constexpr double ALPHA = 32.0;   // ibm_3d_uniform_omp_cpu.f90:170
constexpr int BX = 64, BY = 4;

inline dim3 cell_grid(const Geo &g, int nx, int ny, int nz) {
  return dim3((nx + BX - 1) / BX, (ny + BY - 1) / BY, nz);
}

#define CELL_IJK(i0, j0, k0)                                       \
  const int i = blockIdx.x * BX + threadIdx.x + (i0);              \
  const int j = blockIdx.y * BY + threadIdx.y + (j0);              \
  const int k = (int)blockIdx.z + (k0);

// ----------------------------------------------------------------------------------------------
// divergence   3D: ibm_3d_uniform_omp_cpu.f90:185-195   2D: ibm_2d_uniform_omp_cpu.f90:172-178
// ----------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(BX *BY) divergence_kernel(Geo g, Phys ph, Fields f) {
  CELL_IJK(1, 1, g.kin0)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  if (DIM == 3) {
    f.div[c] = (f.uo[c + 1] - f.uo[c - 1]) / ph.ix * 0.5 +
               (f.vo[c + g.NX] - f.vo[c - g.NX]) / ph.iy * 0.5 +
               (f.wo[c + g.plane] - f.wo[c - g.plane]) / ph.iz * 0.5;
  } else {
    // second term divides by dx, sic (ibm_2d_uniform_omp_cpu.f90:176)
    f.div[c] = (f.uo[c + 1] - f.uo[c - 1]) / ph.ix * .5 + (f.vo[c + g.NX] - f.vo[c - g.NX]) / ph.ix * .5;
  }
}

// periodic y halo of div for i=1..m (3D :206-213, 2D :188-192); x-face halos stay 0 forever
// (the array is zero-initialised and nothing else writes them), like the air-condition halos.
__global__ void div_halo_y_kernel(Geo g, double *div) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int k = (int)blockIdx.y + g.kin0;
  if (i > g.m) return;
  div[nat_idx(g, i, 0, k)] = div[nat_idx(g, i, g.n, k)];
  div[nat_idx(g, i, g.n + 1, k)] = div[nat_idx(g, i, 1, k)];
}

// a(i,j,kd) = a(i,j,ks) for i=1..m, j=1..n   (periodic z of div :215-222, single rank)
__global__ void plane_copy_interior_kernel(Geo g, double *a, int kd, int ks) {
  const int i = blockIdx.x * BX + threadIdx.x + 1;
  const int j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > g.m || j > g.n) return;
  a[nat_idx(g, i, j, kd)] = a[nat_idx(g, i, j, ks)];
}

// a(i,j,kd) = a(i,j,ks) for i=0..m+1, j=0..n+1   (periodic z of boundary :735-748)
__global__ void plane_copy_full_kernel(Geo g, double *a, int kd, int ks) {
  const int i = blockIdx.x * BX + threadIdx.x;
  const int j = blockIdx.y * BY + threadIdx.y;
  if (i > g.m + 1 || j > g.n + 1) return;
  a[nat_idx(g, i, j, kd)] = a[nat_idx(g, i, j, ks)];
}

// u_old <- u by exchanging the two buffers (pf_api.cu: do_copy_old): the predictor rewrites every interior cell of
// the new `u`, so only its halo shell -- i in {0,m+1}, j in {0,n+1}, k in {0,lz+1} -- must carry the previous
// step's values.  One block per (j,k) row: whole row on the shell faces, two cells elsewhere.
__global__ void shell_copy_kernel(Geo g, const double *s0, const double *s1, const double *s2, double *d0, double *d1,
                                  double *d2) {
  const int j = blockIdx.y, k = blockIdx.z;
  const bool face = j == 0 || j == g.n + 1 || (g.dim == 3 && (k == 0 || k == g.lz + 1));
  const long long row = nat_idx(g, 0, j, k);
  if (face) {
    for (int i = threadIdx.x; i <= g.m + 1; i += blockDim.x) {
      d0[row + i] = s0[row + i];
      d1[row + i] = s1[row + i];
      if (s2) d2[row + i] = s2[row + i];
    }
  } else if (threadIdx.x < 2) {
    const int i = threadIdx.x ? g.m + 1 : 0;
    d0[row + i] = s0[row + i];
    d1[row + i] = s1[row + i];
    if (s2) d2[row + i] = s2[row + i];
  }
}

// ----------------------------------------------------------------------------------------------
// momentum predictor.  3D: ibm_3d_uniform_omp_cpu.f90:228-381 (identical in air-condition
// :275-428); the three component sweeps are fused into one pass -- each is still the reference's
// chain of in-place statements, evaluated in a register.
// ----------------------------------------------------------------------------------------------
// The arithmetic of one cell, templated on the divisor type: FastDiv = exact reciprocal division without
// per-division range checks (used when every stencil value is 0 or of moderate magnitude, checked once
// per cell), ExactDiv = IEEE division (the rare fallback).  Both give the same bits.
struct FastDiv { double d, r; };
struct ExactDiv { double d; };
__device__ __forceinline__ double operator/(double a, const FastDiv &b) {
  double q = a * b.r;
  double e = __fma_rn(-q, b.d, a);
  q = __fma_rn(e, b.r, q);
  e = __fma_rn(-q, b.d, a);
  return __fma_rn(e, b.r, q);
}
__device__ __forceinline__ double operator/(double a, const ExactDiv &b) { return a / b.d; }
__device__ __forceinline__ FastDiv as_div(const Inv &v, FastDiv *) { return FastDiv{v.d, v.r}; }
__device__ __forceinline__ ExactDiv as_div(const Inv &v, ExactDiv *) { return ExactDiv{v.d}; }
// 0, or 1e-60 < |x| < 1e60 (biased exponent 0x337 .. 0x4C6): products and differences of a handful of
// such values stay far inside the range where the reciprocal sequence is exact
__device__ __forceinline__ bool moderate(double x) {
  const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
  return (hi - 0x33700000u) < (0x4C700000u - 0x33700000u) || (hi | (unsigned)__double2loint(x)) == 0u;
}

struct Stencil3 {
  double uc, ue, uw, un, us, ut, ub, vc, ve, vw, vn, vs, vt, vb, wc, we, ww, wn, ws, wt, wb;
  double ec, ee, ew, en, es, et, eb, dc, de, dw, dn, ds, dtp, db;
};

template <class D>
__device__ __forceinline__ void predictor3_cell(const Phys &ph, const Stencil3 &q, double &ru, double &rv, double &rw) {
  const D dx = as_div(ph.ix, (D *)nullptr), dy = as_div(ph.iy, (D *)nullptr), dz = as_div(ph.iz, (D *)nullptr);
  const D tx2 = as_div(ph.itx2, (D *)nullptr), ty2 = as_div(ph.ity2, (D *)nullptr), tz2 = as_div(ph.itz2, (D *)nullptr);
  const double dt = ph.dt, xnue = ph.xnue, xlambda = ph.xlambda;
  const int nonslip = ph.nonslip;
  const double uc = q.uc, ue = q.ue, uw = q.uw, un = q.un, us = q.us, ut = q.ut, ub = q.ub;
  const double vc = q.vc, ve = q.ve, vw = q.vw, vn = q.vn, vs = q.vs, vt = q.vt, vb = q.vb;
  const double wc = q.wc, we = q.we, ww = q.ww, wn = q.wn, ws = q.ws, wt = q.wt, wb = q.wb;
  const double ec = q.ec, ee = q.ee, ew = q.ew, en = q.en, es = q.es, et = q.et, eb = q.eb;
  const double dc = q.dc, de = q.de, dw = q.dw, dn = q.dn, ds = q.ds, dtp = q.dtp, db = q.db;
  double r;

  // ---- u :233-271
  r = uc - dt * uc * (ue - uw) / dx * 0.5;
  r = r - dt * vc * (un - us) / dy * 0.5;
  r = r - dt * wc * (ut - ub) / dz * 0.5;
  r = r + dt * xnue * (ue - 2. * uc + uw) / dx / dx;
  r = r + dt * xnue * (un - 2. * uc + us) / dy / dy;
  r = r + dt * xnue * (ut - 2. * uc + ub) / dz / dz;
  r = r + dt * (xnue + xlambda) * (de - dw) / dx * 0.5;
  r = r + dt * (((ue - uw) / dx * 0.5 + (ue - uw) / dx * 0.5) * xnue * (ee - ew) / dx * 0.5 +
                ((un - us) / dy * 0.5 + (ve - vw) / dx * 0.5) * xnue * (en - es) / dy * 0.5 +
                ((ut - ub) / dz * 0.5 + (we - ww) / dx * 0.5) * xnue * (et - eb) / dz * 0.5 +
                dc * (ee - ew) / dx * 0.5 * xlambda) /
              ec;
  if (nonslip)
    r = r - dt * xnue * uc / tx2 * ALPHA * ec * (1. - ec) *
                (1. - ec);
  ru = r;
  // ---- v :286-323
  r = vc - dt * uc * (ve - vw) / dx * 0.5;
  r = r - dt * vc * (vn - vs) / dy * 0.5;
  r = r - dt * wc * (vt - vb) / dz * 0.5;
  r = r + dt * xnue * (ve - 2. * vc + vw) / dx / dx;
  r = r + dt * xnue * (vn - 2. * vc + vs) / dy / dy;
  r = r + dt * xnue * (vt - 2. * vc + vb) / dz / dz;
  r = r + dt * (xnue + xlambda) * (dn - ds) / dy * 0.5;
  r = r + dt * (((ve - vw) / dx * 0.5 + (un - us) / dy * 0.5) * xnue * (ee - ew) / dx * 0.5 +
                ((vn - vs) / dy * .5 + (vn - vs) / dy * 0.5) * xnue * (en - es) / dy * 0.5 +
                ((vt - vb) / dz * .5 + (wn - ws) / dy * 0.5) * xnue * (et - eb) / dz * 0.5 +
                dc * (en - es) / dy * 0.5 * xlambda) /
              ec;
  if (nonslip)
    r = r - dt * xnue * vc / ty2 * ALPHA * ec * (1. - ec) *
                (1. - ec);
  rv = r;
  // ---- w :339-376
  r = wc - dt * uc * (we - ww) / dx * 0.5;
  r = r - dt * vc * (wn - ws) / dy * 0.5;
  r = r - dt * wc * (wt - wb) / dz * 0.5;
  r = r + dt * xnue * (we - 2. * wc + ww) / dx / dx;
  r = r + dt * xnue * (wn - 2. * wc + ws) / dy / dy;
  r = r + dt * xnue * (wt - 2. * wc + wb) / dz / dz;
  r = r + dt * (xnue + xlambda) * (dtp - db) / dz * 0.5;
  r = r + dt * (((we - ww) / dx * 0.5 + (ut - ub) / dz * 0.5) * xnue * (ee - ew) / dx * 0.5 +
                ((wn - ws) / dy * 0.5 + (vt - vb) / dz * 0.5) * xnue * (en - es) / dy * 0.5 +
                ((wt - wb) / dz * 0.5 + (wt - wb) / dz * 0.5) * xnue * (et - eb) / dz * 0.5 +
                dc * (et - eb) / dz * 0.5 * xlambda) /
              ec;
  if (nonslip)
    r = r - dt * xnue * wc / tz2 * ALPHA * ec * (1. - ec) *
                (1. - ec);
  rw = r;
}

__global__ void __launch_bounds__(BX *BY) predictor3_kernel(Geo g, Phys ph, Fields f) {
  CELL_IJK(1, 1, 1)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  const long long sx = 1, sy = g.NX, sz = g.plane;
  Stencil3 q;
  q.uc = f.uo[c]; q.ue = f.uo[c + sx]; q.uw = f.uo[c - sx]; q.un = f.uo[c + sy]; q.us = f.uo[c - sy];
  q.ut = f.uo[c + sz]; q.ub = f.uo[c - sz];
  q.vc = f.vo[c]; q.ve = f.vo[c + sx]; q.vw = f.vo[c - sx]; q.vn = f.vo[c + sy]; q.vs = f.vo[c - sy];
  q.vt = f.vo[c + sz]; q.vb = f.vo[c - sz];
  q.wc = f.wo[c]; q.we = f.wo[c + sx]; q.ww = f.wo[c - sx]; q.wn = f.wo[c + sy]; q.ws = f.wo[c - sy];
  q.wt = f.wo[c + sz]; q.wb = f.wo[c - sz];
  q.ec = f.eps[c]; q.ee = f.eps[c + sx]; q.ew = f.eps[c - sx]; q.en = f.eps[c + sy]; q.es = f.eps[c - sy];
  q.et = f.eps[c + sz]; q.eb = f.eps[c - sz];
  q.dc = f.div[c]; q.de = f.div[c + sx]; q.dw = f.div[c - sx]; q.dn = f.div[c + sy]; q.ds = f.div[c - sy];
  q.dtp = f.div[c + sz]; q.db = f.div[c - sz];
  bool fast = ph.ix.fast != 0;
  const double *qv = reinterpret_cast<const double *>(&q);
#pragma unroll
  for (int t = 0; t < (int)(sizeof(Stencil3) / sizeof(double)); ++t) fast = fast && moderate(qv[t]);
  double ru, rv, rw;
  if (fast) predictor3_cell<FastDiv>(ph, q, ru, rv, rw);
  else      predictor3_cell<ExactDiv>(ph, q, ru, rv, rw);
  f.u[c] = ru;
  f.v[c] = rv;
  f.w[c] = rw;
}

// 2D: ibm_2d_uniform_omp_cpu.f90:200-258.  Convection is dt*(u*(du)/dx/2.) here (not /dx*0.5),
// and the v wall force uses dx (sic :254).
__global__ void __launch_bounds__(BX *BY) predictor2_kernel(Geo g, Phys ph, Fields f) {
  CELL_IJK(1, 1, 0)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  const long long sx = 1, sy = g.NX;
  const Inv dx = ph.ix, dy = ph.iy;
  const double dt = ph.dt;
  const double xnue = ph.xnue, xlambda = ph.xlambda;
  const double uc = f.uo[c], ue = f.uo[c + sx], uw = f.uo[c - sx], un = f.uo[c + sy], us = f.uo[c - sy];
  const double vc = f.vo[c], ve = f.vo[c + sx], vw = f.vo[c - sx], vn = f.vo[c + sy], vs = f.vo[c - sy];
  const double ec = f.eps[c], ee = f.eps[c + sx], ew = f.eps[c - sx], en = f.eps[c + sy], es = f.eps[c - sy];
  const double dc = f.div[c], de = f.div[c + sx], dw = f.div[c - sx], dn = f.div[c + sy], ds = f.div[c - sy];
  double r;
  r = uc - dt * (uc * (ue - uw) / dx / 2.);
  r = r - dt * (vc * (un - us) / dy / 2.);
  r = r + dt * xnue * (ue - 2. * uc + uw) / dx / dx;
  r = r + dt * xnue * (un - 2. * uc + us) / dy / dy;
  r = r + dt * (xnue + xlambda) * (de - dw) / dx * .5;
  r = r + dt * (((ue - uw) / dx * .5 + (ue - uw) / dx * .5) * xnue * (ee - ew) / dx * .5 +
                ((un - us) / dy * .5 + (ve - vw) / dx * .5) * xnue * (en - es) / dy * .5 +
                dc * (ee - ew) / dx * 0.5 * xlambda) /
              ec;
  if (ph.nonslip)
    r = r - dt * xnue * uc / ph.itx2 * ALPHA * ec * (1. - ec) *
                (1. - ec);
  f.u[c] = r;
  r = vc - dt * (uc * (ve - vw) / dx / 2.);
  r = r - dt * (vc * (vn - vs) / dy / 2.);
  r = r + dt * xnue * (ve - 2. * vc + vw) / dx / dx;
  r = r + dt * xnue * (vn - 2. * vc + vs) / dy / dy;
  r = r + dt * (xnue + xlambda) * (dn - ds) / dy * .5;
  r = r + dt * (((ve - vw) / dx * .5 + (un - us) / dy * .5) * xnue * (ee - ew) / dx * .5 +
                ((vn - vs) / dy * .5 + (vn - vs) / dy * .5) * xnue * (en - es) / dy * .5 +
                dc * (en - es) / dy * 0.5 * xlambda) /
              ec;
  if (ph.nonslip)
    r = r - dt * xnue * vc / ph.itx2 * ALPHA * ec * (1. - ec) *
                (1. - ec);
  f.v[c] = r;
}

// ----------------------------------------------------------------------------------------------
// Poisson matrix.  Raw coefficients 3D :390-400 / 2D ibm_2d_uniform_omp_cpu.f90:266-269, raw
// right-hand side 3D :404-409 / 2D :272-275, then the boundary-matrix ladder applied per cell.
// a[] order: ae, aw, an, as, at, ab.
// ----------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void raw_coefficients(const Geo &g, const Phys &ph, const double *eps,
                                                 long long c, double a[6]) {
  const double ec = eps[c];
  a[0] = ph.dt * fmax(SMALL, (eps[c + 1] + ec) * 0.5) / ph.ix / ph.ix;
  a[1] = ph.dt * fmax(SMALL, (ec + eps[c - 1]) * 0.5) / ph.ix / ph.ix;
  a[2] = ph.dt * fmax(SMALL, (eps[c + g.NX] + ec) * 0.5) / ph.iy / ph.iy;
  a[3] = ph.dt * fmax(SMALL, (ec + eps[c - g.NX]) * 0.5) / ph.iy / ph.iy;
  if (DIM == 3) {
    a[4] = ph.dt * fmax(SMALL, (eps[c + g.plane] + ec) * 0.5) / ph.iz / ph.iz;
    a[5] = ph.dt * fmax(SMALL, (ec + eps[c - g.plane]) * 0.5) / ph.iz / ph.iz;
  } else {
    a[4] = 0.;
    a[5] = 0.;
  }
}

template <int DIM>
__device__ __forceinline__ double raw_rhs(const Geo &g, const Phys &ph, const Fields &f, long long c) {
  const double *e = f.eps;
  const double ec = e[c];
  double bb = ((e[c + 1] * f.u[c] + ec * f.u[c + 1]) * 0.5 - (e[c - 1] * f.u[c] + ec * f.u[c - 1]) * 0.5) *
                  ph.density / ph.ix +
              ((e[c + g.NX] * f.v[c] + ec * f.v[c + g.NX]) * 0.5 -
               (e[c - g.NX] * f.v[c] + ec * f.v[c - g.NX]) * 0.5) *
                  ph.density / ph.iy;
  if (DIM == 3)
    bb = bb + ((e[c + g.plane] * f.w[c] + ec * f.w[c + g.plane]) * 0.5 -
               (e[c - g.plane] * f.w[c] + ec * f.w[c - g.plane]) * 0.5) *
                  ph.density / ph.iz;
  return bb;
}

__device__ __forceinline__ bool on_matrix_boundary(const Geo &g, const Phys &ph, int i, int j, int kg) {
  if (ph.scase == PF_IBM3_AIRCOND)
    return i == 1 || i == g.m || j == 1 || j == g.n || kg == 1 || kg == g.l;
  return i == 1 || i == g.m;
}

__device__ __forceinline__ void zero6(double a[6]) {
  a[0] = a[1] = a[2] = a[3] = a[4] = a[5] = 0.;
}

// Applies, to ONE cell, the statements of boundrary_matrix / boundary_matrix that touch it, in the
// reference's order.  uniform (3D :636-658, 2D ibm_2d_uniform_omp_cpu.f90:424-437): inlet fold at
// i=1 then outlet Dirichlet fold at i=m.  air-condition (ibm_3d_air_condition_omp_cpu.f90:685-864):
// top, bottom, east, west, north, south, each either a wall fold or, on an outlet face where
// porosity >= 0.9, the Dirichlet fold.  `grow`/`shrink` index a[].
template <int DIM>
__device__ __forceinline__ void boundary_matrix_cell(const Geo &g, const Phys &ph, const Fields &f,
                                                     int i, int j, int kl, long long c, double a[6],
                                                     double &bb) {
  const int kg = kl + g.koff;
  if (ph.scase != PF_IBM3_AIRCOND) {
    if (i == 1) { a[0] = a[0] + a[1]; a[1] = 0.; }
    if (i == g.m) { bb = bb + a[0] * f.p[c + 1]; zero6(a); }
    return;
  }
  const double ec = f.eps[c];
  auto face = [&](bool on, int code, int grow, int shrink, long long halo, bool top_quirk) {
    if (!on) return;
    if (code == 0 || code == 1 || (code == 2 && ec < 0.9)) {
      a[grow] = a[grow] + a[shrink];
      a[shrink] = 0.;
    } else if (code == 2) {
      // top outlet reads bb(i,j,1), sic (:702): plane 1 of this rank, or -- z-slabs -- the plane rank 0 sent
      const double base = top_quirk ? (f.bb1 ? f.bb1[nat_idx(g, i, j, 0)] : raw_rhs<DIM>(g, ph, f, nat_idx(g, i, j, 1))) : bb;
      bb = base + a[shrink] * f.p[halo];
      zero6(a);
    }
  };
  face(kg == g.l, ph.wall[PF_TOP], 5, 4, c + g.plane, true);
  face(kg == 1, ph.wall[PF_BOTTOM], 4, 5, c - g.plane, false);
  face(i == g.m, ph.wall[PF_EAST], 1, 0, c + 1, false);
  face(i == 1, ph.wall[PF_WEST], 0, 1, c - 1, false);
  face(j == g.n, ph.wall[PF_NORTH], 3, 2, c + g.NX, false);
  face(j == 1, ph.wall[PF_SOUTH], 2, 3, c - g.NX, false);
}

__device__ __forceinline__ int cell_colour(const Geo &g, int i, int j, int kl) {
  return (i + j + kl + g.koff) & 1;
}
__device__ __forceinline__ int cell_ih(int i) { return ((i + 1) >> 1) - 1; }

// one-off: the time-invariant coefficients (they depend only on porosity, dt, dx; the reference
// recomputes them every step :386-402) written straight into the checkerboard arrays.
template <int DIM>
__global__ void __launch_bounds__(BX *BY) coefficients_kernel(Geo g, Phys ph, Fields f, SplitSet S0,
                                                              SplitSet S1) {
  CELL_IJK(1, 1, g.kin0)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  double a[6];
  raw_coefficients<DIM>(g, ph, f.eps, c, a);
  double ap;
  if (DIM == 3) ap = -a[0] - a[1] - a[2] - a[3] - a[4] - a[5];   // :402 (before the fold)
  else          ap = -a[0] - a[1] - a[2] - a[3];                 // 2D :270
  double bb = 0.;
  if (on_matrix_boundary(g, ph, i, j, k + g.koff)) boundary_matrix_cell<DIM>(g, ph, f, i, j, k, c, a, bb);
  const SplitSet &S = cell_colour(g, i, j, k) ? S1 : S0;
  const long long h = split_row(g, j, k) + cell_ih(i);
  S.ap[h] = ap;
  S.ae[h] = a[0]; S.aw[h] = a[1]; S.an[h] = a[2]; S.as[h] = a[3];
  if (DIM == 3) { S.at[h] = a[4]; S.ab[h] = a[5]; }
}

// per step: bb (with the Dirichlet fold where a face is an outlet), written in checkerboard layout
template <int DIM>
__global__ void __launch_bounds__(BX *BY) rhs_kernel(Geo g, Phys ph, Fields f, SplitSet S0, SplitSet S1) {
  CELL_IJK(1, 1, g.kin0)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  double bb = raw_rhs<DIM>(g, ph, f, c);
  if (on_matrix_boundary(g, ph, i, j, k + g.koff)) {
    double a[6];
    raw_coefficients<DIM>(g, ph, f.eps, c, a);
    boundary_matrix_cell<DIM>(g, ph, f, i, j, k, c, a, bb);
  }
  const SplitSet &S = cell_colour(g, i, j, k) ? S1 : S0;
  S.bb[split_row(g, j, k) + cell_ih(i)] = bb;
}

__global__ void __launch_bounds__(BX *BY) raw_rhs_plane_kernel(Geo g, Phys ph, Fields f, int kl, double *out) {
  const int i = blockIdx.x * BX + threadIdx.x + 1;
  const int j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > g.m || j > g.n) return;
  out[nat_idx(g, i, j, 0)] = raw_rhs<3>(g, ph, f, nat_idx(g, i, j, kl));
}

// ----------------------------------------------------------------------------------------------
// projection   3D :110-125   2D ibm_2d_uniform_omp_cpu.f90:103-115
// ----------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(BX *BY) project_kernel(Geo g, Phys ph, Fields f) {
  CELL_IJK(1, 1, g.kin0)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  f.u[c] = f.u[c] - ph.dtrho * (f.p[c + 1] - f.p[c - 1]) / ph.ix * 0.5;
  f.v[c] = f.v[c] - ph.dtrho * (f.p[c + g.NX] - f.p[c - g.NX]) / ph.iy * 0.5;
  if (DIM == 3) f.w[c] = f.w[c] - ph.dtrho * (f.p[c + g.plane] - f.p[c - g.plane]) / ph.iz * 0.5;
}

// ----------------------------------------------------------------------------------------------
// boundary conditions, uniform cases.  3D :691-716 (x faces), :719-732 (periodic y);
// 2D ibm_2d_uniform_omp_cpu.f90:476-503 ; backstep inlet * porosity(1,j) ibm_2d_backstep_omp_cpu.f90:533-534
// ----------------------------------------------------------------------------------------------
template <int DIM>
__global__ void bc_uniform_x_kernel(Geo g, Phys ph, Fields f) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int k = (int)blockIdx.y + g.kin0;
  if (j > g.n) return;
  const long long c0 = nat_idx(g, 0, j, k);
  double uin = ph.uin, vin = ph.vin;
  if (ph.scase == PF_IBM2_BACKSTEP) {
    uin = uin * f.eps[c0 + 1];
    vin = vin * f.eps[c0 + 1];
  }
  f.u[c0 + 1] = uin;
  f.v[c0 + 1] = vin;
  f.u[c0] = uin;
  f.v[c0] = vin;
  if (DIM == 3) { f.w[c0 + 1] = 0.; f.w[c0] = 0.; }
  f.p[c0] = f.p[c0 + 2];
  const long long cm = c0 + g.m;
  f.u[cm + 1] = f.u[cm - 1];
  f.v[cm + 1] = f.v[cm - 1];
  if (DIM == 3) f.w[cm + 1] = f.w[cm - 1];
  f.p[cm + 1] = ph.outlet_pressure;
}

template <int DIM>
__global__ void bc_periodic_y_kernel(Geo g, Fields f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = (int)blockIdx.y + g.kin0;
  if (i > g.m + 1) return;
  const long long lo = nat_idx(g, i, 0, k), hi = nat_idx(g, i, g.n + 1, k);
  const long long s1 = nat_idx(g, i, 1, k), sn = nat_idx(g, i, g.n, k);
  f.u[lo] = f.u[sn]; f.v[lo] = f.v[sn]; f.p[lo] = f.p[sn];
  f.u[hi] = f.u[s1]; f.v[hi] = f.v[s1]; f.p[hi] = f.p[s1];
  if (DIM == 3) { f.w[lo] = f.w[sn]; f.w[hi] = f.w[s1]; }
}

// ----------------------------------------------------------------------------------------------
// boundary conditions, air-condition: ibm_3d_air_condition_omp_cpu.f90:873-1170.  One launch per
// face, in the reference's (serial) order top, bottom, west, east, north, south; every face point
// reads and writes only its own line normal to the face, so the points of a face are independent.
// FACE: 0 top 1 bottom 2 west 3 east 4 north 5 south.
// ----------------------------------------------------------------------------------------------
template <int FACE>
__global__ void bc_air_face_kernel(Geo g, Phys ph, Fields f) {
  const int a = blockIdx.x * BX + threadIdx.x;   // fast in-face index
  const int b = blockIdx.y * BY + threadIdx.y;   // slow in-face index
  int i, j, kl;           // the boundary-layer cell
  long long nrm;          // stride towards the ghost cell
  int code;
  if (FACE == 0 || FACE == 1) {          // faces k = l / k = 1, loops i=0..m+1, j=0..n+1
    if (a > g.m + 1 || b > g.n + 1) return;
    i = a; j = b;
    const int kg = (FACE == 0) ? g.l : 1;
    kl = kg - g.koff;
    if (kl < 1 || kl > g.lz) return;     // not on this rank
    nrm = (FACE == 0) ? g.plane : -g.plane;
    code = ph.wall[FACE == 0 ? PF_TOP : PF_BOTTOM];
  } else if (FACE == 2 || FACE == 3) {   // faces i = 1 / i = m, loops j=0..n+1, k=0..l+1
    if (a > g.n + 1 || b > g.lz + 1) return;
    j = a; kl = b;
    i = (FACE == 2) ? 1 : g.m;
    nrm = (FACE == 2) ? -1 : 1;
    code = ph.wall[FACE == 2 ? PF_WEST : PF_EAST];
  } else {                               // faces j = n / j = 1, loops i=0..m+1, k=0..l+1
    if (a > g.m + 1 || b > g.lz + 1) return;
    i = a; kl = b;
    j = (FACE == 4) ? g.n : 1;
    nrm = (FACE == 4) ? g.NX : -(long long)g.NX;
    code = ph.wall[FACE == 4 ? PF_NORTH : PF_SOUTH];
  }
  const long long on = nat_idx(g, i, j, kl), gh = on + nrm, in = on - nrm;
  // fluid test: porosity of the boundary-layer cell; the bottom INLET branch tests porosity(i,j,l), sic (:948)
  double eref = f.eps[on];
  if (FACE == 1 && code == 1) eref = f.eps_top ? f.eps_top[nat_idx(g, i, j, 0)] : f.eps[nat_idx(g, i, j, g.l - g.koff)];
  const bool fluid = eref >= 0.9;
  const double uin = ph.inlet_velocity;
  if (code == 1 && fluid) {
    double iu = 0., iv = 0., iw = 0.;
    if (FACE == 0) iw = -uin;            // :903
    if (FACE == 1) iw = uin;             // :952
    if (FACE == 2) iu = uin;             // :996
    if (FACE == 3) iu = -uin;            // :1041
    if (FACE == 4) iu = -uin;            // :1084 (sic: u, not v)
    if (FACE == 5) iu = uin;             // :1131 (sic)
    f.u[on] = iu; f.v[on] = iv; f.w[on] = iw;
    f.u[gh] = iu; f.v[gh] = iv; f.w[gh] = iw;
    f.p[gh] = f.p[in];
  } else if (code == 2 && fluid) {
    // outlet ghosts: top copies l-1 (:916-918), south copies j=2 (:1149-1151), the others copy the
    // boundary layer itself (:965-967, :1010-1012, :1056-1058, :1098-1100)
    const long long src = (FACE == 0 || FACE == 5) ? in : on;
    f.u[gh] = f.u[src]; f.v[gh] = f.v[src]; f.w[gh] = f.w[src];
    f.p[gh] = ph.outlet_pressure;
  } else {
    f.u[on] = 0.; f.v[on] = 0.; f.w[on] = 0.;
    // mirrored ghost component: top/bottom w, west/east u, north u (sic :1078), south v
    if (FACE == 0 || FACE == 1) f.w[gh] = -f.w[in];
    else if (FACE == 5)         f.v[gh] = -f.v[in];
    else                        f.u[gh] = -f.u[in];
    f.p[gh] = f.p[in];
  }
}

// ----------------------------------------------------------------------------------------------
// initial conditions (3D :778-789 AoA/360, air :1194-1205 zeros, 2D ibm_2d_uniform_omp_cpu.f90:557-565,
// backstep * porosity ibm_2d_backstep_omp_cpu.f90:615-617)
// ----------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(BX *BY) initial_kernel(Geo g, Phys ph, Fields f) {
  CELL_IJK(1, 1, g.kin0)
  if (i > g.m || j > g.n) return;
  const long long c = nat_idx(g, i, j, k);
  double u0 = ph.u0, v0 = ph.v0;
  if (ph.scase == PF_IBM2_BACKSTEP) { u0 = u0 * f.eps[c]; v0 = v0 * f.eps[c]; }
  f.u[c] = u0;
  f.v[c] = v0;
  if (DIM == 3) f.w[c] = 0.;
  f.p[c] = ph.outlet_pressure;
}

// ----------------------------------------------------------------------------------------------
// natural <-> checkerboard conversion of a whole array (halos, edges and corners included)
// ----------------------------------------------------------------------------------------------
template <bool TO_SPLIT>
__global__ void __launch_bounds__(BX *BY) convert_kernel(Geo g, double *nat, double *s0, double *s1) {
  CELL_IJK(0, 0, 0)
  if (i > g.m + 1 || j > g.n + 1) return;
  const long long c = nat_idx(g, i, j, k);
  double *s = cell_colour(g, i, j, k) ? s1 : s0;
  const long long h = split_row(g, j, k) + cell_ih(i);
  if (TO_SPLIT) s[h] = nat[c];
  else          nat[c] = s[h];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static const dim3 kBlock(BX, BY, 1);

void k_divergence(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st) {
  const dim3 grid = cell_grid(g, g.m, g.n, g.lz);
  if (g.dim == 3) divergence_kernel<3><<<grid, kBlock, 0, st>>>(g, ph, f);
  else            divergence_kernel<2><<<grid, kBlock, 0, st>>>(g, ph, f);
  LAUNCHED();
}

void k_div_halo_y(const Geo &g, const Phys &, const Fields &f, cudaStream_t st) {
  div_halo_y_kernel<<<dim3((g.m + 127) / 128, g.lz), 128, 0, st>>>(g, f.div);
  LAUNCHED();
}

void k_plane_copy_interior(const Geo &g, double *a, int kd, int ks, cudaStream_t st) {
  plane_copy_interior_kernel<<<cell_grid(g, g.m, g.n, 1), kBlock, 0, st>>>(g, a, kd, ks);
  LAUNCHED();
}

void k_plane_copy_full(const Geo &g, double *a, int kd, int ks, cudaStream_t st) {
  plane_copy_full_kernel<<<cell_grid(g, g.m + 2, g.n + 2, 1), kBlock, 0, st>>>(g, a, kd, ks);
  LAUNCHED();
}

// halo shell of (d0,d1,d2) <- (s0,s1,s2); s2/d2 null in 2D
void k_shell_copy(const Geo &g, const double *s0, const double *s1, const double *s2, double *d0, double *d1, double *d2,
                  cudaStream_t st) {
  shell_copy_kernel<<<dim3(1, g.n + 2, g.dim == 3 ? g.lz + 2 : 1), 128, 0, st>>>(g, s0, s1, s2, d0, d1, d2);
  LAUNCHED();
}

void k_predictor(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st) {
  const dim3 grid = cell_grid(g, g.m, g.n, g.lz);
  if (g.dim == 3) predictor3_kernel<<<grid, kBlock, 0, st>>>(g, ph, f);
  else            predictor2_kernel<<<grid, kBlock, 0, st>>>(g, ph, f);
  LAUNCHED();
}

void k_coefficients(const Geo &g, const Phys &ph, const Fields &f, const SplitSet S[2], cudaStream_t st) {
  const dim3 grid = cell_grid(g, g.m, g.n, g.lz);
  if (g.dim == 3) coefficients_kernel<3><<<grid, kBlock, 0, st>>>(g, ph, f, S[0], S[1]);
  else            coefficients_kernel<2><<<grid, kBlock, 0, st>>>(g, ph, f, S[0], S[1]);
  LAUNCHED();
}

void k_rhs(const Geo &g, const Phys &ph, const Fields &f, const SplitSet S[2], cudaStream_t st) {
  const dim3 grid = cell_grid(g, g.m, g.n, g.lz);
  if (g.dim == 3) rhs_kernel<3><<<grid, kBlock, 0, st>>>(g, ph, f, S[0], S[1]);
  else            rhs_kernel<2><<<grid, kBlock, 0, st>>>(g, ph, f, S[0], S[1]);
  LAUNCHED();
}

void k_raw_rhs_plane(const Geo &g, const Phys &ph, const Fields &f, int kl, double *out, cudaStream_t st) {
  raw_rhs_plane_kernel<<<cell_grid(g, g.m, g.n, 1), kBlock, 0, st>>>(g, ph, f, kl, out);
  LAUNCHED();
}

void k_project(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st) {
  const dim3 grid = cell_grid(g, g.m, g.n, g.lz);
  if (g.dim == 3) project_kernel<3><<<grid, kBlock, 0, st>>>(g, ph, f);
  else            project_kernel<2><<<grid, kBlock, 0, st>>>(g, ph, f);
  LAUNCHED();
}

// The part of `boundary` that needs no neighbour rank: x faces + periodic y (uniform cases), or the
// six face ladders (air-condition).  The periodic-z / slab-interface plane copies are done by the
// caller (local copies on one rank, exchanges otherwise).
void k_boundary_local(const Geo &g, const Phys &ph, const Fields &f, int, int, cudaStream_t st) {
  if (ph.scase == PF_IBM3_AIRCOND) {
    const dim3 gxy = cell_grid(g, g.m + 2, g.n + 2, 1);
    const dim3 gyz = cell_grid(g, g.n + 2, g.lz + 2, 1);
    const dim3 gxz = cell_grid(g, g.m + 2, g.lz + 2, 1);
    bc_air_face_kernel<0><<<gxy, kBlock, 0, st>>>(g, ph, f);
    bc_air_face_kernel<1><<<gxy, kBlock, 0, st>>>(g, ph, f);
    bc_air_face_kernel<2><<<gyz, kBlock, 0, st>>>(g, ph, f);
    bc_air_face_kernel<3><<<gyz, kBlock, 0, st>>>(g, ph, f);
    bc_air_face_kernel<4><<<gxz, kBlock, 0, st>>>(g, ph, f);
    bc_air_face_kernel<5><<<gxz, kBlock, 0, st>>>(g, ph, f);
    g_launches += 6;
    return;
  }
  if (g.dim == 3) {
    bc_uniform_x_kernel<3><<<dim3((g.n + 127) / 128, g.lz), 128, 0, st>>>(g, ph, f);
    bc_periodic_y_kernel<3><<<dim3((g.m + 2 + 127) / 128, g.lz), 128, 0, st>>>(g, f);
  } else {
    bc_uniform_x_kernel<2><<<dim3((g.n + 127) / 128, 1), 128, 0, st>>>(g, ph, f);
    bc_periodic_y_kernel<2><<<dim3((g.m + 2 + 127) / 128, 1), 128, 0, st>>>(g, f);
  }
  g_launches += 2;
}

void k_initial(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st) {
  const dim3 grid = cell_grid(g, g.m, g.n, g.lz);
  if (g.dim == 3) initial_kernel<3><<<grid, kBlock, 0, st>>>(g, ph, f);
  else            initial_kernel<2><<<grid, kBlock, 0, st>>>(g, ph, f);
  LAUNCHED();
}

void k_nat_to_split(const Geo &g, const double *nat, double *s0, double *s1, cudaStream_t st) {
  convert_kernel<true><<<cell_grid(g, g.m + 2, g.n + 2, g.NZ), kBlock, 0, st>>>(g, const_cast<double *>(nat), s0, s1);
  LAUNCHED();
}

void k_split_to_nat(const Geo &g, const double *s0, const double *s1, double *nat, cudaStream_t st) {
  convert_kernel<false><<<cell_grid(g, g.m + 2, g.n + 2, g.NZ), kBlock, 0, st>>>(
      g, nat, const_cast<double *>(s0), const_cast<double *>(s1));
  LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// self-check of the exact reciprocal division (pf_internal.cuh, struct Inv): counts the inputs for
// which `a / Inv{d}` differs from the IEEE quotient a / d.  Must return 0 (tests/test_gpu_parity.py).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void fastdiv_check_kernel(Inv inv, long long n, unsigned long long seed, unsigned long long *bad) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (; t < n; t += stride) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(t + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    // random sign and mantissa, exponent in [-60, 60] around 1.0 (plus exact zero now and then)
    const unsigned long long mant = x & 0xFFFFFFFFFFFFFull;
    const int e = (int)((x >> 52) % 121) - 60;
    const unsigned long long sign = (x >> 63) << 63;
    double a = __longlong_as_double((long long)(sign | ((unsigned long long)(1023 + e) << 52) | mant));
    if ((x & 0xFFF000) == 0) a = 0.0;
    const double q1 = a / inv;
    const double q2 = a / inv.d;
    if (__double_as_longlong(q1) != __double_as_longlong(q2)) ++local;
  }
  if (local) atomicAdd(bad, local);
}
}  // namespace

extern "C" int pf_debug_fastdiv_mismatches(double d, long long n, unsigned long long seed, long long *mismatches) {
  if (!mismatches) return 1;
  Inv inv;
  inv.d = d;
  inv.r = 1.0 / d;
  inv.fast = 1;
  unsigned long long *bad = nullptr;
  if (cudaMalloc(&bad, sizeof(*bad)) != cudaSuccess) return 1;
  cudaMemset(bad, 0, sizeof(*bad));
  fastdiv_check_kernel<<<pf_sm_count() * 8, 256>>>(inv, n, seed, bad);
  unsigned long long h = 0;
  const cudaError_t e = cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(bad);
  *mismatches = (long long)h;
  return e == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
// drag / lift force log of ibm2_drag: lib/output.f90:244-305 (output_force_log_2d), called after every
// step by ibm_2d_drag_omp_cpu.f90:121.  Per-cell terms are evaluated exactly as typed; the four sums
// are formed deterministically (fixed tree inside a block, partials added in block order), so they
// agree with the reference's serial sums to rounding (not bit for bit: the summation order differs).
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int FB = 256;
__global__ void __launch_bounds__(FB) force2d_partial_kernel(Geo g, Phys ph, Fields f, double *partial) {
  constexpr double small = 1.e-6, alpha = 32.0;
  const long long ncell = (long long)g.m * g.n;
  double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
  for (long long t = blockIdx.x * (long long)FB + threadIdx.x; t < ncell; t += (long long)gridDim.x * FB) {
    const int i = (int)(t % g.m) + 1, j = (int)(t / g.m) + 1;
    const long long c = nat_idx(g, i, j, 0);
    const double e = f.eps[c];
    const double gx = (f.eps[c + 1] - f.eps[c - 1]) * 0.5, gy = (f.eps[c + g.NX] - f.eps[c - g.NX]) * 0.5;
    const double normal_abs = sqrt(gx * gx + gy * gy);
    const double nx = gx / fmax(normal_abs, small), ny = gy / fmax(normal_abs, small);
    const double dx = ph.dx, dy = ph.dy, th = ph.thickness;
    s0 += -dx * dy * f.p[c] * 2 * e * (1.0 - e) / (th * dx) * nx;
    s1 += -dx * dy * f.p[c] * 2 * e * (1.0 - e) / (th * dy) * ny;
    const double qx = (e * (1.0 - e)) / (th * dx), qy = (e * (1.0 - e)) / (th * dy);
    s2 += +dx * dy * alpha * ph.density * ph.xnue * (qx * qx) * f.u[c];
    s3 += +dx * dy * alpha * ph.density * ph.xnue * (qy * qy) * f.v[c];
  }
  __shared__ double sh[4][FB];
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; sh[2][threadIdx.x] = s2; sh[3][threadIdx.x] = s3;
  __syncthreads();
  for (int o = FB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 4; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[4 * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void force2d_final_kernel(const double *partial, int nblocks, double *out4) {
  if (threadIdx.x < 4) {
    double s = 0.;
    for (int b = 0; b < nblocks; ++b) s += partial[4 * b + threadIdx.x];
    out4[threadIdx.x] = s;
  }
}
// output_force_log_3d (lib/output.f90:1090-1165): the same sums in 3D, six of them, over this rank's planes
__global__ void __launch_bounds__(FB) force3d_partial_kernel(Geo g, Phys ph, Fields f, double *partial) {
  constexpr double small = 1.e-6, alpha = 32.0;
  const long long ncell = (long long)g.m * g.n * g.lz;
  double s[6] = {0., 0., 0., 0., 0., 0.};
  const double dx = ph.dx, dy = ph.dy, dz = ph.dz, th = ph.thickness;
  for (long long t = blockIdx.x * (long long)FB + threadIdx.x; t < ncell; t += (long long)gridDim.x * FB) {
    const int i = (int)(t % g.m) + 1;
    const long long r = t / g.m;
    const int j = (int)(r % g.n) + 1, k = (int)(r / g.n) + 1;
    const long long c = nat_idx(g, i, j, k);
    const double e = f.eps[c];
    const double gx = (f.eps[c + 1] - f.eps[c - 1]) * 0.5, gy = (f.eps[c + g.NX] - f.eps[c - g.NX]) * 0.5,
                 gz = (f.eps[c + g.plane] - f.eps[c - g.plane]) * 0.5;
    const double normal_abs = sqrt(gx * gx + gy * gy + gz * gz);
    const double den = fmax(normal_abs, small);
    const double nx = gx / den, ny = gy / den, nz = gz / den;
    const double pp = f.p[c];
    s[0] += -dx * dy * dz * pp * 2 * e * (1.0 - e) / (th * dx) * nx;
    s[1] += -dx * dy * dz * pp * 2 * e * (1.0 - e) / (th * dy) * ny;
    s[2] += -dx * dy * dz * pp * 2 * e * (1.0 - e) / (th * dz) * nz;
    const double qx = (e * (1.0 - e)) / (th * dx), qy = (e * (1.0 - e)) / (th * dy), qz = (e * (1.0 - e)) / (th * dz);
    s[3] += +dx * dy * dz * alpha * ph.density * ph.xnue * (qx * qx) * f.u[c];
    s[4] += +dx * dy * dz * alpha * ph.density * ph.xnue * (qy * qy) * f.v[c];
    s[5] += +dx * dy * dz * alpha * ph.density * ph.xnue * (qz * qz) * f.w[c];
  }
  __shared__ double sh[6][FB];
  for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] = s[q];
  __syncthreads();
  for (int o = FB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 6) partial[6 * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void force3d_final_kernel(const double *partial, int nblocks, double *out6) {
  if (threadIdx.x < 6) {
    double s = 0.;
    for (int b = 0; b < nblocks; ++b) s += partial[6 * b + threadIdx.x];
    out6[threadIdx.x] = s;
  }
}
}  // namespace

// partial: >= 6*blocks doubles of scratch; out6: Fpx, Fpy, Fpz, Fvx, Fvy, Fvz of this rank's slab
void k_force3d(const Geo &g, const Phys &ph, const Fields &f, double *partial, int blocks, double *out6, cudaStream_t st) {
  force3d_partial_kernel<<<blocks, FB, 0, st>>>(g, ph, f, partial);
  force3d_final_kernel<<<1, 32, 0, st>>>(partial, blocks, out6);
  g_launches += 2;
}

// partial: >= 4*blocks doubles of scratch; out4: Fpx, Fpy, Fvx, Fvy
void k_force2d(const Geo &g, const Phys &ph, const Fields &f, double *partial, int blocks, double *out4, cudaStream_t st) {
  force2d_partial_kernel<<<blocks, FB, 0, st>>>(g, ph, f, partial);
  force2d_final_kernel<<<1, 32, 0, st>>>(partial, blocks, out4);
  g_launches += 2;
}
