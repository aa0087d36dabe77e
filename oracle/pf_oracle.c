/*
 * pf_oracle.c -- CPU ORACLE for the PixelFlow per-timestep hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (pixelflow_b200/csrc) never links, calls or falls back to anything in oracle/.
 *
 * What it is: a plain-C restatement, at fp64, of the hot path of the five
 * src/omp_parallel programs of nobu-n2002/PixelFlow, keeping the reference's loop
 * structure (separate u/v/w sweeps, stored div, two full p_old copies per SOR
 * iteration, linear-index -> (i,j,k) div/mod colour mapping, separate error pass),
 * statement order and expression association, so that it evaluates what
 * `gfortran -O3 -fopenmp -fdefault-real-8` (no FMA contraction) would evaluate.
 * Build with -ffp-contract=off and without -ffast-math (oracle/Makefile).
 *
 * PARITY STATUS: pinned against the reference's own source, machine-translated and run
 * here.  The reference ships no golden vectors / expected outputs for this path and no
 * Fortran compiler exists in this environment; oracle/f90toc.py
 * translates the reference's Fortran files mechanically into C (oracle/build_ref.py ->
 * oracle/_ref/, git-ignored), oracle/ref_translated.py runs the translated programs on
 * project directories, and tests/test_ref_translation.py shows this restatement
 * bit-identical to them (all five programs, seeded decks, the three shipped decks; the
 * outputs are committed as tests/golden/ref_translated.npz).  Not a gfortran build.
 * Also: (1) every function cites the file:line it transcribes, (2) an independent numpy
 * restatement (oracle/oracle_np.py) agrees bit for bit (tests/test_oracle_cross.py),
 * (3) invariants (uniform-flow fixed point, colour coverage, in-place SOR == p_old-copy SOR).
 *
 * Array layout: the reference's `real, dimension(0:md,0:nd,0:ld)` column-major arrays
 * become dense C arrays of logical shape [l+2][n+2][m+2] (2D: [n+2][m+2]), i fastest:
 *   A(i,j,k)  ->  a[i + (m+2)*(j + (n+2)*k)],  0<=i<=m+1, 0<=j<=n+1, 0<=k<=l+1.
 * The reference's arrays are static and zero-initialised (-fno-automatic / BSS); callers
 * must pass zero-initialised arrays, and the solver-local arrays (ap..bb, div, p_old)
 * live in a calloc'ed workspace that persists across steps for the same reason.
 *
 * Reference files restated (all under /root/reference/src/omp_parallel/):
 *   ibm_3d_uniform_omp_cpu.f90        (ibm3 uniform:  inlet x=1, outlet x=m, periodic y,z)
 *   ibm_3d_air_condition_omp_cpu.f90  (ibm3 room:     six configurable wall/inlet/outlet faces)
 *   ibm_2d_uniform_omp_cpu.f90        (ibm2 uniform / drag: inlet, outlet, periodic y)
 *   ibm_2d_backstep_omp_cpu.f90       (ibm2 backstep: inlet and initial velocity times porosity)
 *   lib/grid.f90                      (porosity halo rules)
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define PFO_EXPORT __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------- */
/* parameters                                                                          */
/* ---------------------------------------------------------------------------------- */
typedef struct pfo_params {
  int m, n, l;               /* interior cells; l ignored in 2D                          */
  double dx, dy, dz, dt;     /* lib/grid.f90:297-300                                     */
  double xnue, xlambda, density, thickness;
  int nonslip;               /* logical                                                  */
  int iter_max;
  double relux_factor;
  double inlet_velocity, outlet_pressure, AoA;
  /* module wall_conditions, ibm_3d_air_condition_omp_cpu.f90:4-16
   * order: top(k=l), bottom(k=1), east(i=m), west(i=1), south(j=1), north(j=n)
   * 0 wall, 1 inlet where porosity>=0.9, 2 outlet where porosity>=0.9                  */
  int wall[6];
} pfo_params;

enum { PFO_TOP = 0, PFO_BOTTOM = 1, PFO_EAST = 2, PFO_WEST = 3, PFO_SOUTH = 4, PFO_NORTH = 5 };

/* reference parameters `small`, `alpha`: ibm_3d_uniform_omp_cpu.f90:169-170 */
static const double SMALL = 1.e-6;
static const double ALPHA = 32.0;

/* solver-local static arrays of the reference (solve_p locals :172, p_old :445) */
typedef struct pfo_ws {
  size_t nelem;
  double *ap, *ae, *aw, *an, *as, *at, *ab, *bb, *div, *p_old;
} pfo_ws;

PFO_EXPORT pfo_ws *pfo_ws_create(size_t nelem) {
  pfo_ws *ws = (pfo_ws *)calloc(1, sizeof(pfo_ws));
  if (!ws) return NULL;
  ws->nelem = nelem;
  double **arr[10] = {&ws->ap, &ws->ae, &ws->aw, &ws->an, &ws->as,
                      &ws->at, &ws->ab, &ws->bb, &ws->div, &ws->p_old};
  for (int a = 0; a < 10; ++a) {
    *arr[a] = (double *)calloc(nelem, sizeof(double));
    if (!*arr[a]) return NULL;
  }
  return ws;
}

PFO_EXPORT void pfo_ws_destroy(pfo_ws *ws) {
  if (!ws) return;
  free(ws->ap); free(ws->ae); free(ws->aw); free(ws->an); free(ws->as);
  free(ws->at); free(ws->ab); free(ws->bb); free(ws->div); free(ws->p_old);
  free(ws);
}

/* expose workspace arrays to the tests: 0 ap,1 ae,2 aw,3 an,4 as,5 at,6 ab,7 bb,8 div,9 p_old */
PFO_EXPORT double *pfo_ws_array(pfo_ws *ws, int which) {
  switch (which) {
    case 0: return ws->ap; case 1: return ws->ae; case 2: return ws->aw;
    case 3: return ws->an; case 4: return ws->as; case 5: return ws->at;
    case 6: return ws->ab; case 7: return ws->bb; case 8: return ws->div;
    case 9: return ws->p_old;
  }
  return NULL;
}

static inline double dmax(double a, double b) { return (a > b) ? a : b; }

/* ================================================================================== */
/*                                   3 D                                               */
/* ================================================================================== */
#define LX ((size_t)(P->m + 2))
#define LY ((size_t)(P->n + 2))
#define I3(i, j, k) ((size_t)(i) + LX * ((size_t)(j) + LY * (size_t)(k)))

/* lib/grid.f90:349-378 (grid_conditions_yz_periodic): x zero-gradient on j=1..n+1,k=1..l+1
 * only, then periodic y over all i,k, then periodic z over all i,j. */
PFO_EXPORT void pfo3u_porosity_halo(const pfo_params *P, double *porosity) {
  const int m = P->m, n = P->n, l = P->l;
  for (int j = 1; j <= n + 1; ++j)
    for (int k = 1; k <= l + 1; ++k) {
      porosity[I3(0, j, k)] = porosity[I3(1, j, k)];
      porosity[I3(m + 1, j, k)] = porosity[I3(m, j, k)];
    }
  for (int i = 0; i <= m + 1; ++i)
    for (int k = 0; k <= l + 1; ++k) {
      porosity[I3(i, 0, k)] = porosity[I3(i, n, k)];
      porosity[I3(i, n + 1, k)] = porosity[I3(i, 1, k)];
    }
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      porosity[I3(i, j, 0)] = porosity[I3(i, j, l)];
      porosity[I3(i, j, l + 1)] = porosity[I3(i, j, 1)];
    }
}

/* lib/grid.f90:215-243 (grid_conditions_wall): zero-gradient on all six faces, x then y then z */
PFO_EXPORT void pfo3a_porosity_halo(const pfo_params *P, double *porosity) {
  const int m = P->m, n = P->n, l = P->l;
  for (int j = 0; j <= n + 1; ++j)
    for (int k = 0; k <= l + 1; ++k) {
      porosity[I3(0, j, k)] = porosity[I3(1, j, k)];
      porosity[I3(m + 1, j, k)] = porosity[I3(m, j, k)];
    }
  for (int i = 0; i <= m + 1; ++i)
    for (int k = 0; k <= l + 1; ++k) {
      porosity[I3(i, 0, k)] = porosity[I3(i, 1, k)];
      porosity[I3(i, n + 1, k)] = porosity[I3(i, n, k)];
    }
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      porosity[I3(i, j, 0)] = porosity[I3(i, j, 1)];
      porosity[I3(i, j, l + 1)] = porosity[I3(i, j, l)];
    }
}

/* ibm_3d_uniform_omp_cpu.f90:756-793 (AoA/360) ; air: ibm_3d_air_condition_omp_cpu.f90:1174-1211 */
PFO_EXPORT void pfo3_initial_conditions(const pfo_params *P, int air, double *p, double *u,
                                        double *v, double *w) {
  const int m = P->m, n = P->n, l = P->l;
  const double pi = atan(1.) * 4.;
  const double u0 = air ? 0. : P->inlet_velocity * cos(P->AoA / 360 * pi);
  const double v0 = air ? 0. : P->inlet_velocity * sin(P->AoA / 360 * pi);
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        u[I3(i, j, k)] = u0;
        v[I3(i, j, k)] = v0;
        w[I3(i, j, k)] = 0.;
        p[I3(i, j, k)] = P->outlet_pressure;
      }
}

/* ibm_3d_uniform_omp_cpu.f90:85-100 (air :99-114): u_old = u incl. halos */
PFO_EXPORT void pfo3_copy_old(const pfo_params *P, const double *u, const double *v,
                              const double *w, double *u_old, double *v_old, double *w_old) {
  const int m = P->m, n = P->n, l = P->l;
#pragma omp parallel for
  for (int k = 0; k <= l + 1; ++k)
    for (int j = 0; j <= n + 1; ++j)
      for (int i = 0; i <= m + 1; ++i) {
        u_old[I3(i, j, k)] = u[I3(i, j, k)];
        v_old[I3(i, j, k)] = v[I3(i, j, k)];
        w_old[I3(i, j, k)] = w[I3(i, j, k)];
      }
}

/* divergence: ibm_3d_uniform_omp_cpu.f90:185-222 ; air halos all zero :210-237 */
PFO_EXPORT void pfo3_divergence(const pfo_params *P, int air, const double *u_old,
                                const double *v_old, const double *w_old, double *div) {
  const int m = P->m, n = P->n, l = P->l;
  const double dx = P->dx, dy = P->dy, dz = P->dz;
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i)
        div[I3(i, j, k)] = (u_old[I3(i + 1, j, k)] - u_old[I3(i - 1, j, k)]) / dx * 0.5 +
                           (v_old[I3(i, j + 1, k)] - v_old[I3(i, j - 1, k)]) / dy * 0.5 +
                           (w_old[I3(i, j, k + 1)] - w_old[I3(i, j, k - 1)]) / dz * 0.5;
  if (!air) {
    for (int k = 1; k <= l; ++k)
      for (int j = 1; j <= n; ++j) {
        div[I3(0, j, k)] = 0.;
        div[I3(m + 1, j, k)] = 0.;
      }
    for (int k = 1; k <= l; ++k)
      for (int i = 1; i <= m; ++i) {
        div[I3(i, 0, k)] = div[I3(i, n, k)];
        div[I3(i, n + 1, k)] = div[I3(i, 1, k)];
      }
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        div[I3(i, j, 0)] = div[I3(i, j, l)];
        div[I3(i, j, l + 1)] = div[I3(i, j, 1)];
      }
  } else {
    for (int i = 1; i <= m; ++i)
      for (int j = 1; j <= n; ++j) {
        div[I3(i, j, 0)] = 0.;
        div[I3(i, j, l + 1)] = 0.;
      }
    for (int j = 1; j <= n; ++j)
      for (int k = 1; k <= l; ++k) {
        div[I3(0, j, k)] = 0.;
        div[I3(m + 1, j, k)] = 0.;
      }
    for (int i = 1; i <= m; ++i)
      for (int k = 1; k <= l; ++k) {
        div[I3(i, 0, k)] = 0.;
        div[I3(i, n + 1, k)] = 0.;
      }
  }
}

/* predictor: ibm_3d_uniform_omp_cpu.f90:228-381 (identical in air :275-428).
 * Three separate sweeps, each a chain of in-place statements; kept verbatim. */
PFO_EXPORT void pfo3_predictor(const pfo_params *P, const double *u_old, const double *v_old,
                               const double *w_old, const double *porosity, const double *div,
                               double *u, double *v, double *w) {
  const int m = P->m, n = P->n, l = P->l;
  const double dx = P->dx, dy = P->dy, dz = P->dz, dt = P->dt;
  const double xnue = P->xnue, xlambda = P->xlambda, thickness = P->thickness;
  const int nonslip = P->nonslip;
#define U(a, b, c) u_old[I3(a, b, c)]
#define V(a, b, c) v_old[I3(a, b, c)]
#define W(a, b, c) w_old[I3(a, b, c)]
#define E(a, b, c) porosity[I3(a, b, c)]
#define D(a, b, c) div[I3(a, b, c)]
  /* velocity u :228-276 */
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        double r;
        r = U(i, j, k) - dt * U(i, j, k) * (U(i + 1, j, k) - U(i - 1, j, k)) / dx * 0.5;
        r = r - dt * V(i, j, k) * (U(i, j + 1, k) - U(i, j - 1, k)) / dy * 0.5;
        r = r - dt * W(i, j, k) * (U(i, j, k + 1) - U(i, j, k - 1)) / dz * 0.5;
        r = r + dt * xnue * (U(i + 1, j, k) - 2. * U(i, j, k) + U(i - 1, j, k)) / dx / dx;
        r = r + dt * xnue * (U(i, j + 1, k) - 2. * U(i, j, k) + U(i, j - 1, k)) / dy / dy;
        r = r + dt * xnue * (U(i, j, k + 1) - 2. * U(i, j, k) + U(i, j, k - 1)) / dz / dz;
        r = r + dt * (xnue + xlambda) * (D(i + 1, j, k) - D(i - 1, j, k)) / dx * 0.5;
        r = r + dt * (((U(i + 1, j, k) - U(i - 1, j, k)) / dx * 0.5 +
                       (U(i + 1, j, k) - U(i - 1, j, k)) / dx * 0.5) *
                          xnue * (E(i + 1, j, k) - E(i - 1, j, k)) / dx * 0.5 +
                      ((U(i, j + 1, k) - U(i, j - 1, k)) / dy * 0.5 +
                       (V(i + 1, j, k) - V(i - 1, j, k)) / dx * 0.5) *
                          xnue * (E(i, j + 1, k) - E(i, j - 1, k)) / dy * 0.5 +
                      ((U(i, j, k + 1) - U(i, j, k - 1)) / dz * 0.5 +
                       (W(i + 1, j, k) - W(i - 1, j, k)) / dx * 0.5) *
                          xnue * (E(i, j, k + 1) - E(i, j, k - 1)) / dz * 0.5 +
                      D(i, j, k) * (E(i + 1, j, k) - E(i - 1, j, k)) / dx * 0.5 * xlambda) /
                    E(i, j, k);
        if (nonslip)
          r = r - dt * xnue * U(i, j, k) / ((thickness * dx) * (thickness * dx)) * ALPHA *
                      E(i, j, k) * (1. - E(i, j, k)) * (1. - E(i, j, k));
        u[I3(i, j, k)] = r;
      }
  /* velocity v :281-328 */
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        double r;
        r = V(i, j, k) - dt * U(i, j, k) * (V(i + 1, j, k) - V(i - 1, j, k)) / dx * 0.5;
        r = r - dt * V(i, j, k) * (V(i, j + 1, k) - V(i, j - 1, k)) / dy * 0.5;
        r = r - dt * W(i, j, k) * (V(i, j, k + 1) - V(i, j, k - 1)) / dz * 0.5;
        r = r + dt * xnue * (V(i + 1, j, k) - 2. * V(i, j, k) + V(i - 1, j, k)) / dx / dx;
        r = r + dt * xnue * (V(i, j + 1, k) - 2. * V(i, j, k) + V(i, j - 1, k)) / dy / dy;
        r = r + dt * xnue * (V(i, j, k + 1) - 2. * V(i, j, k) + V(i, j, k - 1)) / dz / dz;
        r = r + dt * (xnue + xlambda) * (D(i, j + 1, k) - D(i, j - 1, k)) / dy * 0.5;
        r = r + dt * (((V(i + 1, j, k) - V(i - 1, j, k)) / dx * 0.5 +
                       (U(i, j + 1, k) - U(i, j - 1, k)) / dy * 0.5) *
                          xnue * (E(i + 1, j, k) - E(i - 1, j, k)) / dx * 0.5 +
                      ((V(i, j + 1, k) - V(i, j - 1, k)) / dy * .5 +
                       (V(i, j + 1, k) - V(i, j - 1, k)) / dy * 0.5) *
                          xnue * (E(i, j + 1, k) - E(i, j - 1, k)) / dy * 0.5 +
                      ((V(i, j, k + 1) - V(i, j, k - 1)) / dz * .5 +
                       (W(i, j + 1, k) - W(i, j - 1, k)) / dy * 0.5) *
                          xnue * (E(i, j, k + 1) - E(i, j, k - 1)) / dz * 0.5 +
                      D(i, j, k) * (E(i, j + 1, k) - E(i, j - 1, k)) / dy * 0.5 * xlambda) /
                    E(i, j, k);
        if (nonslip)
          r = r - dt * xnue * V(i, j, k) / ((thickness * dy) * (thickness * dy)) * ALPHA *
                      E(i, j, k) * (1. - E(i, j, k)) * (1. - E(i, j, k));
        v[I3(i, j, k)] = r;
      }
  /* velocity w :334-381 */
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        double r;
        r = W(i, j, k) - dt * U(i, j, k) * (W(i + 1, j, k) - W(i - 1, j, k)) / dx * 0.5;
        r = r - dt * V(i, j, k) * (W(i, j + 1, k) - W(i, j - 1, k)) / dy * 0.5;
        r = r - dt * W(i, j, k) * (W(i, j, k + 1) - W(i, j, k - 1)) / dz * 0.5;
        r = r + dt * xnue * (W(i + 1, j, k) - 2. * W(i, j, k) + W(i - 1, j, k)) / dx / dx;
        r = r + dt * xnue * (W(i, j + 1, k) - 2. * W(i, j, k) + W(i, j - 1, k)) / dy / dy;
        r = r + dt * xnue * (W(i, j, k + 1) - 2. * W(i, j, k) + W(i, j, k - 1)) / dz / dz;
        r = r + dt * (xnue + xlambda) * (D(i, j, k + 1) - D(i, j, k - 1)) / dz * 0.5;
        r = r + dt * (((W(i + 1, j, k) - W(i - 1, j, k)) / dx * 0.5 +
                       (U(i, j, k + 1) - U(i, j, k - 1)) / dz * 0.5) *
                          xnue * (E(i + 1, j, k) - E(i - 1, j, k)) / dx * 0.5 +
                      ((W(i, j + 1, k) - W(i, j - 1, k)) / dy * 0.5 +
                       (V(i, j, k + 1) - V(i, j, k - 1)) / dz * 0.5) *
                          xnue * (E(i, j + 1, k) - E(i, j - 1, k)) / dy * 0.5 +
                      ((W(i, j, k + 1) - W(i, j, k - 1)) / dz * 0.5 +
                       (W(i, j, k + 1) - W(i, j, k - 1)) / dz * 0.5) *
                          xnue * (E(i, j, k + 1) - E(i, j, k - 1)) / dz * 0.5 +
                      D(i, j, k) * (E(i, j, k + 1) - E(i, j, k - 1)) / dz * 0.5 * xlambda) /
                    E(i, j, k);
        if (nonslip)
          r = r - dt * xnue * W(i, j, k) / ((thickness * dz) * (thickness * dz)) * ALPHA *
                      E(i, j, k) * (1. - E(i, j, k)) * (1. - E(i, j, k));
        w[I3(i, j, k)] = r;
      }
#undef U
#undef V
#undef W
#undef D
}

/* matrix build: ibm_3d_uniform_omp_cpu.f90:386-414 (air :432-460) */
PFO_EXPORT void pfo3_matrix(const pfo_params *P, const double *u, const double *v,
                            const double *w, const double *porosity, pfo_ws *ws) {
  const int m = P->m, n = P->n, l = P->l;
  const double dx = P->dx, dy = P->dy, dz = P->dz, dt = P->dt, density = P->density;
  double *ap = ws->ap, *ae = ws->ae, *aw = ws->aw, *an = ws->an, *as = ws->as, *at = ws->at,
         *ab = ws->ab, *bb = ws->bb;
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        const size_t c = I3(i, j, k);
        ae[c] = dt * dmax(SMALL, (E(i + 1, j, k) + E(i, j, k)) * 0.5) / dx / dx;
        aw[c] = dt * dmax(SMALL, (E(i, j, k) + E(i - 1, j, k)) * 0.5) / dx / dx;
        an[c] = dt * dmax(SMALL, (E(i, j + 1, k) + E(i, j, k)) * 0.5) / dy / dy;
        as[c] = dt * dmax(SMALL, (E(i, j, k) + E(i, j - 1, k)) * 0.5) / dy / dy;
        at[c] = dt * dmax(SMALL, (E(i, j, k + 1) + E(i, j, k)) * 0.5) / dz / dz;
        ab[c] = dt * dmax(SMALL, (E(i, j, k) + E(i, j, k - 1)) * 0.5) / dz / dz;
        ap[c] = -ae[c] - aw[c] - an[c] - as[c] - at[c] - ab[c];
        bb[c] = ((E(i + 1, j, k) * u[c] + E(i, j, k) * u[I3(i + 1, j, k)]) * 0.5 -
                 (E(i - 1, j, k) * u[c] + E(i, j, k) * u[I3(i - 1, j, k)]) * 0.5) *
                    density / dx +
                ((E(i, j + 1, k) * v[c] + E(i, j, k) * v[I3(i, j + 1, k)]) * 0.5 -
                 (E(i, j - 1, k) * v[c] + E(i, j, k) * v[I3(i, j - 1, k)]) * 0.5) *
                    density / dy +
                ((E(i, j, k + 1) * w[c] + E(i, j, k) * w[I3(i, j, k + 1)]) * 0.5 -
                 (E(i, j, k - 1) * w[c] + E(i, j, k) * w[I3(i, j, k - 1)]) * 0.5) *
                    density / dz;
      }
}

/* boundrary_matrix (sic), uniform: ibm_3d_uniform_omp_cpu.f90:618-663 */
PFO_EXPORT void pfo3u_boundary_matrix(const pfo_params *P, const double *p, pfo_ws *ws) {
  const int m = P->m, n = P->n, l = P->l;
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j) {
      ws->ae[I3(1, j, k)] = ws->ae[I3(1, j, k)] + ws->aw[I3(1, j, k)];
      ws->aw[I3(1, j, k)] = 0.;
    }
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j) {
      const size_t c = I3(m, j, k);
      ws->bb[c] = ws->bb[c] + ws->ae[c] * p[I3(m + 1, j, k)];
      ws->ae[c] = 0.; ws->aw[c] = 0.; ws->an[c] = 0.;
      ws->as[c] = 0.; ws->at[c] = 0.; ws->ab[c] = 0.;
    }
}

/* boundary_matrix, air-condition: ibm_3d_air_condition_omp_cpu.f90:665-867.
 * Faces in the reference's order top, bottom, east, west, north, south; loops include the
 * halo indices; the dead-branch quirk bb(i,j,l)=bb(i,j,1)+... (:702) is kept verbatim. */
static void zero6(pfo_ws *ws, size_t c) {
  ws->ae[c] = 0.; ws->aw[c] = 0.; ws->an[c] = 0.;
  ws->as[c] = 0.; ws->at[c] = 0.; ws->ab[c] = 0.;
}
PFO_EXPORT void pfo3a_boundary_matrix(const pfo_params *P, const double *p,
                                      const double *porosity, pfo_ws *ws) {
  const int m = P->m, n = P->n, l = P->l;
  const int top = P->wall[PFO_TOP], bottom = P->wall[PFO_BOTTOM], east = P->wall[PFO_EAST],
            west = P->wall[PFO_WEST], south = P->wall[PFO_SOUTH], north = P->wall[PFO_NORTH];
  double *ae = ws->ae, *aw = ws->aw, *an = ws->an, *as = ws->as, *at = ws->at, *ab = ws->ab,
         *bb = ws->bb;
  /* top :686-714 */
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      const size_t c = I3(i, j, l);
      if (top == 0 || top == 1 || (top == 2 && E(i, j, l) < 0.9)) {
        ab[c] = ab[c] + at[c]; at[c] = 0.;
      } else if (top == 2) {
        bb[c] = bb[I3(i, j, 1)] + at[c] * p[I3(i, j, l + 1)];
        zero6(ws, c);
      }
    }
  /* bottom :717-745 */
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      const size_t c = I3(i, j, 1);
      if (bottom == 0 || bottom == 1 || (bottom == 2 && E(i, j, 1) < 0.9)) {
        at[c] = at[c] + ab[c]; ab[c] = 0.;
      } else if (bottom == 2) {
        bb[c] = bb[c] + ab[c] * p[I3(i, j, 0)];
        zero6(ws, c);
      }
    }
  /* east :748-776 */
  for (int j = 0; j <= n + 1; ++j)
    for (int k = 0; k <= l + 1; ++k) {
      const size_t c = I3(m, j, k);
      if (east == 0 || east == 1 || (east == 2 && E(m, j, k) < 0.9)) {
        aw[c] = aw[c] + ae[c]; ae[c] = 0.;
      } else if (east == 2) {
        bb[c] = bb[c] + ae[c] * p[I3(m + 1, j, k)];
        zero6(ws, c);
      }
    }
  /* west :779-807 */
  for (int j = 0; j <= n + 1; ++j)
    for (int k = 0; k <= l + 1; ++k) {
      const size_t c = I3(1, j, k);
      if (west == 0 || west == 1 || (west == 2 && E(1, j, k) < 0.9)) {
        ae[c] = ae[c] + aw[c]; aw[c] = 0.;
      } else if (west == 2) {
        bb[c] = bb[c] + aw[c] * p[I3(0, j, k)];
        zero6(ws, c);
      }
    }
  /* north :810-836 */
  for (int i = 0; i <= m + 1; ++i)
    for (int k = 0; k <= l + 1; ++k) {
      const size_t c = I3(i, n, k);
      if (north == 0 || north == 1 || (north == 2 && E(i, n, k) < 0.9)) {
        as[c] = as[c] + an[c]; an[c] = 0.;
      } else if (north == 2) {
        bb[c] = bb[c] + an[c] * p[I3(i, n + 1, k)];
        zero6(ws, c);
      }
    }
  /* south :839-863 */
  for (int i = 0; i <= m + 1; ++i)
    for (int k = 0; k <= l + 1; ++k) {
      const size_t c = I3(i, 1, k);
      if (south == 0 || south == 1 || (south == 2 && E(i, 1, k) < 0.9)) {
        an[c] = an[c] + as[c]; as[c] = 0.;
      } else if (south == 2) {
        bb[c] = bb[c] + as[c] * p[I3(i, 0, k)];
        zero6(ws, c);
      }
    }
}

/* reference's linear-index -> (i,j,k) colour mapping, ibm_3d_uniform_omp_cpu.f90:493-508 (even
 * space, first) and :550-565 (odd space, second). */
static inline void map3(int ii, int m, int n, int second, int *pi, int *pj, int *pk) {
  int k = (ii - 1) / (m * n) + 1;
  int j = ((ii - 1) / m + 1) - (k - 1) * n;
  int i = (ii - (j - 1) * m) - (k - 1) * m * n;
  if ((m % 2) != 0 && (n % 2) == 0 && (k % 2) == 0) {
    if ((i % 2) != 0) j = second ? j + 1 : j - 1;
    else              j = second ? j - 1 : j + 1;
  } else if ((m % 2) == 0 && ((j % 2) + (k % 2) == 1)) {
    i = second ? i + 1 : i - 1;
  }
  *pi = i; *pj = j; *pk = k;
}

static void halo3_p(const pfo_params *P, double *p) {
  const int m = P->m, n = P->n, l = P->l;
#pragma omp parallel for
  for (int i = 1; i <= m; ++i)
    for (int k = 1; k <= l; ++k) {
      p[I3(i, 0, k)] = p[I3(i, n, k)];
      p[I3(i, n + 1, k)] = p[I3(i, 1, k)];
    }
#pragma omp parallel for
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j) {
      p[I3(i, j, 0)] = p[I3(i, j, l)];
      p[I3(i, j, l + 1)] = p[I3(i, j, 1)];
    }
}

static void copy3_pold(const pfo_params *P, const double *p, double *p_old) {
  const int m = P->m, n = P->n, l = P->l;
  /* reference loop order i,j,k (k innermost), :482-490 */
#pragma omp parallel for
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j)
      for (int k = 0; k <= l + 1; ++k) p_old[I3(i, j, k)] = p[I3(i, j, k)];
}

/* solve_matrix_vec_omp: ibm_3d_uniform_omp_cpu.f90:433-614 (periodic=1);
 * air-condition :480-661 has the halo refreshes commented out (periodic=0).
 * Runs `iters` iterations and returns the running-max |p-p_old| of the reference (:575-583). */
PFO_EXPORT double pfo3_sor(const pfo_params *P, int periodic, int iters, double *p, pfo_ws *ws) {
  const int m = P->m, n = P->n, l = P->l;
  const double relux_factor = P->relux_factor;
  const double *ap = ws->ap, *ae = ws->ae, *aw = ws->aw, *an = ws->an, *as = ws->as,
               *at = ws->at, *ab = ws->ab, *bb = ws->bb;
  double *p_old = ws->p_old;
  double error = 0.0;
  const int N = m * n * l;
  for (int iter = 1; iter <= iters; ++iter) {
    for (int second = 0; second <= 1; ++second) {
      if (periodic) halo3_p(P, p);
      copy3_pold(P, p, p_old);
#pragma omp parallel for
      for (int ii = (second ? 1 : 2); ii <= N; ii += 2) {
        int i, j, k;
        map3(ii, m, n, second, &i, &j, &k);
        const size_t c = I3(i, j, k);
        p[c] = (bb[c] - ae[c] * p_old[I3(i + 1, j, k)] - aw[c] * p_old[I3(i - 1, j, k)] -
                an[c] * p_old[I3(i, j + 1, k)] - as[c] * p_old[I3(i, j - 1, k)] -
                at[c] * p_old[I3(i, j, k + 1)] - ab[c] * p_old[I3(i, j, k - 1)]) /
                   ap[c] * relux_factor +
               p_old[c] * (1. - relux_factor);
      }
    }
#pragma omp parallel for reduction(max : error)
    for (int i = 1; i <= m; ++i)
      for (int j = 1; j <= n; ++j)
        for (int k = 1; k <= l; ++k)
          error = dmax(error, fabs(p[I3(i, j, k)] - p_old[I3(i, j, k)]));
  }
  if (periodic) halo3_p(P, p);
  return error;
}

/* projection: ibm_3d_uniform_omp_cpu.f90:110-125 */
PFO_EXPORT void pfo3_project(const pfo_params *P, const double *p, double *u, double *v,
                             double *w) {
  const int m = P->m, n = P->n, l = P->l;
  const double dx = P->dx, dy = P->dy, dz = P->dz, dt = P->dt, density = P->density;
#pragma omp parallel for
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        const size_t c = I3(i, j, k);
        u[c] = u[c] - dt / density * (p[I3(i + 1, j, k)] - p[I3(i - 1, j, k)]) / dx * 0.5;
        v[c] = v[c] - dt / density * (p[I3(i, j + 1, k)] - p[I3(i, j - 1, k)]) / dy * 0.5;
        w[c] = w[c] - dt / density * (p[I3(i, j, k + 1)] - p[I3(i, j, k - 1)]) / dz * 0.5;
      }
}

/* boundary, uniform: ibm_3d_uniform_omp_cpu.f90:669-752 (inlet angle AoA/1300, sic) */
PFO_EXPORT void pfo3u_boundary(const pfo_params *P, double *p, double *u, double *v, double *w) {
  const int m = P->m, n = P->n, l = P->l;
  const double pi = atan(1.) * 4.;
  const double uin = P->inlet_velocity * cos(P->AoA / 1300. * pi);
  const double vin = P->inlet_velocity * sin(P->AoA / 1300. * pi);
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j) {
      u[I3(1, j, k)] = uin;
      v[I3(1, j, k)] = vin;
      w[I3(1, j, k)] = 0.;
      u[I3(0, j, k)] = u[I3(1, j, k)];
      v[I3(0, j, k)] = v[I3(1, j, k)];
      w[I3(0, j, k)] = w[I3(1, j, k)];
      p[I3(0, j, k)] = p[I3(2, j, k)];
    }
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j) {
      u[I3(m + 1, j, k)] = u[I3(m - 1, j, k)];
      v[I3(m + 1, j, k)] = v[I3(m - 1, j, k)];
      w[I3(m + 1, j, k)] = w[I3(m - 1, j, k)];
      p[I3(m + 1, j, k)] = P->outlet_pressure;
    }
  for (int k = 0; k <= l + 1; ++k)
    for (int i = 0; i <= m + 1; ++i) {
      u[I3(i, 0, k)] = u[I3(i, n, k)];
      v[I3(i, 0, k)] = v[I3(i, n, k)];
      w[I3(i, 0, k)] = w[I3(i, n, k)];
      p[I3(i, 0, k)] = p[I3(i, n, k)];
      u[I3(i, n + 1, k)] = u[I3(i, 1, k)];
      v[I3(i, n + 1, k)] = v[I3(i, 1, k)];
      w[I3(i, n + 1, k)] = w[I3(i, 1, k)];
      p[I3(i, n + 1, k)] = p[I3(i, 1, k)];
    }
  for (int j = 0; j <= n + 1; ++j)
    for (int i = 0; i <= m + 1; ++i) {
      u[I3(i, j, 0)] = u[I3(i, j, l)];
      v[I3(i, j, 0)] = v[I3(i, j, l)];
      w[I3(i, j, 0)] = w[I3(i, j, l)];
      p[I3(i, j, 0)] = p[I3(i, j, l)];
      u[I3(i, j, l + 1)] = u[I3(i, j, 1)];
      v[I3(i, j, l + 1)] = v[I3(i, j, 1)];
      w[I3(i, j, l + 1)] = w[I3(i, j, 1)];
      p[I3(i, j, l + 1)] = p[I3(i, j, 1)];
    }
}

/* boundary, air-condition: ibm_3d_air_condition_omp_cpu.f90:873-1170.  Serial, faces in the
 * order top, bottom, west, east, north, south.  One helper per face type would hide the
 * reference's per-face differences (which ghost component is mirrored, the bottom-inlet test
 * on porosity(i,j,l) :948, outlet ghosts copying k=1 vs k=l-1), so each face is spelled out. */
PFO_EXPORT void pfo3a_boundary(const pfo_params *P, const double *porosity, double *p, double *u,
                               double *v, double *w) {
  const int m = P->m, n = P->n, l = P->l;
  const double uin = P->inlet_velocity, pout = P->outlet_pressure;
  const int top = P->wall[PFO_TOP], bottom = P->wall[PFO_BOTTOM], east = P->wall[PFO_EAST],
            west = P->wall[PFO_WEST], south = P->wall[PFO_SOUTH], north = P->wall[PFO_NORTH];
  /* top :890-933 */
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      const int fluid = E(i, j, l) >= 0.9;
      if (top == 1 && fluid) {
        u[I3(i, j, l)] = 0.; v[I3(i, j, l)] = 0.; w[I3(i, j, l)] = -uin;
        u[I3(i, j, l + 1)] = u[I3(i, j, l)];
        v[I3(i, j, l + 1)] = v[I3(i, j, l)];
        w[I3(i, j, l + 1)] = w[I3(i, j, l)];
        p[I3(i, j, l + 1)] = p[I3(i, j, l - 1)];
      } else if (top == 2 && fluid) {
        u[I3(i, j, l + 1)] = u[I3(i, j, l - 1)];
        v[I3(i, j, l + 1)] = v[I3(i, j, l - 1)];
        w[I3(i, j, l + 1)] = w[I3(i, j, l - 1)];
        p[I3(i, j, l + 1)] = pout;
      } else if (top == 0 || top == 1 || top == 2) {
        u[I3(i, j, l)] = 0.; v[I3(i, j, l)] = 0.; w[I3(i, j, l)] = 0.;
        w[I3(i, j, l + 1)] = -w[I3(i, j, l - 1)];
        p[I3(i, j, l + 1)] = p[I3(i, j, l - 1)];
      }
    }
  /* bottom :936-979 (inlet branch tests porosity(i,j,l), sic :948) */
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      if (bottom == 1 && E(i, j, l) >= 0.9) {
        u[I3(i, j, 1)] = 0.; v[I3(i, j, 1)] = 0.; w[I3(i, j, 1)] = uin;
        u[I3(i, j, 0)] = u[I3(i, j, 1)];
        v[I3(i, j, 0)] = v[I3(i, j, 1)];
        w[I3(i, j, 0)] = w[I3(i, j, 1)];
        p[I3(i, j, 0)] = p[I3(i, j, 2)];
      } else if (bottom == 2 && E(i, j, 1) >= 0.9) {
        u[I3(i, j, 0)] = u[I3(i, j, 1)];
        v[I3(i, j, 0)] = v[I3(i, j, 1)];
        w[I3(i, j, 0)] = w[I3(i, j, 1)];
        p[I3(i, j, 0)] = pout;
      } else if (bottom == 0 || bottom == 1 || bottom == 2) {
        u[I3(i, j, 1)] = 0.; v[I3(i, j, 1)] = 0.; w[I3(i, j, 1)] = 0.;
        w[I3(i, j, 0)] = -w[I3(i, j, 2)];
        p[I3(i, j, 0)] = p[I3(i, j, 2)];
      }
    }
  /* west :982-1024 */
  for (int j = 0; j <= n + 1; ++j)
    for (int k = 0; k <= l + 1; ++k) {
      const int fluid = E(1, j, k) >= 0.9;
      if (west == 1 && fluid) {
        u[I3(1, j, k)] = uin; v[I3(1, j, k)] = 0.; w[I3(1, j, k)] = 0.;
        u[I3(0, j, k)] = u[I3(1, j, k)];
        v[I3(0, j, k)] = v[I3(1, j, k)];
        w[I3(0, j, k)] = w[I3(1, j, k)];
        p[I3(0, j, k)] = p[I3(2, j, k)];
      } else if (west == 2 && fluid) {
        u[I3(0, j, k)] = u[I3(1, j, k)];
        v[I3(0, j, k)] = v[I3(1, j, k)];
        w[I3(0, j, k)] = w[I3(1, j, k)];
        p[I3(0, j, k)] = pout;
      } else if (west == 0 || west == 1 || west == 2) {
        u[I3(1, j, k)] = 0.; v[I3(1, j, k)] = 0.; w[I3(1, j, k)] = 0.;
        u[I3(0, j, k)] = -u[I3(2, j, k)];
        p[I3(0, j, k)] = p[I3(2, j, k)];
      }
    }
  /* east :1027-1071 */
  for (int j = 0; j <= n + 1; ++j)
    for (int k = 0; k <= l + 1; ++k) {
      const int fluid = E(m, j, k) >= 0.9;
      if (east == 1 && fluid) {
        u[I3(m, j, k)] = -uin; v[I3(m, j, k)] = 0.; w[I3(m, j, k)] = 0.;
        u[I3(m + 1, j, k)] = u[I3(m, j, k)];
        v[I3(m + 1, j, k)] = v[I3(m, j, k)];
        w[I3(m + 1, j, k)] = w[I3(m, j, k)];
        p[I3(m + 1, j, k)] = p[I3(m - 1, j, k)];
      } else if (east == 2 && fluid) {
        u[I3(m + 1, j, k)] = u[I3(m, j, k)];
        v[I3(m + 1, j, k)] = v[I3(m, j, k)];
        w[I3(m + 1, j, k)] = w[I3(m, j, k)];
        p[I3(m + 1, j, k)] = pout;
      } else if (east == 0 || east == 1 || east == 2) {
        u[I3(m, j, k)] = 0.; v[I3(m, j, k)] = 0.; w[I3(m, j, k)] = 0.;
        u[I3(m + 1, j, k)] = -u[I3(m - 1, j, k)];
        p[I3(m + 1, j, k)] = p[I3(m - 1, j, k)];
      }
    }
  /* north :1074-1116 (wall ghost mirrors u, sic) */
  for (int i = 0; i <= m + 1; ++i)
    for (int k = 0; k <= l + 1; ++k) {
      const int fluid = E(i, n, k) >= 0.9;
      if (north == 1 && fluid) {
        u[I3(i, n, k)] = -uin; v[I3(i, n, k)] = 0.; w[I3(i, n, k)] = 0.;
        u[I3(i, n + 1, k)] = u[I3(i, n, k)];
        v[I3(i, n + 1, k)] = v[I3(i, n, k)];
        w[I3(i, n + 1, k)] = w[I3(i, n, k)];
        p[I3(i, n + 1, k)] = p[I3(i, n - 1, k)];
      } else if (north == 2 && fluid) {
        u[I3(i, n + 1, k)] = u[I3(i, n, k)];
        v[I3(i, n + 1, k)] = v[I3(i, n, k)];
        w[I3(i, n + 1, k)] = w[I3(i, n, k)];
        p[I3(i, n + 1, k)] = pout;
      } else if (north == 0 || north == 1 || north == 2) {
        u[I3(i, n, k)] = 0.; v[I3(i, n, k)] = 0.; w[I3(i, n, k)] = 0.;
        u[I3(i, n + 1, k)] = -u[I3(i, n - 1, k)];
        p[I3(i, n + 1, k)] = p[I3(i, n - 1, k)];
      }
    }
  /* south :1119-1166 (outlet ghosts copy j=2, wall ghost mirrors v) */
  for (int i = 0; i <= m + 1; ++i)
    for (int k = 0; k <= l + 1; ++k) {
      const int fluid = E(i, 1, k) >= 0.9;
      if (south == 1 && fluid) {
        u[I3(i, 1, k)] = uin; v[I3(i, 1, k)] = 0.; w[I3(i, 1, k)] = 0.;
        u[I3(i, 0, k)] = u[I3(i, 1, k)];
        v[I3(i, 0, k)] = v[I3(i, 1, k)];
        w[I3(i, 0, k)] = w[I3(i, 1, k)];
        p[I3(i, 0, k)] = p[I3(i, 2, k)];
      } else if (south == 2 && fluid) {
        u[I3(i, 0, k)] = u[I3(i, 2, k)];
        v[I3(i, 0, k)] = v[I3(i, 2, k)];
        w[I3(i, 0, k)] = w[I3(i, 2, k)];
        p[I3(i, 0, k)] = pout;
      } else if (south == 0 || south == 1 || south == 2) {
        u[I3(i, 1, k)] = 0.; v[I3(i, 1, k)] = 0.; w[I3(i, 1, k)] = 0.;
        v[I3(i, 0, k)] = -v[I3(i, 2, k)];
        p[I3(i, 0, k)] = p[I3(i, 2, k)];
      }
    }
}
#undef E

/* One or more whole time steps, ibm_3d_uniform_omp_cpu.f90:81-132 / air :95-146.
 * p_error[s] receives the per-step `p error` the reference prints (:608). */
PFO_EXPORT void pfo3_step(const pfo_params *P, int air, int nsteps, double *p, double *u,
                          double *v, double *w, double *u_old, double *v_old, double *w_old,
                          const double *porosity, pfo_ws *ws, double *p_error) {
  for (int s = 0; s < nsteps; ++s) {
    pfo3_copy_old(P, u, v, w, u_old, v_old, w_old);
    pfo3_divergence(P, air, u_old, v_old, w_old, ws->div);
    pfo3_predictor(P, u_old, v_old, w_old, porosity, ws->div, u, v, w);
    pfo3_matrix(P, u, v, w, porosity, ws);
    if (air) pfo3a_boundary_matrix(P, p, porosity, ws);
    else     pfo3u_boundary_matrix(P, p, ws);
    const double err = pfo3_sor(P, !air, P->iter_max, p, ws);
    if (p_error) p_error[s] = err;
    pfo3_project(P, p, u, v, w);
    if (air) pfo3a_boundary(P, porosity, p, u, v, w);
    else     pfo3u_boundary(P, p, u, v, w);
  }
}
#undef LX
#undef LY
#undef I3

/* ================================================================================== */
/*                                   2 D                                               */
/* ================================================================================== */
#define LX ((size_t)(P->m + 2))
#define I2(i, j) ((size_t)(i) + LX * (size_t)(j))
#define E(a, b) porosity[I2(a, b)]

/* lib/grid.f90:92-106: x zero-gradient on j=1..n+1, then periodic y on i=0..m+1 */
PFO_EXPORT void pfo2_porosity_halo(const pfo_params *P, double *porosity) {
  const int m = P->m, n = P->n;
  for (int j = 1; j <= n + 1; ++j) {
    porosity[I2(0, j)] = porosity[I2(1, j)];
    porosity[I2(m + 1, j)] = porosity[I2(m, j)];
  }
  for (int i = 0; i <= m + 1; ++i) {
    porosity[I2(i, 0)] = porosity[I2(i, n)];
    porosity[I2(i, n + 1)] = porosity[I2(i, 1)];
  }
}

/* ibm_2d_uniform_omp_cpu.f90:542-568 ; backstep multiplies by porosity ibm_2d_backstep_omp_cpu.f90:615-617 */
PFO_EXPORT void pfo2_initial_conditions(const pfo_params *P, int backstep, const double *porosity,
                                        double *p, double *u, double *v) {
  const int m = P->m, n = P->n;
  const double pai = atan(1.) * 4.;
  for (int j = 1; j <= n; ++j)
    for (int i = 1; i <= m; ++i) {
      if (backstep) {
        u[I2(i, j)] = P->inlet_velocity * cos(P->AoA / 180 * pai) * E(i, j);
        v[I2(i, j)] = P->inlet_velocity * sin(P->AoA / 180 * pai) * E(i, j);
      } else {
        u[I2(i, j)] = P->inlet_velocity * cos(P->AoA / 180 * pai);
        v[I2(i, j)] = P->inlet_velocity * sin(P->AoA / 180 * pai);
      }
      p[I2(i, j)] = P->outlet_pressure;
    }
}

/* ibm_2d_uniform_omp_cpu.f90:84-96 */
PFO_EXPORT void pfo2_copy_old(const pfo_params *P, const double *u, const double *v,
                              double *u_old, double *v_old) {
  const int m = P->m, n = P->n;
#pragma omp parallel for
  for (int i = 0; i <= m + 1; ++i)
    for (int j = 0; j <= n + 1; ++j) {
      u_old[I2(i, j)] = u[I2(i, j)];
      v_old[I2(i, j)] = v[I2(i, j)];
    }
}

/* ibm_2d_uniform_omp_cpu.f90:172-194 (second term divides by dx, sic :176) */
PFO_EXPORT void pfo2_divergence(const pfo_params *P, const double *u_old, const double *v_old,
                                double *div) {
  const int m = P->m, n = P->n;
  const double dx = P->dx;
#pragma omp parallel for
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j)
      div[I2(i, j)] = (u_old[I2(i + 1, j)] - u_old[I2(i - 1, j)]) / dx * .5 +
                      (v_old[I2(i, j + 1)] - v_old[I2(i, j - 1)]) / dx * .5;
  for (int j = 1; j <= n; ++j) {
    div[I2(0, j)] = 0.;
    div[I2(m + 1, j)] = 0.;
  }
  for (int i = 1; i <= m; ++i) {
    div[I2(i, 0)] = div[I2(i, n)];
    div[I2(i, n + 1)] = div[I2(i, 1)];
  }
}

/* ibm_2d_uniform_omp_cpu.f90:200-258.  Note the 2D association of the convection terms,
 * dt*(u*(du)/dx/2.), and the v wall force using dx (:254). */
PFO_EXPORT void pfo2_predictor(const pfo_params *P, const double *u_old, const double *v_old,
                               const double *porosity, const double *div, double *u, double *v) {
  const int m = P->m, n = P->n;
  const double dx = P->dx, dy = P->dy, dt = P->dt;
  const double xnue = P->xnue, xlambda = P->xlambda, thickness = P->thickness;
  const int nonslip = P->nonslip;
#define U(a, b) u_old[I2(a, b)]
#define V(a, b) v_old[I2(a, b)]
#define D(a, b) div[I2(a, b)]
#pragma omp parallel for
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j) {
      double r;
      r = U(i, j) - dt * (U(i, j) * (U(i + 1, j) - U(i - 1, j)) / dx / 2.);
      r = r - dt * (V(i, j) * (U(i, j + 1) - U(i, j - 1)) / dy / 2.);
      r = r + dt * xnue * (U(i + 1, j) - 2. * U(i, j) + U(i - 1, j)) / dx / dx;
      r = r + dt * xnue * (U(i, j + 1) - 2. * U(i, j) + U(i, j - 1)) / dy / dy;
      r = r + dt * (xnue + xlambda) * (D(i + 1, j) - D(i - 1, j)) / dx * .5;
      r = r + dt * (((U(i + 1, j) - U(i - 1, j)) / dx * .5 + (U(i + 1, j) - U(i - 1, j)) / dx * .5) *
                        xnue * (E(i + 1, j) - E(i - 1, j)) / dx * .5 +
                    ((U(i, j + 1) - U(i, j - 1)) / dy * .5 + (V(i + 1, j) - V(i - 1, j)) / dx * .5) *
                        xnue * (E(i, j + 1) - E(i, j - 1)) / dy * .5 +
                    D(i, j) * (E(i + 1, j) - E(i - 1, j)) / dx * 0.5 * xlambda) /
                  E(i, j);
      if (nonslip)
        r = r - dt * xnue * U(i, j) / ((thickness * dx) * (thickness * dx)) * ALPHA * E(i, j) *
                    (1. - E(i, j)) * (1. - E(i, j));
      u[I2(i, j)] = r;
    }
#pragma omp parallel for
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j) {
      double r;
      r = V(i, j) - dt * (U(i, j) * (V(i + 1, j) - V(i - 1, j)) / dx / 2.);
      r = r - dt * (V(i, j) * (V(i, j + 1) - V(i, j - 1)) / dy / 2.);
      r = r + dt * xnue * (V(i + 1, j) - 2. * V(i, j) + V(i - 1, j)) / dx / dx;
      r = r + dt * xnue * (V(i, j + 1) - 2. * V(i, j) + V(i, j - 1)) / dy / dy;
      r = r + dt * (xnue + xlambda) * (D(i, j + 1) - D(i, j - 1)) / dy * .5;
      r = r + dt * (((V(i + 1, j) - V(i - 1, j)) / dx * .5 + (U(i, j + 1) - U(i, j - 1)) / dy * .5) *
                        xnue * (E(i + 1, j) - E(i - 1, j)) / dx * .5 +
                    ((V(i, j + 1) - V(i, j - 1)) / dy * .5 + (V(i, j + 1) - V(i, j - 1)) / dy * .5) *
                        xnue * (E(i, j + 1) - E(i, j - 1)) / dy * .5 +
                    D(i, j) * (E(i, j + 1) - E(i, j - 1)) / dy * 0.5 * xlambda) /
                  E(i, j);
      if (nonslip)
        r = r - dt * xnue * V(i, j) / ((thickness * dx) * (thickness * dx)) * ALPHA * E(i, j) *
                    (1. - E(i, j)) * (1. - E(i, j));
      v[I2(i, j)] = r;
    }
#undef U
#undef V
#undef D
}

/* ibm_2d_uniform_omp_cpu.f90:262-278 */
PFO_EXPORT void pfo2_matrix(const pfo_params *P, const double *u, const double *v,
                            const double *porosity, pfo_ws *ws) {
  const int m = P->m, n = P->n;
  const double dx = P->dx, dy = P->dy, dt = P->dt, density = P->density;
#pragma omp parallel for
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j) {
      const size_t c = I2(i, j);
      ws->ae[c] = dt * dmax(SMALL, (E(i + 1, j) + E(i, j)) * 0.5) / dx / dx;
      ws->aw[c] = dt * dmax(SMALL, (E(i, j) + E(i - 1, j)) * 0.5) / dx / dx;
      ws->an[c] = dt * dmax(SMALL, (E(i, j + 1) + E(i, j)) * 0.5) / dy / dy;
      ws->as[c] = dt * dmax(SMALL, (E(i, j) + E(i, j - 1)) * 0.5) / dy / dy;
      ws->ap[c] = -ws->ae[c] - ws->aw[c] - ws->an[c] - ws->as[c];
      ws->bb[c] = ((E(i + 1, j) * u[c] + E(i, j) * u[I2(i + 1, j)]) * 0.5 -
                   (E(i - 1, j) * u[c] + E(i, j) * u[I2(i - 1, j)]) * 0.5) *
                      density / dx +
                  ((E(i, j + 1) * v[c] + E(i, j) * v[I2(i, j + 1)]) * 0.5 -
                   (E(i, j - 1) * v[c] + E(i, j) * v[I2(i, j - 1)]) * 0.5) *
                      density / dy;
    }
}

/* ibm_2d_uniform_omp_cpu.f90:410-454 */
PFO_EXPORT void pfo2_boundary_matrix(const pfo_params *P, const double *p, pfo_ws *ws) {
  const int m = P->m, n = P->n;
  for (int j = 1; j <= n; ++j) {
    ws->ae[I2(1, j)] = ws->ae[I2(1, j)] + ws->aw[I2(1, j)];
    ws->aw[I2(1, j)] = 0.;
  }
  for (int j = 1; j <= n; ++j) {
    const size_t c = I2(m, j);
    ws->bb[c] = ws->bb[c] + ws->ae[c] * p[I2(m + 1, j)];
    ws->ae[c] = 0.; ws->aw[c] = 0.; ws->an[c] = 0.; ws->as[c] = 0.;
  }
}

/* solve_matrix_vec_omp 2D: ibm_2d_uniform_omp_cpu.f90:293-406.  First colour is the "even
 * space" k=2,m*n,2 == (i+j) odd; error accumulates in BOTH half-sweeps (:351,:385). */
PFO_EXPORT double pfo2_sor(const pfo_params *P, int iters, double *p, pfo_ws *ws) {
  const int m = P->m, n = P->n;
  const double relux_factor = P->relux_factor;
  const double *ap = ws->ap, *ae = ws->ae, *aw = ws->aw, *an = ws->an, *as = ws->as,
               *bb = ws->bb;
  double *p_old = ws->p_old;
  double error = 0.0;
  for (int iter = 1; iter <= iters; ++iter) {
    for (int second = 0; second <= 1; ++second) {
      for (int i = 1; i <= m; ++i) {
        p[I2(i, 0)] = p[I2(i, n)];
        p[I2(i, n + 1)] = p[I2(i, 1)];
      }
#pragma omp parallel for
      for (int i = 0; i <= m + 1; ++i)
        for (int j = 0; j <= n + 1; ++j) p_old[I2(i, j)] = p[I2(i, j)];
#pragma omp parallel for reduction(max : error)
      for (int k = (second ? 1 : 2); k <= m * n; k += 2) {
        int j = (k - 1) / m + 1;
        int i = k - (j - 1) * m;
        if ((m % 2) == 0 && (j % 2) == 0) i = second ? i + 1 : i - 1;
        const size_t c = I2(i, j);
        p[c] = (bb[c] - ae[c] * p_old[I2(i + 1, j)] - aw[c] * p_old[I2(i - 1, j)] -
                an[c] * p_old[I2(i, j + 1)] - as[c] * p_old[I2(i, j - 1)]) /
                   ap[c] * relux_factor +
               p_old[c] * (1. - relux_factor);
        error = dmax(error, fabs(p[c] - p_old[c]));
      }
    }
  }
  for (int i = 1; i <= m; ++i) {
    p[I2(i, 0)] = p[I2(i, n)];
    p[I2(i, n + 1)] = p[I2(i, 1)];
  }
  return error;
}

/* ibm_2d_uniform_omp_cpu.f90:103-115 */
PFO_EXPORT void pfo2_project(const pfo_params *P, const double *p, double *u, double *v) {
  const int m = P->m, n = P->n;
  const double dx = P->dx, dy = P->dy, dt = P->dt, density = P->density;
#pragma omp parallel for
  for (int j = 1; j <= n; ++j)
    for (int i = 1; i <= m; ++i) {
      const size_t c = I2(i, j);
      u[c] = u[c] - dt / density * (p[I2(i + 1, j)] - p[I2(i - 1, j)]) / dx * 0.5;
      v[c] = v[c] - dt / density * (p[I2(i, j + 1)] - p[I2(i, j - 1)]) / dy * 0.5;
    }
}

/* ibm_2d_uniform_omp_cpu.f90:460-538 ; backstep inlet times porosity(1,j), ibm_2d_backstep_omp_cpu.f90:533-534 */
PFO_EXPORT void pfo2_boundary(const pfo_params *P, int backstep, const double *porosity,
                              double *p, double *u, double *v) {
  const int m = P->m, n = P->n;
  const double pai = atan(1.) * 4.;
  for (int j = 1; j <= n; ++j) {
    if (backstep) {
      u[I2(1, j)] = P->inlet_velocity * cos(P->AoA / 180. * pai) * E(1, j);
      v[I2(1, j)] = P->inlet_velocity * sin(P->AoA / 180. * pai) * E(1, j);
    } else {
      u[I2(1, j)] = P->inlet_velocity * cos(P->AoA / 180. * pai);
      v[I2(1, j)] = P->inlet_velocity * sin(P->AoA / 180. * pai);
    }
    u[I2(0, j)] = u[I2(1, j)];
    v[I2(0, j)] = v[I2(1, j)];
    p[I2(0, j)] = p[I2(2, j)];
  }
  for (int j = 1; j <= n; ++j) {
    u[I2(m + 1, j)] = u[I2(m - 1, j)];
    v[I2(m + 1, j)] = v[I2(m - 1, j)];
    p[I2(m + 1, j)] = P->outlet_pressure;
  }
  for (int i = 0; i <= m + 1; ++i) {
    u[I2(i, 0)] = u[I2(i, n)];
    v[I2(i, 0)] = v[I2(i, n)];
    p[I2(i, 0)] = p[I2(i, n)];
    u[I2(i, n + 1)] = u[I2(i, 1)];
    v[I2(i, n + 1)] = v[I2(i, 1)];
    p[I2(i, n + 1)] = p[I2(i, 1)];
  }
}

/* ibm_2d_uniform_omp_cpu.f90:80-125 */
PFO_EXPORT void pfo2_step(const pfo_params *P, int backstep, int nsteps, double *p, double *u,
                          double *v, double *u_old, double *v_old, const double *porosity,
                          pfo_ws *ws, double *p_error) {
  for (int s = 0; s < nsteps; ++s) {
    pfo2_copy_old(P, u, v, u_old, v_old);
    pfo2_divergence(P, u_old, v_old, ws->div);
    pfo2_predictor(P, u_old, v_old, porosity, ws->div, u, v);
    pfo2_matrix(P, u, v, porosity, ws);
    pfo2_boundary_matrix(P, p, ws);
    const double err = pfo2_sor(P, P->iter_max, p, ws);
    if (p_error) p_error[s] = err;
    pfo2_project(P, p, u, v);
    pfo2_boundary(P, backstep, porosity, p, u, v);
  }
}
#undef E
#undef LX
#undef I2

PFO_EXPORT int pfo_sizeof_params(void) { return (int)sizeof(pfo_params); }

/* output_force_log_2d: lib/output.f90:244-305 (called every step by ibm_2d_drag_omp_cpu.f90:121).
 * Serial summation in the reference's loop order (i outer, j inner); out = Fpx, Fpy, Fvx, Fvy, Fx, Fy, Cd, Cl */
#define LX ((size_t)(P->m + 2))
#define I2(i, j) ((size_t)(i) + LX * (size_t)(j))
PFO_EXPORT void pfo2_force_log(const pfo_params *P, double radius, const double *p, const double *u,
                               const double *v, const double *porosity, double *out) {
  const int m = P->m, n = P->n;
  const double dx = P->dx, dy = P->dy, thickness = P->thickness, density = P->density, xnue = P->xnue;
  const double small = 1.e-6, alpha = 32.0;
  double force_px = 0.0, force_vx = 0.0, force_py = 0.0, force_vy = 0.0;
  for (int i = 1; i <= m; ++i)
    for (int j = 1; j <= n; ++j) {
      const double e = porosity[I2(i, j)];
      const double gx = (porosity[I2(i + 1, j)] - porosity[I2(i - 1, j)]) * 0.5;
      const double gy = (porosity[I2(i, j + 1)] - porosity[I2(i, j - 1)]) * 0.5;
      const double normal_abs = sqrt(gx * gx + gy * gy);
      const double nx = gx / dmax(normal_abs, small), ny = gy / dmax(normal_abs, small);
      force_px = force_px + (-dx * dy * p[I2(i, j)] * 2 * e * (1.0 - e) / (thickness * dx) * nx);
      force_py = force_py + (-dx * dy * p[I2(i, j)] * 2 * e * (1.0 - e) / (thickness * dy) * ny);
      const double qx = (e * (1.0 - e)) / (thickness * dx), qy = (e * (1.0 - e)) / (thickness * dy);
      force_vx = force_vx + (+dx * dy * alpha * density * xnue * (qx * qx) * u[I2(i, j)]);
      force_vy = force_vy + (+dx * dy * alpha * density * xnue * (qy * qy) * v[I2(i, j)]);
    }
  out[0] = force_px; out[1] = force_py; out[2] = force_vx; out[3] = force_vy;
  out[4] = force_px + force_vx; out[5] = force_py + force_vy;
  out[6] = out[4] / (density * (P->inlet_velocity * P->inlet_velocity) * radius);
  out[7] = out[5] / (density * (P->inlet_velocity * P->inlet_velocity) * radius);
}
#undef LX
#undef I2

/* output_force_log_3d: lib/output.f90:1090-1165 (defined by the reference, called by none of its programs).
 * Serial summation in the reference's loop order (k outer, j, i inner);
 * out = Fpx, Fpy, Fpz, Fvx, Fvy, Fvz, Fx, Fy, Fz, Cd(x), Cl, Cd(z) */
#define LX ((size_t)(P->m + 2))
#define LY ((size_t)(P->n + 2))
#define I3F(i, j, k) ((size_t)(i) + LX * ((size_t)(j) + LY * (size_t)(k)))
PFO_EXPORT void pfo3_force_log(const pfo_params *P, double radius, const double *p, const double *u, const double *v,
                               const double *w, const double *porosity, double *out) {
  const int m = P->m, n = P->n, l = P->l;
  const double dx = P->dx, dy = P->dy, dz = P->dz, thickness = P->thickness, density = P->density, xnue = P->xnue;
  const double small = 1.e-6, alpha = 32.0;
  double fp[3] = {0., 0., 0.}, fv[3] = {0., 0., 0.};
  for (int k = 1; k <= l; ++k)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= m; ++i) {
        const double e = porosity[I3F(i, j, k)];
        const double gx = (porosity[I3F(i + 1, j, k)] - porosity[I3F(i - 1, j, k)]) * 0.5;
        const double gy = (porosity[I3F(i, j + 1, k)] - porosity[I3F(i, j - 1, k)]) * 0.5;
        const double gz = (porosity[I3F(i, j, k + 1)] - porosity[I3F(i, j, k - 1)]) * 0.5;
        const double normal_abs = sqrt(gx * gx + gy * gy + gz * gz);
        const double nx = gx / dmax(normal_abs, small), ny = gy / dmax(normal_abs, small), nz = gz / dmax(normal_abs, small);
        const double pp = p[I3F(i, j, k)];
        fp[0] = fp[0] + (-dx * dy * dz * pp * 2 * e * (1.0 - e) / (thickness * dx) * nx);
        fp[1] = fp[1] + (-dx * dy * dz * pp * 2 * e * (1.0 - e) / (thickness * dy) * ny);
        fp[2] = fp[2] + (-dx * dy * dz * pp * 2 * e * (1.0 - e) / (thickness * dz) * nz);
        const double qx = (e * (1.0 - e)) / (thickness * dx), qy = (e * (1.0 - e)) / (thickness * dy),
                     qz = (e * (1.0 - e)) / (thickness * dz);
        fv[0] = fv[0] + (+dx * dy * dz * alpha * density * xnue * (qx * qx) * u[I3F(i, j, k)]);
        fv[1] = fv[1] + (+dx * dy * dz * alpha * density * xnue * (qy * qy) * v[I3F(i, j, k)]);
        fv[2] = fv[2] + (+dx * dy * dz * alpha * density * xnue * (qz * qz) * w[I3F(i, j, k)]);
      }
  for (int q = 0; q < 3; ++q) {
    out[q] = fp[q];
    out[3 + q] = fv[q];
    out[6 + q] = fp[q] + fv[q];
    out[9 + q] = out[6 + q] / (density * (P->inlet_velocity * P->inlet_velocity) * radius);
  }
}
#undef LX
#undef LY
#undef I3F

/* ------------------------------------------------------------------------------------------------------
 * voxel -> porosity: tools/voxel2poro/voxel2poro.py:33
 *     porosity = scipy.ndimage.convolve(array_3d, kernel, mode='nearest', cval=1.0)
 * The arithmetic lives in a third-party dependency that is not part of /root/reference: SciPy (1.18.1 in
 * this image), scipy/ndimage/src/ni_filters.c, NI_Correlate.  Its published algorithm, restated: convolve =
 * correlate with the weights reversed along every axis (no origin shift for odd sizes); for each output
 * element a double accumulator sums input*weight over the footprint -- the weights with fabs(w) > DBL_EPSILON
 * -- in C order of the (reversed) weight array, inputs beyond the edges replaced by the nearest edge element; the sum is stored as float32
 * because the input array is float32.  PINNED: tests/golden/voxel2poro.npz holds outputs of the reference's
 * own code (its create_tanh_kernel + that scipy call) produced by tests/golden/make_voxel2poro.py.
 * in/out [n0][n1][n2] C order, w [k0][k1][k2], odd sizes. */
static inline int pfo_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
PFO_EXPORT void pfo_convolve3d_nearest(const float *in, int n0, int n1, int n2, const double *w, int k0, int k1,
                                       int k2, float *out) {
  const int h0 = k0 / 2, h1 = k1 / 2, h2 = k2 / 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int i0 = 0; i0 < n0; ++i0)
    for (int i1 = 0; i1 < n1; ++i1)
      for (int i2 = 0; i2 < n2; ++i2) {
        double tmp = 0.0;
        for (int a0 = 0; a0 < k0; ++a0) {
          const int z = pfo_clampi(i0 + a0 - h0, 0, n0 - 1);
          for (int a1 = 0; a1 < k1; ++a1) {
            const int y = pfo_clampi(i1 + a1 - h1, 0, n1 - 1);
            const float *row = in + ((size_t)z * n1 + y) * n2;
            const double *wr = w + ((size_t)(k0 - 1 - a0) * k1 + (k1 - 1 - a1)) * k2;
            for (int a2 = 0; a2 < k2; ++a2) {
              const double ww = wr[k2 - 1 - a2];
              /* NI_Correlate's footprint: weights with fabs(w) <= DBL_EPSILON are dropped */
              if (fabs(ww) > 2.220446049250313e-16) tmp += (double)row[pfo_clampi(i2 + a2 - h2, 0, n2 - 1)] * ww;
            }
          }
        }
        out[((size_t)i0 * n1 + i1) * n2 + i2] = (float)tmp;
      }
}
