/* abi_double.c -- a TEST DOUBLE of the C ABI (include/pixelflow_gpu.h) on the CPU, backed by the oracle.
 *
 * TEST INFRASTRUCTURE ONLY, and deliberately NOT named or built like the product library: it becomes
 * oracle/_ref/libpf_abi_double.so, is linked only into the translated Fortran driver that
 * tests/test_fortran_driver.py runs, and nothing under pixelflow_b200/ can load it (the product has no CPU path).
 *
 * Purpose: the Fortran drivers (pixelflow_b200/fortran/ibm*_gpu.f90 + pixelflow_gpu_mod.f90) cannot be
 * compiled here (no Fortran compiler) and cannot reach a GPU in the build container.  Translated to C by
 * oracle/f90toc.py it CAN run if something answers its ABI calls.  This double answers the entry points that driver
 * uses — pf_config_init, pf_create, pf_last_error, pf_set_porosity, pf_upload, pf_initial_conditions, pf_step,
 * pf_download, pf_force_log_2d, pf_destroy — with the contract of the header: struct_size guard, host arrays in Fortran order with
 * leading dimensions host_ldx / host_ldy, p_error per step.  Compiled against the REAL header, so a drift between the
 * header and this file is a compile error.  What the test then shows: the driver's statement order, its marshalling
 * of the namelist values into pf_config through the bind(C) type, and its use of the reference's own grid and output
 * routines give the reference's run directory byte for byte.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pixelflow_gpu.h"

/* the oracle's C API (oracle/pf_oracle.c) */
typedef struct pfo_params {
  int m, n, l;
  double dx, dy, dz, dt;
  double xnue, xlambda, density, thickness;
  int nonslip;
  int iter_max;
  double relux_factor;
  double inlet_velocity, outlet_pressure, AoA;
  int wall[6];
} pfo_params;
typedef struct pfo_ws pfo_ws;
pfo_ws *pfo_ws_create(size_t nelem);
void pfo_ws_destroy(pfo_ws *ws);
int pfo_sizeof_params(void);
void pfo3_initial_conditions(const pfo_params *P, int air, double *p, double *u, double *v, double *w);
void pfo3u_boundary(const pfo_params *P, double *p, double *u, double *v, double *w);
void pfo3a_boundary(const pfo_params *P, const double *porosity, double *p, double *u, double *v, double *w);
void pfo3_step(const pfo_params *P, int air, int nsteps, double *p, double *u, double *v, double *w, double *uo,
               double *vo, double *wo, const double *porosity, pfo_ws *ws, double *err);
void pfo2_initial_conditions(const pfo_params *P, int backstep, const double *porosity, double *p, double *u, double *v);
void pfo2_boundary(const pfo_params *P, int backstep, const double *porosity, double *p, double *u, double *v);
void pfo2_step(const pfo_params *P, int backstep, int nsteps, double *p, double *u, double *v, double *uo, double *vo,
               const double *porosity, pfo_ws *ws, double *err);
void pfo2_force_log(const pfo_params *P, double radius, const double *p, const double *u, const double *v,
                    const double *porosity, double *out8);

struct pf_solver {
  pfo_params P;
  int air, d3, backstep;
  size_t ldx, ldy;          /* host leading dimensions */
  size_t nelem;             /* dense (l+2)(n+2)(m+2) */
  double *u, *v, *w, *p, *uo, *vo, *wo, *e;
  pfo_ws *ws;
  char err[256];
};

static char g_create_error[256];

#define PUB __attribute__((visibility("default")))

PUB void pf_config_init(pf_config *c) {
  /* the defaults of the product's pf_config_init (pf_api.cu) */
  memset(c, 0, sizeof *c);
  c->struct_size = (int)sizeof(pf_config);
  c->solver_case = PF_IBM3_UNIFORM;
  c->l = 1;
  c->dx = c->dy = c->dz = c->dt = 1.0;
  c->density = 1.0;
  c->thickness = 1.5;
  c->nonslip = 1;
  c->iter_max = 100;
  c->relux_factor = 1.7;
  c->inlet_velocity = 1.0;
  c->wall[PF_TOP] = 1;
  c->wall[PF_SOUTH] = 2;
  c->device = -1;
  c->nranks = 1;
  c->use_graph = 1;
}

PUB int pf_create(pf_solver **out, const pf_config *cfg) {
  if (!out) { snprintf(g_create_error, sizeof g_create_error, "null out pointer"); return 1; }
  *out = 0;
  if (!cfg || cfg->struct_size != (int)sizeof(pf_config)) {
    snprintf(g_create_error, sizeof g_create_error, "pf_config.struct_size = %d, this ABI expects %d",
             cfg ? cfg->struct_size : -1, (int)sizeof(pf_config));
    return 1;
  }
  const int d3 = cfg->solver_case >= PF_IBM3_UNIFORM;
  if (cfg->solver_case < PF_IBM2_UNIFORM || cfg->solver_case > PF_IBM3_AIRCOND) {
    snprintf(g_create_error, sizeof g_create_error, "unknown solver_case %d", cfg->solver_case);
    return 1;
  }
  if (cfg->nranks != 1 || cfg->m < 2 || cfg->n < 2 || (d3 && cfg->l < 2) || pfo_sizeof_params() != (int)sizeof(pfo_params)) {
    snprintf(g_create_error, sizeof g_create_error, "bad configuration");
    return 1;
  }
  pf_solver *s = calloc(1, sizeof *s);
  pfo_params *P = &s->P;
  P->m = cfg->m; P->n = cfg->n; P->l = d3 ? cfg->l : 1;
  s->d3 = d3;
  s->backstep = cfg->solver_case == PF_IBM2_BACKSTEP;
  P->dx = cfg->dx; P->dy = cfg->dy; P->dz = cfg->dz; P->dt = cfg->dt;
  P->xnue = cfg->xnue; P->xlambda = cfg->xlambda; P->density = cfg->density; P->thickness = cfg->thickness;
  P->nonslip = cfg->nonslip; P->iter_max = cfg->iter_max; P->relux_factor = cfg->relux_factor;
  P->inlet_velocity = cfg->inlet_velocity; P->outlet_pressure = cfg->outlet_pressure; P->AoA = cfg->AoA;
  memcpy(P->wall, cfg->wall, sizeof P->wall);
  s->air = cfg->solver_case == PF_IBM3_AIRCOND;
  s->ldx = cfg->host_ldx ? (size_t)cfg->host_ldx : (size_t)cfg->m + 2;
  s->ldy = cfg->host_ldy ? (size_t)cfg->host_ldy : (size_t)cfg->n + 2;
  s->nelem = (size_t)(d3 ? P->l + 2 : 1) * (size_t)(P->n + 2) * (size_t)(P->m + 2);
  double **arr[] = {&s->u, &s->v, &s->w, &s->p, &s->uo, &s->vo, &s->wo, &s->e};
  for (size_t q = 0; q < sizeof arr / sizeof arr[0]; q++) *arr[q] = calloc(s->nelem, sizeof(double));
  s->ws = pfo_ws_create(s->nelem);
  *out = s;
  return 0;
}

PUB void pf_destroy(pf_solver *s) {
  if (!s) return;
  free(s->u); free(s->v); free(s->w); free(s->p); free(s->uo); free(s->vo); free(s->wo); free(s->e);
  pfo_ws_destroy(s->ws);
  free(s);
}

PUB const char *pf_last_error(const pf_solver *s) { return s ? s->err : g_create_error; }

/* the rank helpers of pf_ranks.cu on one rank: nothing forks, the gather is the download */
PUB int pf_ranks_launch(int nranks, int *rank) { (void)nranks; if (rank) *rank = 0; return 0; }
PUB int pf_ranks_rank(void) { return 0; }
PUB int pf_ranks_count(void) { return 1; }
PUB const void *pf_ranks_unique_id(void) { return 0; }
PUB int pf_ranks_barrier(void) { return 0; }
PUB int pf_ranks_finish(int status) { return status; }
PUB int pf_download(pf_solver *s, double *u, double *v, double *w, double *p);
PUB int pf_gather(pf_solver *s, double *u, double *v, double *w, double *p) { return pf_download(s, u, v, w, p); }

/* host (Fortran order, leading dimensions ldx, ldy) <-> dense [l+2][n+2][m+2] */
static void gather(const pf_solver *s, const double *host, double *dense) {
  const size_t nx = (size_t)s->P.m + 2, ny = (size_t)s->P.n + 2, nz = s->d3 ? (size_t)s->P.l + 2 : 1;
  for (size_t k = 0; k < nz; k++)
    for (size_t j = 0; j < ny; j++)
      memcpy(dense + nx * (j + ny * k), host + s->ldx * (j + s->ldy * k), nx * sizeof(double));
}
static void scatter(const pf_solver *s, const double *dense, double *host) {
  const size_t nx = (size_t)s->P.m + 2, ny = (size_t)s->P.n + 2, nz = s->d3 ? (size_t)s->P.l + 2 : 1;
  for (size_t k = 0; k < nz; k++)
    for (size_t j = 0; j < ny; j++)
      memcpy(host + s->ldx * (j + s->ldy * k), dense + nx * (j + ny * k), nx * sizeof(double));
}

PUB int pf_set_porosity(pf_solver *s, const double *porosity) { gather(s, porosity, s->e); return 0; }

PUB int pf_upload(pf_solver *s, const double *u, const double *v, const double *w, const double *p) {
  gather(s, u, s->u); gather(s, v, s->v); gather(s, p, s->p);
  if (s->d3) gather(s, w, s->w);      /* w ignored in 2D (header) */
  return 0;
}

PUB int pf_download(pf_solver *s, double *u, double *v, double *w, double *p) {
  scatter(s, s->u, u); scatter(s, s->v, v); scatter(s, s->p, p);
  if (s->d3) scatter(s, s->w, w);
  return 0;
}

PUB int pf_initial_conditions(pf_solver *s) {
  if (!s->d3) {
    pfo2_initial_conditions(&s->P, s->backstep, s->e, s->p, s->u, s->v);
    pfo2_boundary(&s->P, s->backstep, s->e, s->p, s->u, s->v);
    return 0;
  }
  pfo3_initial_conditions(&s->P, s->air, s->p, s->u, s->v, s->w);
  if (s->air) pfo3a_boundary(&s->P, s->e, s->p, s->u, s->v, s->w);
  else pfo3u_boundary(&s->P, s->p, s->u, s->v, s->w);
  return 0;
}

PUB int pf_step(pf_solver *s, int nsteps, double *p_error) {
  if (nsteps < 0) { snprintf(s->err, sizeof s->err, "nsteps < 0"); return 1; }
  double *err = calloc((size_t)(nsteps > 0 ? nsteps : 1), sizeof(double));
  if (s->d3) pfo3_step(&s->P, s->air, nsteps, s->p, s->u, s->v, s->w, s->uo, s->vo, s->wo, s->e, s->ws, err);
  else pfo2_step(&s->P, s->backstep, nsteps, s->p, s->u, s->v, s->uo, s->vo, s->e, s->ws, err);
  if (p_error) memcpy(p_error, err, (size_t)nsteps * sizeof(double));
  free(err);
  return 0;
}

PUB int pf_force_log_2d(pf_solver *s, double radius, double *out8) {
  if (s->d3) { snprintf(s->err, sizeof s->err, "pf_force_log_2d on a 3D case"); return 1; }
  pfo2_force_log(&s->P, radius, s->p, s->u, s->v, s->e, out8);
  return 0;
}
