/*
 * pixelflow_gpu.h -- C ABI of the B200-native PixelFlow hot path (libpixelflow_gpu.so).
 *
 * Drop-in boundary for the per-timestep hot path of nobu-n2002/PixelFlow's ibm2 / ibm3
 * solvers.  The reference has no plugin API; the seam is the body of `program main`'s time
 * loop (src/omp_parallel/ibm_3d_uniform_omp_cpu.f90:81-132 and the same lines of the other four
 * programs):
 *
 *     u_old = u ...                                   (:85-100)
 *     call solve_p(p,u,v,w,u_old,v_old,w_old,porosity, xnue,xlambda,density,height,thickness,
 *                  yp,dx,dy,dz,dt,m,n,l, nonslip,iter_max,relux_factor)      (:104-107, :151-165)
 *     u = u - dt/density * grad p                     (:110-125)
 *     call boundary(p,u,v,w,xp,yp,zp,width,height,depth,inlet_velocity,outlet_pressure,AoA,
 *                   porosity,m,n,l)                                          (:127-128, :669-679)
 *
 * Every entry point below replaces one of those pieces; the Fortran driver keeps its namelists,
 * porosity CSV, logs and output files and calls these through iso_c_binding
 * (pixelflow_b200/fortran/pixelflow_gpu_mod.f90).
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success, non-zero on error;
 *     the message is available from pf_last_error().  No exceptions cross the ABI.
 *   - arithmetic is fp64, evaluated in the reference's statement order with no FMA contraction.
 *   - host arrays are the Fortran arrays themselves: `real(8), dimension(0:md,0:nd,0:ld)`,
 *     column-major, element (i,j,k) at  i + host_ldx*(j + host_ldy*k),  host_ldx = md+1,
 *     host_ldy = nd+1 (0 in pf_config means dense: m+2, n+2).  2D arrays are (0:md,0:nd).
 *     The library never keeps a host pointer after a call returns and never frees one.
 *   - one solver handle drives one GPU from one host thread (not re-entrant per handle).
 *     Multi-GPU: one process (or thread) per GPU, each with its own handle, rank r owning the
 *     z-slab  k in [k_first, k_first+k_count)  of the global grid (pf_local_slab).
 *   - there is NO CPU fallback: if no CUDA device is usable pf_create fails.
 */
#ifndef PIXELFLOW_GPU_H
#define PIXELFLOW_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define PF_ABI_VERSION 2

/* which of the five reference programs' numerics to run */
enum pf_case {
  PF_IBM2_UNIFORM  = 0, /* ibm_2d_uniform_omp_cpu.f90                                   */
  PF_IBM2_BACKSTEP = 1, /* ibm_2d_backstep_omp_cpu.f90 (inlet/initial velocity * porosity) */
  PF_IBM2_DRAG     = 2, /* ibm_2d_drag_omp_cpu.f90 (hot path identical to uniform)       */
  PF_IBM3_UNIFORM  = 3, /* ibm_3d_uniform_omp_cpu.f90                                    */
  PF_IBM3_AIRCOND  = 4  /* ibm_3d_air_condition_omp_cpu.f90                              */
};

/* indices into pf_config.wall (module wall_conditions, ibm_3d_air_condition_omp_cpu.f90:4-16) */
enum pf_face { PF_TOP = 0, PF_BOTTOM = 1, PF_EAST = 2, PF_WEST = 3, PF_SOUTH = 4, PF_NORTH = 5 };

/* named device arrays for pf_get_field / pf_set_field */
enum pf_field {
  PF_F_U = 0, PF_F_V, PF_F_W, PF_F_P, PF_F_UOLD, PF_F_VOLD, PF_F_WOLD, PF_F_POROSITY, PF_F_DIV,
  PF_F_AP, PF_F_AE, PF_F_AW, PF_F_AN, PF_F_AS, PF_F_AT, PF_F_AB, PF_F_BB,
  PF_F_COUNT
};

typedef struct pf_solver pf_solver; /* opaque */

typedef struct pf_config {
  int struct_size;   /* = sizeof(pf_config); ABI guard                                          */
  int solver_case;   /* enum pf_case                                                            */
  int m, n, l;       /* GLOBAL interior cells (l = 1 for the 2D cases)                          */
  int host_ldx;      /* leading dimension of the host arrays in x (md+1); 0 = m+2               */
  int host_ldy;      /* leading dimension in y (nd+1); 0 = n+2                                  */
  int host_is_slab;  /* 0: host arrays have the global shape; 1: they hold only this rank's slab
                        planes k_first-1 .. k_first+k_count (k_count+2 planes)                  */
  double dx, dy, dz, dt;                    /* lib/grid.f90:297-300                              */
  double xnue, xlambda, density, thickness; /* &physical / &porosity_control                     */
  int nonslip;                              /* &calculation_method (logical as int)              */
  int iter_max;                             /* &solver_control                                   */
  double relux_factor;
  double inlet_velocity, outlet_pressure, AoA;
  int wall[6];       /* air-condition only: 0 wall, 1 inlet, 2 outlet (enum pf_face order)       */
  /* --- device / multi-GPU --- */
  int device;        /* CUDA device ordinal; -1 = keep the calling thread's current device      */
  int rank, nranks;  /* z-slab decomposition; nranks = 1 for a single GPU                       */
  const void *nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks (pf_comm_unique_id);
                                 NULL when nranks == 1                                          */
  /* --- tuning --- */
  int sor_variant;   /* 0 = auto; 1 half-sweeps, 2 coefficients-from-porosity, 3/4 fused red+black
                        (register prefetch), 6 fused red+black (TMA pipeline), 5 = 1 with in-order halo exchange; 7 = the
                        half-sweeps of a whole solve in one cooperative launch (2D cases, 3D air-condition,
                        one GPU); 8 = temporally blocked iterations (2D cases, one GPU: several red-black
                        iterations per launch on shared-memory tiles with a recomputed ring).  A variant that does not apply to the case is replaced by the one that
                        does, and pf_get_sor_variant() reports the kernel that runs; see DESIGN.md section 4 */
  int use_graph;     /* 1 = replay the SOR solve from a CUDA graph (default), 0 = direct launches;
                        -1 = auto                                                               */
  int halo_transport;/* z-slab ranks running a fused SOR kernel (variants 3/4/6): how the planes next to
                        the slab faces reach the neighbour ranks.  0 = auto (2 when every rank can map
                        its neighbours, else 1); 1 = one NCCL send/recv group per iteration; 2 = the
                        kernel stores them into the neighbours' ghost planes over NVLink (CUDA IPC
                        peer mapping) and the ranks meet at a one-thread flag-barrier kernel after
                        every launch; the launches and barriers of a solve replay from one CUDA
                        graph; 3 = as 2, but the TMA kernel (variant 6) meets its neighbours itself
                        (flag words published by the blocks of the boundary z-chunks; pays only
                        when a slab has interior z-chunks to overlap with, kept as an option) --
                        pf_create fails for 2 and 3 if the mapping is not possible              */
} pf_config;

/* ---- lifetime ------------------------------------------------------------------------ */
int  pf_abi_version(void);
/* fills cfg with defaults (struct_size, device=-1, nranks=1, thickness=1.5, iter_max=100, ...) */
void pf_config_init(pf_config *cfg);
int  pf_create(pf_solver **out, const pf_config *cfg);
void pf_destroy(pf_solver *s);
/* message of the last failed call on this handle (s == NULL: last pf_create failure) */
const char *pf_last_error(const pf_solver *s);
/* rank 0 obtains the 128-byte id that every rank then passes in pf_config.nccl_unique_id */
int  pf_comm_unique_id(void *out128);
/* global planes owned by this rank: k = *k_first .. *k_first + *k_count - 1   (1-based) */
int  pf_local_slab(const pf_solver *s, int *k_first, int *k_count);

/* ---- several GPUs from ONE driver program, no launcher (pf_ranks.cu) ---- */
/* The reference's `program main` is started once (src/omp_parallel/ibm_3d_uniform_omp_cpu.f90:4).  Call this first
 * -- before anything touches CUDA -- and the process forks: on return `nranks` copies of the program are running,
 * *rank = 0 .. nranks-1 (rank 0 is the original process and keeps stdout, i.e. the log; the others' stdout is
 * discarded).  Each copy then fills pf_config.rank / nranks / device / nccl_unique_id and continues as on one GPU,
 * with global-shaped host arrays (every copy holds the deck).  nranks = 0: the count is taken from the environment
 * variable PIXELFLOW_GPUS (the analogue of OMP_NUM_THREADS in the reference's config/omp_config.conf; default 1).
 * One rank: nothing happens, *rank = 0. */
int  pf_ranks_launch(int nranks, int *rank);
int  pf_ranks_rank(void);
int  pf_ranks_count(void);
/* the 128-byte NCCL id of this run, for pf_config.nccl_unique_id: made by rank 0 on first use, awaited by the
 * others.  NULL on a single rank or on failure (pf_last_error(NULL)). */
const void *pf_ranks_unique_id(void);
/* every rank arrives (0), or some rank has failed (non-zero, on every rank still alive) */
int  pf_ranks_barrier(void);
/* end of the run: ranks > 0 leave the process with `status` and never return; rank 0 waits for them and returns 0
 * only if every rank finished with status 0 */
int  pf_ranks_finish(int status);

/* ---- data movement ------------------------------------------------------------------- */
/* porosity incl. halos as produced by lib/grid.f90 (grid_conditions*).  Also builds the
 * time-invariant Poisson coefficients ae..ap (ibm_3d_uniform_omp_cpu.f90:390-402 + boundrary_matrix). */
int  pf_set_porosity(pf_solver *s, const double *porosity);
/* u, v, w, p with halos (w ignored / may be NULL in 2D) */
int  pf_upload(pf_solver *s, const double *u, const double *v, const double *w, const double *p);
int  pf_download(pf_solver *s, double *u, double *v, double *w, double *p);
/* Collective over the z-slab ranks: afterwards rank 0's host arrays -- GLOBAL shape (0:md,0:nd,0:ld), whatever
 * host_is_slab says -- hold the whole fields, as the reference's output routines (lib/output.f90) expect them.
 * The other ranks send their planes to rank 0's GPU; their host pointers are ignored and may be NULL.  On one
 * rank this is pf_download. */
int  pf_gather(pf_solver *s, double *u, double *v, double *w, double *p);
/* any named device array, converted to the host layout (tests, output paths) */
int  pf_get_field(pf_solver *s, int field, double *host);
int  pf_set_field(pf_solver *s, int field, const double *host);

/* ---- the hot path -------------------------------------------------------------------- */
/* nsteps whole time steps (:81-132).  p_error[s] receives the `p error` the reference prints
 * after each solve (:608); may be NULL.  Asynchronous work is complete on return. */
int  pf_step(pf_solver *s, int nsteps, double *p_error);
/* same, but through HOST arrays every call: upload u,v,w,p -> nsteps -> download (what a Fortran
 * driver does when it needs the fields on the host after every step). */
int  pf_step_host(pf_solver *s, int nsteps, double *u, double *v, double *w, double *p,
                  double *p_error);

/* fine-grained entry points mirroring the reference's phases (parity tests, custom drivers) */
int  pf_initial_conditions(pf_solver *s); /* initial_conditions + boundary (:68-71)              */
int  pf_copy_old(pf_solver *s);           /* u_old = u ...            (:85-100)                  */
int  pf_divergence(pf_solver *s);         /* div + its halos          (:185-222)                 */
int  pf_predictor(pf_solver *s);          /* u*, v*, w*               (:228-381)                 */
int  pf_build_poisson(pf_solver *s);      /* bb (+ boundary fold)     (:404-409, :650)           */
int  pf_sor(pf_solver *s, int iters, double *p_error); /* solve_matrix_vec_omp (:433-614)        */
int  pf_project(pf_solver *s);            /* velocity correction      (:110-125)                 */
int  pf_boundary(pf_solver *s);           /* boundary                 (:669-752)                 */

/* drag / lift log of ibm2_drag: output_force_log_2d (lib/output.f90:244-305), which that program calls
 * after every step (ibm_2d_drag_omp_cpu.f90:121).  out8 = Fp_x, Fp_y, Fv_x, Fv_y, F_x, F_y, Cd, Cl.
 * 2D cases only.  The four sums are deterministic but ordered differently from the reference's serial
 * loop: equal to rounding, not bit for bit. */
int  pf_force_log_2d(pf_solver *s, double radius, double *out8);
/* output_force_log_3d (lib/output.f90:1090-1165; defined by the reference, called by none of its programs).
 * out12 = Fp_x, Fp_y, Fp_z, Fv_x, Fv_y, Fv_z, F_x, F_y, F_z, Cd(x), Cl, Cd(z).  3D cases; on z-slab ranks the
 * sums are reduced over all ranks (every rank must call).  Equal to the serial reference to rounding. */
int  pf_force_log_3d(pf_solver *s, double radius, double *out12);

/* ---- input: porosity CSV records parsed on the GPU ------------------------ */
/* `text` = the records of a porosity file AFTER its header line (`m,n,l`), nbytes of them: one
 * `index_x, index_y, index_z, porosity_value` record per line, as lib/grid.f90:281-294 (3D) / :38-47 (2D)
 * reads with `read(52,*) x, y, z, poro_val`.  Stores porosity(x,y,z) = max(value, threshold) into the host
 * array porosity[(m+2)*(n+2)*(l+2)] (Fortran order, l = 0 for a 2D file: (m+2)*(n+2), z ignored); cells
 * without a record keep their value; halos are the caller's job (lib/grid.f90).  Every value is the
 * correctly rounded double of its decimal text.  As the reference's loop: only the first m*n*l non-blank lines are
 * read (the rest of the file is never looked at), indices 0 .. m+1 are legal (the arrays are (0:md,...)), and of
 * several records for one cell the last one read wins.  *nrecords = records stored.  Fails on malformed records
 * or indices outside the array among those lines (message: pf_last_error(NULL)). */
int  pf_parse_porosity_csv(const char *text, size_t nbytes, int m, int n, int l, double threshold,
                           double *porosity, long long *nrecords, int device);

/* ---- output: bodies of the ASCII VTK snapshots ----------------------------- */
/* Sections of output_paraview_temp_3d / _2d (lib/output.f90:968-1088 / :421-537), in file order.
 * DIMLESS_V and ABS_DIMLESS_V exist in the 2D files only. */
enum pf_vtk_section {
  PF_VTK_POINTS = 0,            /* xp(i), yp(j), zp(k) | 0.0                     vector records */
  PF_VTK_VELOCITY = 1,          /* u, v, w | 0.0                                                 */
  PF_VTK_VELOCITY_IN_FLUID = 2, /* u*porosity, ...                                               */
  PF_VTK_DIMLESS_V = 3,         /* u*porosity/inlet_velocity, ... (2D)                           */
  PF_VTK_POROSITY = 4,          /*                                               scalar records */
  PF_VTK_PRESSURE = 5,
  PF_VTK_DIVERGENT = 6,         /* (u(i+1)-u(i-1))/(xp(i+1)-xp(i-1)) + ...                       */
  PF_VTK_ABS_DIMLESS_V = 7      /* sqrt((u*porosity/Uin)**2 + (v*porosity/Uin)**2) (2D)          */
};
/* bytes of one section body for `nplanes` planes (3D; the argument is ignored in 2D): m*n*nplanes records of
 * 51 bytes ("(3(f16.4,1x))" with three items: 3 x 16 columns, two blanks, newline) or 17 bytes (one item). */
size_t pf_vtk_section_bytes(const pf_solver *s, int section, int nplanes);
/* Formats the records of `section` for the interior points of the local planes k_local0 .. k_local0+nplanes-1
 * (1-based; 2D: pass 0, 1) on the GPU from the device-resident fields, in the reference's loop order, and
 * copies the text to `out` (host, pf_vtk_section_bytes bytes, no terminator).  xp[0..m+1], yp[0..n+1] and
 * zp[0..l+1] (global k; NULL in 2D) are the driver's grid coordinates (lib/grid.f90).  The driver writes the
 * header lines itself and fwrite()s the bodies. */
int  pf_vtk_section(pf_solver *s, int section, int k_local0, int nplanes, const double *xp, const double *yp,
                    const double *zp, char *out);

/* ---- input preparation: voxel model -> porosity -------------------------- */
/* scipy.ndimage.convolve(in, weights, mode='nearest') as tools/voxel2poro/voxel2poro.py:33 calls it:
 * in/out float32 [n0][n1][n2] (C order, the numpy array_3d), weights float64 [k0][k1][k2] with odd sizes
 * (create_tanh_kernel, voxel2poro.py:189-197: 1 - tanh(r/thickness), normalised).  Double accumulation in
 * scipy's tap order, result rounded to float32: the same bits as the reference's output.  Host pointers;
 * device = CUDA ordinal or -1 for the current one.  Errors: pf_last_error(NULL). */
int  pf_convolve3d_nearest(const float *in, int n0, int n1, int n2, const double *weights, int k0, int k1,
                           int k2, float *out, int device);

/* ---- input preparation: STL surface -> signed distance -------------------- */
/* The numerical core of tools/stl2poro/stl2poro.py (calculate_sdf, :71-84: vtkImplicitPolyDataDistance at every cell
 * centre): signed distance from `npoints` points (points[npoints][3]) to the triangle mesh tri[ntri][3][3] (float32,
 * the vertex triples of a binary STL), negative inside.  Exact point-triangle distance over all triangles; the sign
 * comes from the pseudo-normal of the closest feature (face / edge / vertex); vertices are merged by exact equality.
 * Host pointers; device = CUDA ordinal or -1 for the current one.  Errors: pf_last_error(NULL).  The porosity is
 * 0.5*tanh(d/(thickness*pitch)) + 0.5 (stl2poro.py:87-97): pixelflow_b200/stl2poro.py. */
int  pf_stl_signed_distance(const float *tri, long long ntri, const double *points, long long npoints,
                            double *dist, int device);

/* ---- measurement hooks --------------------------------------------------------------- */
int  pf_sync(pf_solver *s);
/* device-side timings of the last pf_step call, in milliseconds (CUDA events on the solver's
 * stream): total, time inside the SOR solves, and kernel launches issued. */
int  pf_last_timing(const pf_solver *s, double *ms_total, double *ms_sor, long long *launches);
/* the SOR kernel that RUNS: pf_config.sor_variant after auto-selection and after replacing a requested variant
 * that does not apply to this case (e.g. 6 on odd n -> 1); see DESIGN.md section 4 */
int  pf_get_sor_variant(const pf_solver *s);
/* slab-face transport of the fused SOR kernels in use: 0 = single rank (none), 1 = NCCL, 2 = peer stores */
int  pf_get_halo_transport(const pf_solver *s);
/* the CUDA stream (cudaStream_t) the solver launches on, for external event timing */
void *pf_stream(const pf_solver *s);
/* self-check: number of random inputs a (n of them) for which the kernels' exact reciprocal
 * division by the loop-invariant divisor d differs from the IEEE quotient a/d.  Must be 0. */
int  pf_debug_fastdiv_mismatches(double d, long long n, unsigned long long seed, long long *mismatches);
/* self-check of the branch-free division of the fused SOR kernel (variant 6): n random operand pairs with binary exponents in
 * [-exp_range, exp_range] (zeros, denormals, infinities, NaNs mixed in).  *mismatches = pairs inside the guard for
 * which the straight-line sequence differs from the IEEE quotient r/d -- must be 0; *outside (may be NULL) = pairs
 * the guard hands to the plain division. */
int  pf_debug_quot_mismatches(long long n, unsigned long long seed, int exp_range, long long *mismatches,
                              long long *outside);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PIXELFLOW_GPU_H */
