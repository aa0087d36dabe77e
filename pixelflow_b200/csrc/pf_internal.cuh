// pf_internal.cuh -- shared declarations of libpixelflow_gpu.so (sm_100a, fp64, no FMA contraction).
//
// Device data layout (see DESIGN.md "Data layout in HBM"):
//
//  natural arrays  (u v w p u_old v_old w_old porosity div):
//      element (i,j,kl) at  X0 + i + NX*(j + NY*kl),   i=0..m+1, j=0..n+1, kl=0..lz+1 (3D) / kl=0 (2D)
//      X0 = 15 so that the first interior cell i=1 starts a 128-byte line; NX % 16 == 0.
//
//  checkerboard ("split") arrays  (p, bb and the seven Poisson coefficients during SOR):
//      one array per colour c = (i+j+kg)&1, kg = global k.  Row (j,kl) of colour c holds the cells
//      of that row whose i has parity s = (c+j+kg)&1, at column H0 + ih, ih = (i-1)>>1
//      (i = 2*ih+1 for s=1, i = 2*ih+2 for s=0; the x-halo i=0 sits at ih=-1).
//      With this numbering the six neighbours of element ih of a colour-c row are, in the OTHER
//      colour's array:  rows (j+-1,kl), (j,kl+-1) element ih, and row (j,kl) elements
//      {ih-1, ih} (s=1) or {ih, ih+1} (s=0)  -> every stream in a half-sweep is unit stride.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pixelflow_gpu.h"

struct Geo {
  int dim;        // 2 or 3
  int m, n;       // interior cells in x, y
  int lz;         // local interior planes (1 in 2D)
  int l;          // global interior planes
  int koff;       // global k = kl + koff   (0 in 2D / rank 0)
  int kin0;       // first interior local plane index: 1 (3D), 0 (2D)
  int NX, NY, NZ; // natural pitch, rows (n+2), planes (lz+2 | 1)
  int X0;         // column of i=0 in a natural row
  int HX, H0;     // split pitch, column of ih=0 in a split row
  long long plane, hplane;  // NX*NY, HX*NY
  long long nat_elems, split_elems;
};

// A loop-invariant divisor with its correctly rounded reciprocal.  `a / inv` (operator below) returns
// the correctly rounded quotient RN(a/d) -- bit-identical to the IEEE division the reference performs
// -- in 5 FMA-pipe operations instead of the ~25 of a generic fp64 division (Markstein's theorem: if
// r = RN(1/d) and q is a faithful quotient, then RN(q + r*(a - q*d)) == RN(a/d); the first
// correction makes q faithful, the second makes it correctly rounded).  Outside a safe exponent
// range (or when `fast` is 0) it falls back to the IEEE division.
struct Inv {
  double d, r;
  int fast;
};

struct Phys {
  Inv ix, iy, iz;           // dx, dy, dz as divisors
  Inv itx2, ity2, itz2;     // (thickness*dx)**2 etc. as divisors (:270, :322, :375)
  double dtrho;             // dt/density (:118)
  double dx, dy, dz, dt;
  double xnue, xlambda, density, thickness;
  double relux;
  double uin, vin;       // inlet velocity components (host libm cos/sin, like the reference)
  double u0, v0;         // initial velocity components
  double inlet_velocity, outlet_pressure;
  int nonslip;
  int scase;             // enum pf_case
  int wall[6];
};

// pointers to the split (checkerboard) operands of one colour
struct SplitSet {
  double *ap, *ae, *aw, *an, *as, *at, *ab, *bb, *p;
  double *eps;   // porosity in checkerboard layout (SOR variant 2 recomputes the coefficients from it)
};

__host__ __device__ inline long long nat_idx(const Geo &g, int i, int j, int kl) {
  return (long long)g.X0 + i + (long long)g.NX * (j + (long long)g.NY * kl);
}
__host__ __device__ inline long long split_row(const Geo &g, int j, int kl) {
  return (long long)g.H0 + (long long)g.HX * (j + (long long)g.NY * kl);
}

#ifdef __CUDACC__
__device__ __noinline__ static double pf_slow_div(double a, double d) { return a / d; }
__device__ __forceinline__ double operator/(double a, const Inv &b) {
  const double fa = fabs(a);
  if (b.fast && ((fa > 1e-280 && fa < 1e280) || a == 0.0)) {
    double q = a * b.r;
    double e = __fma_rn(-q, b.d, a);
    q = __fma_rn(e, b.r, q);
    e = __fma_rn(-q, b.d, a);
    return __fma_rn(e, b.r, q);
  }
  return pf_slow_div(a, b.d);
}
#endif

#define PF_CUDA_OK(call)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      char _b[512];                                                                        \
      snprintf(_b, sizeof(_b), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
               cudaGetErrorString(_e));                                                    \
      throw std::string(_b);                                                               \
    }                                                                                      \
  } while (0)

// ---- communicator (pf_comm.cu): NCCL through dlopen, only when nranks > 1 ----
struct PfComm;
PfComm *pf_comm_create(int rank, int nranks, const void *unique_id, cudaStream_t stream);
void pf_comm_destroy(PfComm *c);
int pf_comm_get_unique_id(void *out128, std::string &err);
// ring exchange of `count` doubles: send_lo -> prev rank's recv_hi, send_hi -> next rank's recv_lo.
// `wrap` = 1 also exchanges across the periodic seam (rank 0 <-> rank P-1); 0 = open chain.
void pf_comm_exchange(PfComm *c, const double *send_lo, const double *send_hi, double *recv_lo,
                      double *recv_hi, size_t count, int wrap, cudaStream_t on = nullptr);
// plain point-to-point transfers of `count` doubles on the communicator's stream (pf_gather)
void pf_comm_send(PfComm *c, const double *src, size_t count, int to);
void pf_comm_recv(PfComm *c, double *dst, size_t count, int from);
void pf_comm_allreduce_max(PfComm *c, double *dev_value, size_t count);
void pf_comm_allreduce_sum(PfComm *c, double *dev_value, size_t count);
// several exchanges issued between these two calls travel as ONE NCCL group (one launch)
void pf_comm_group_begin(PfComm *c);
void pf_comm_group_end(PfComm *c);

// ---- neighbour slabs mapped into this process (CUDA IPC over NVLink peer access, pf_comm.cu) ----
// The fused SOR kernels store the images of their boundary planes straight into the neighbour ranks' ghost
// planes; `block` is the one allocation that holds this rank's pressure buffers and its two arrival flags.
struct PfPeer {
  void *prev = nullptr, *next = nullptr;   // the neighbours' blocks, addressable from this device
  bool same = false;                       // nranks == 2: prev and next are one mapping
  void *prev_map = nullptr, *next_map = nullptr;   // what cudaIpcOpenMemHandle returned (to close)
};
// Collective over the ring.  Returns nullptr ON EVERY RANK (reason in `why`) if any rank cannot export its
// block or map a neighbour's; the caller then keeps the NCCL transport.
PfPeer *pf_peer_open(PfComm *c, void *block, std::string &why);
void pf_peer_close(PfComm *c, PfPeer *p);   // collective
// "my iteration is done" to both neighbours, then wait for theirs: flags hold the sequence number
void k_slab_barrier(unsigned long long *to_prev, unsigned long long *to_next, const unsigned long long *from_prev,
                    const unsigned long long *from_next, unsigned long long *my_seq, cudaStream_t st);

// ---- kernels (pf_kernels.cu / pf_sor.cu), all launched on `st` ----
struct Fields {
  double *u, *v, *w, *p, *uo, *vo, *wo, *eps, *div;
  // air-condition on z-slabs, where the reference reads the OPPOSITE z face (null on one rank: the local arrays hold
  // those planes): the raw right-hand side of global plane 1, on the rank that owns plane l
  // (ibm_3d_air_condition_omp_cpu.f90:702), and the porosity of global plane l, on the rank that owns plane 1 (:948)
  double *bb1, *eps_top;
};

void k_divergence(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st);
void k_div_halo_y(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st);
void k_plane_copy_interior(const Geo &g, double *a, int kl_dst, int kl_src, cudaStream_t st);
void k_shell_copy(const Geo &g, const double *s0, const double *s1, const double *s2, double *d0, double *d1, double *d2,
                  cudaStream_t st);
void k_predictor(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st);
void k_coefficients(const Geo &g, const Phys &ph, const Fields &f, const SplitSet S[2], cudaStream_t st);
void k_rhs(const Geo &g, const Phys &ph, const Fields &f, const SplitSet S[2], cudaStream_t st);
// out[in-plane index] = the raw right-hand side (before any boundary fold) of local plane kl
void k_raw_rhs_plane(const Geo &g, const Phys &ph, const Fields &f, int kl, double *out, cudaStream_t st);
void k_project(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st);
void k_boundary_local(const Geo &g, const Phys &ph, const Fields &f, int rank, int nranks, cudaStream_t st);
void k_plane_copy_full(const Geo &g, double *a, int kl_dst, int kl_src, cudaStream_t st);
void k_initial(const Geo &g, const Phys &ph, const Fields &f, cudaStream_t st);
void k_nat_to_split(const Geo &g, const double *nat, double *s0, double *s1, cudaStream_t st);
void k_force3d(const Geo &g, const Phys &ph, const Fields &f, double *partial, int blocks, double *out6, cudaStream_t st);
void k_force2d(const Geo &g, const Phys &ph, const Fields &f, double *partial, int blocks, double *out4, cudaStream_t st);
void k_split_to_nat(const Geo &g, const double *s0, const double *s1, double *nat, cudaStream_t st);

// SOR (pf_sor.cu)
// planes swept: k = k0 + kstride*z, z = 0..nplanes-1   (k0=1,kstride=1,nplanes=lz: the whole slab)
void k_sor_sweep(const Geo &g, const Phys &ph, const SplitSet S[2], int colour, int with_error,
                 unsigned long long *err_bits, int variant, cudaStream_t st, int k0 = 1, int kstride = 1,
                 int nplanes = -1, int pdl = 0);
bool pf_sor_stores_y_images(const Geo &g);
void k_sor_halo_y(const Geo &g, double *p0, double *p1, int colour_mask, cudaStream_t st);
void k_sor_halo_z_local(const Geo &g, double *p0, double *p1, int colour_mask, cudaStream_t st);

// SOR variant 7 (pf_sor_persistent.cu, opt-in): the half-sweeps of a whole solve in ONE cooperative launch
bool pf_persistent_applicable(const Geo &g, bool air, int nranks);
// SOR variant 8 (pf_sor_tb2d.cu): T red-black iterations per launch on tiles with a ring 2T deep; 2D, one GPU
bool pf_tb2d_applicable(const Geo &g, int nranks);
void k_sor_tb2d(const Geo &g, const Phys &ph, const SplitSet S[2], double *const alt[2], int iters,
                unsigned long long *err_bits, cudaStream_t st);
void k_sor_persistent(const Geo &g, const Phys &ph, const SplitSet S[2], int iters, unsigned long long *err_bits,
                      cudaStream_t st);

// SOR variant 3 (pf_sor_fused.cu): fused red+black pass on depth-2-ghost checkerboard arrays
struct FusedArrays {
  double *cx[2], *cy[2], *cz[2];   // face coefficients (raw ae / an / at of the owning cell), per colour
  double *bb[2];
  double *p[2][2];                 // ping-pong pressure buffers [buffer][colour]
  int cz_planes;                   // planes per z-chunk (set by k_fused_build_faces)
  int rpt;                         // tile shape selector of the register-prefetch kernel (1: 32x8, 2: 32x16)
  bool enabled;
  bool tma;                        // variant 6: TMA-staged pipeline (pf_sor_tma.cu)
  void *tma_cache;                 // host-side CUtensorMap sets (owned by pf_sor_tma.cu)
  int tma_tA, tma_nzA, tma_nzB;    // z-chunk schedule of the TMA kernel (pf_tma_schedule)
  // Where the kernels store the images of the planes next to the slab faces (planes 1,2 -> *_lo, planes
  // lz-1,lz -> *_hi), per [buffer][colour], and the element offset added to the cell's own index:
  //   one rank   : the same array, +-lz planes (the periodic wrap)
  //   slab, P2P  : the neighbour rank's array, mapped over NVLink (prev: +lz_prev planes, next: -lz planes)
  //   slab, NCCL : nullptr (the planes are exchanged after the launch)
  double *img_lo[2][2], *img_hi[2][2];
  long long dk_lo, dk_hi;
  bool slab;                       // this rank owns a z-slab of a larger domain
  // in-kernel neighbour handshake of the TMA kernel on slab ranks with the peer-store transport (pf_sor_tma.cu):
  // this rank's flag words, and the words of the previous / next rank it publishes into (null otherwise)
  unsigned long long *sync, *sync_to_prev, *sync_to_next;
};
// flag words of the in-kernel handshake inside a rank's peer-visible block (see pf_sor_tma.cu)
constexpr int PF_SY_FROM_PREV = 8, PF_SY_FROM_NEXT = 9;
bool pf_tma_applicable(const Geo &g, const Phys &ph, int nranks);
void pf_tma_schedule(const Geo &g, FusedArrays &A);
void pf_tma_release(FusedArrays &A);   // frees the host-side tensor-map cache
void k_tma_iteration(const Geo &g, const Phys &ph, FusedArrays &A, int in, unsigned long long *err_bits,
                     cudaStream_t st);
// multiprocessors of the current device (cached per device) -- never a hard-coded 148
int pf_sm_count();
// planes per z-chunk of a z-streaming tile kernel: `tiles` blocks per chunk, each taking (cz + 2) z-steps (two
// redundant red planes), list-scheduled on `slots` resident blocks; even splits of lz only, chunks of >= 8 planes
int pf_chunk_planes(int lz, long long tiles, int slots);
bool pf_fused_applicable(const Geo &g, const Phys &ph, int nranks);
long long pf_fused_elems(const Geo &g);
void k_fused_build_faces(const Geo &g, const Phys &ph, const double *eps_nat, FusedArrays &A, cudaStream_t st);
// one rank: natural pressure -> both split2 buffers / the final split2 buffer -> natural (pf_sor_fused.cu)
void k_fused_gather_nat(const Geo &g, const FusedArrays &A, const double *nat, cudaStream_t st);
void k_fused_scatter_nat(const Geo &g, const FusedArrays &A, int fin, double *nat, cudaStream_t st);
void k_fused_gather(const Geo &g, const FusedArrays &A, const double *s0, const double *s1, double *d0, double *d1,
                    cudaStream_t st);
void k_fused_scatter(const Geo &g, const FusedArrays &A, const double *s0, const double *s1, double *d0, double *d1,
                     cudaStream_t st);
void k_fused_iteration(const Geo &g, const Phys &ph, const FusedArrays &A, int in, unsigned long long *err_bits,
                       cudaStream_t st);

// ASCII VTK section bodies (pf_output.cu)
bool pf_vtk_section_valid(const Geo &g, int section);
size_t pf_vtk_record_bytes(int section);
void k_vtk_section(const Geo &g, const Fields &f, int section, int k0, int nplanes, const double *xp_dev,
                   const double *yp_dev, const double *zp_dev, double inlet_velocity, char *out, cudaStream_t st);

// message returned by pf_last_error(NULL): failures of calls that have no solver handle
void pf_set_global_error(const std::string &e);

long long pf_launch_count();
void pf_launch_count_reset();
void pf_count_launch();
