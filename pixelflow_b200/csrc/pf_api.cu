// pf_api.cu -- the C ABI of libpixelflow_gpu.so (include/pixelflow_gpu.h) and the per-step schedule.
//
// One pf_solver owns the device mirrors of one z-slab of the reference's arrays and replays the
// body of `program main`'s time loop (src/omp_parallel/ibm_3d_uniform_omp_cpu.f90:81-132 and the
// same lines of the other four programs) on one CUDA stream.  No CPU fallback exists: every phase
// is a kernel launch, and pf_create fails if no device is usable.
#include <math.h>
#include <cmath>
#include <string.h>

#include <algorithm>

#include "pf_internal.cuh"

struct pf_solver {
  pf_config cfg;
  Geo g;
  Phys ph;
  int rank = 0, nranks = 1;
  bool air = false, uniform3 = false;
  cudaStream_t st = nullptr;
  Fields f{};
  double *tmp = nullptr;  // natural-layout scratch (pf_get_field of checkerboard arrays)
  SplitSet S[2]{};
  bool opp_top = false, opp_bottom = false;   // air-condition slabs: top outlet / bottom inlet read the opposite z face
  double *p_alt[2] = {nullptr, nullptr};   // second pair of checkerboard p arrays (SOR variant 8 ping-pongs)
  FusedArrays fused{};
  double *force_scratch = nullptr;
  double *coord_dev = nullptr;         // xp, yp, zp of the last pf_vtk_section call
  char *text_dev = nullptr;            // formatted section body
  size_t text_cap = 0;
  std::vector<void *> allocs;
  unsigned long long *err_bits = nullptr;
  double *errs_dev = nullptr;
  int errs_cap = 0;
  PfComm *comm = nullptr;
  PfPeer *peer = nullptr;              // neighbour slabs mapped over NVLink (fused SOR kernels on z-slab ranks)
  unsigned long long *flags = nullptr; // [0] written by the previous rank, [1] by the next one
  std::string peer_why;                // why the NCCL transport is in use instead
  cudaStream_t comm_st = nullptr;      // high-priority stream for halo exchanges that overlap the interior sweep
  cudaEvent_t ev_edge = nullptr, ev_comm = nullptr;
  cudaGraphExec_t sor_graph = nullptr;
  int sor_graph_iters = -1;
  long long sor_graph_nodes = 0;
  std::vector<cudaEvent_t> events;
  double ms_total = 0, ms_sor = 0;
  long long launches = 0;
  bool porosity_set = false;
  int device = -1;                     // the CUDA device this solver lives on; made current by every entry point
  int host_ldx = 0, host_ldy = 0;
  std::string err;
};

static thread_local std::string g_create_error;   // failures of calls without a handle, per calling thread
void pf_set_global_error(const std::string &e) { g_create_error = e; }

namespace {

int round_up(int v, int a) { return (v + a - 1) / a * a; }

double *dalloc(pf_solver *s, long long elems) {
  void *p = nullptr;
  PF_CUDA_OK(cudaMalloc(&p, (size_t)elems * sizeof(double)));
  PF_CUDA_OK(cudaMemsetAsync(p, 0, (size_t)elems * sizeof(double), s->st));
  s->allocs.push_back(p);
  return static_cast<double *>(p);
}

// host <-> device transfer of `nplanes` planes starting at local plane kl0 / host plane hk0.
// Host element (i,j,k) lives at i + ldx*(j + ldy*k)  (the Fortran array itself).
void xfer(pf_solver *s, double *dev, const double *host, bool h2d, int kl0, int hk0, int nplanes) {
  const Geo &g = s->g;
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  double *hbase = const_cast<double *>(host) + (size_t)s->host_ldx * s->host_ldy * (size_t)hk0;
  double *dbase = dev + g.X0 + g.plane * kl0;
  cudaPitchedPtr hp = make_cudaPitchedPtr(hbase, (size_t)s->host_ldx * 8, s->host_ldx, s->host_ldy);
  cudaPitchedPtr dp = make_cudaPitchedPtr(dbase, (size_t)g.NX * 8, g.NX, g.NY);
  p.srcPtr = h2d ? hp : dp;
  p.dstPtr = h2d ? dp : hp;
  p.extent = make_cudaExtent((size_t)(g.m + 2) * 8, g.n + 2, nplanes);
  p.kind = h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  PF_CUDA_OK(cudaMemcpy3DAsync(&p, s->st));
}

void upload_field(pf_solver *s, double *dev, const double *host) {
  if (!host) throw std::string("null host array");
  const Geo &g = s->g;
  if (g.dim == 2) { xfer(s, dev, host, true, 0, 0, 1); return; }
  const int hk0 = s->cfg.host_is_slab ? 0 : g.koff;
  xfer(s, dev, host, true, 0, hk0, g.lz + 2);
}

void download_field(pf_solver *s, const double *dev, double *host) {
  if (!host) throw std::string("null host array");
  const Geo &g = s->g;
  double *d = const_cast<double *>(dev);
  if (g.dim == 2) { xfer(s, d, host, false, 0, 0, 1); return; }
  if (s->cfg.host_is_slab) { xfer(s, d, host, false, 0, 0, g.lz + 2); return; }
  // global-shaped host array: own planes, plus the global ghost planes on the end ranks
  const int k0 = (s->rank == 0) ? 0 : 1;
  const int k1 = (s->rank == s->nranks - 1) ? g.lz + 1 : g.lz;
  xfer(s, d, host, false, k0, g.koff + k0, k1 - k0 + 1);
}

// ring / chain exchange of whole planes of a natural-layout array
void exchange_nat(pf_solver *s, double *a, int wrap) {
  const Geo &g = s->g;
  pf_comm_exchange(s->comm, a + g.plane * 1, a + g.plane * g.lz, a, a + g.plane * (g.lz + 1),
                   (size_t)g.plane, wrap);
}

// ring exchange of the two planes next to each slab face of a depth-2-ghost checkerboard array: own planes
// 1,2 -> the previous rank's ghost planes lz+1,lz+2 ; own planes lz-1,lz -> the next rank's ghost planes -1,0
void exchange_split2(pf_solver *s, double *a) {
  const Geo &g = s->g;
  const size_t hp2 = (size_t)g.HX * (g.n + 4);
  pf_comm_exchange(s->comm, a + hp2 * 2, a + hp2 * g.lz, a, a + hp2 * (g.lz + 2), 2 * hp2, 1);
}

// neighbour barrier of the peer-store transport: every store into the neighbours' ghost planes issued so far
// has landed, and both neighbours have finished reading the buffer the next launch overwrites
// PF_PDL=0 switches the programmatic dependent launches of the half-sweep chains off (A/B measurements)
static bool pf_pdl_enabled() {
  const char *e = getenv("PF_PDL");
  return !(e && e[0] == '0');
}

void slab_barrier(pf_solver *s) {
  unsigned long long *prev_flags = static_cast<unsigned long long *>(s->peer->prev);
  unsigned long long *next_flags = static_cast<unsigned long long *>(s->peer->next);
  k_slab_barrier(prev_flags + 1, next_flags + 0, s->flags + 0, s->flags + 1, s->flags + 2, s->st);
}

// ---------------------------------------------------------------------------------------------
// phases
// ---------------------------------------------------------------------------------------------
void do_copy_old(pf_solver *s) {
  // u_old = u incl. halos (:85-100).  Kept as a device copy: the predictor writes only the interior
  // and the Poisson source then reads PREVIOUS-step halo values of u,v,w.
  const size_t bytes = (size_t)s->g.nat_elems * sizeof(double);
  PF_CUDA_OK(cudaMemcpyAsync(s->f.uo, s->f.u, bytes, cudaMemcpyDeviceToDevice, s->st));
  PF_CUDA_OK(cudaMemcpyAsync(s->f.vo, s->f.v, bytes, cudaMemcpyDeviceToDevice, s->st));
  if (s->g.dim == 3) PF_CUDA_OK(cudaMemcpyAsync(s->f.wo, s->f.w, bytes, cudaMemcpyDeviceToDevice, s->st));
}

// the same state by exchanging the buffers: u_old becomes the previous u (halos included) at no cost, and the
// new u receives only the halo shell -- its interior is about to be overwritten by the predictor (a1 in
// 48 B/cell/step of pure copy traffic removed).  Used inside pf_step; pf_copy_old keeps the copy.
void do_copy_old_by_swap(pf_solver *s) {
  std::swap(s->f.u, s->f.uo);
  std::swap(s->f.v, s->f.vo);
  if (s->g.dim == 3) std::swap(s->f.w, s->f.wo);
  k_shell_copy(s->g, s->f.uo, s->f.vo, s->g.dim == 3 ? s->f.wo : nullptr, s->f.u, s->f.v,
               s->g.dim == 3 ? s->f.w : nullptr, s->st);
}

void do_divergence(pf_solver *s) {
  const Geo &g = s->g;
  k_divergence(g, s->ph, s->f, s->st);
  if (!s->air) k_div_halo_y(g, s->ph, s->f, s->st);          // :206-213 ; air halos are zero
  if (g.dim == 3) {
    if (s->nranks == 1) {
      if (!s->air) {                                          // :215-222
        k_plane_copy_interior(g, s->f.div, 0, g.l, s->st);
        k_plane_copy_interior(g, s->f.div, g.l + 1, 1, s->st);
      }
    } else {
      exchange_nat(s, s->f.div, s->air ? 0 : 1);
    }
  }
}

void do_predictor(pf_solver *s) {
  k_predictor(s->g, s->ph, s->f, s->st);
  // slab interfaces are interior cells of the reference: the Poisson source needs the NEW w of the
  // neighbour plane there, while the periodic seam keeps the stale previous-step plane (H2).
  if (s->nranks > 1) exchange_nat(s, s->f.w, 0);
}

void do_rhs(pf_solver *s) {
  if (s->opp_top) {
    // air-condition slabs with a top outlet: its Dirichlet fold starts from bb(i,j,1) (sic, :702) -- the raw right-hand
    // side of GLOBAL plane 1, which rank 0 computes from its own planes and hands to the rank that owns plane l
    if (s->rank == 0) {
      k_raw_rhs_plane(s->g, s->ph, s->f, 1, s->tmp, s->st);
      pf_comm_send(s->comm, s->tmp, (size_t)s->g.plane, s->nranks - 1);
    } else if (s->rank == s->nranks - 1) {
      pf_comm_recv(s->comm, s->f.bb1, (size_t)s->g.plane, 0);
    }
  }
  k_rhs(s->g, s->ph, s->f, s->S, s->st);
}

void sor_refresh(pf_solver *s, int mask) {
  const Geo &g = s->g;
  if (!s->air) k_sor_halo_y(g, s->S[0].p, s->S[1].p, mask, s->st);
  if (g.dim != 3) return;
  if (s->nranks == 1) {
    if (!s->air) k_sor_halo_z_local(g, s->S[0].p, s->S[1].p, mask, s->st);
    return;
  }
  const int flip = g.l & 1;  // across the periodic seam the colour flips iff l is odd
  for (int cd = 0; cd < 2; ++cd) {
    if (!((mask >> cd) & 1)) continue;
    const int cs_lo = (s->rank == 0 && flip) ? cd ^ 1 : cd;
    const int cs_hi = (s->rank == s->nranks - 1 && flip) ? cd ^ 1 : cd;
    double *dst = s->S[cd].p;
    pf_comm_exchange(s->comm, s->S[cs_lo].p + g.hplane * 1, s->S[cs_hi].p + g.hplane * g.lz, dst,
                     dst + g.hplane * (g.lz + 1), (size_t)g.hplane, s->air ? 0 : 1);
  }
}

// the `iters` fused red+black launches of a solve.  On z-slab ranks the boundary planes reach the neighbours after
// every launch: stored by the kernel itself over NVLink -- and then either the TMA kernel also meets its neighbours
// itself (A.sync: nothing between the launches) or a one-thread barrier kernel follows each launch -- or, without
// peer access, by one grouped NCCL send/recv per iteration.
void sor_fused_launches(pf_solver *s, int iters) {
  const Geo &g = s->g;
  FusedArrays &A = s->fused;
  for (int it = 0; it < iters; ++it) {
    if (A.tma) k_tma_iteration(g, s->ph, A, it & 1, s->err_bits, s->st);
    else       k_fused_iteration(g, s->ph, A, it & 1, s->err_bits, s->st);
    if (!A.slab || A.sync) continue;
    if (s->peer) {
      slab_barrier(s);              // the kernel stored its boundary planes into the neighbours itself
    } else {
      const int out = (it & 1) ^ 1;
      pf_comm_group_begin(s->comm);
      exchange_split2(s, A.p[out][0]);
      exchange_split2(s, A.p[out][1]);
      pf_comm_group_end(s->comm);
    }
  }
}

// ... replayed from a CUDA graph where no NCCL call sits between the launches (one rank: the caller captures the whole
// solve; slab ranks with peer stores: this loop, barrier kernels included -- their sequence number lives on the device)
void sor_fused_loop(pf_solver *s, int iters) {
  FusedArrays &A = s->fused;
  const bool graph = A.slab && s->peer && s->cfg.use_graph != 0 && iters > 0;
  if (!graph) { sor_fused_launches(s, iters); return; }
  if (s->sor_graph_iters != iters) {
    if (s->sor_graph) { cudaGraphExecDestroy(s->sor_graph); s->sor_graph = nullptr; }
    cudaGraph_t gr = nullptr;
    const long long before = pf_launch_count();
    PF_CUDA_OK(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
    try {
      sor_fused_launches(s, iters);
    } catch (...) {
      cudaStreamEndCapture(s->st, &gr);
      if (gr) cudaGraphDestroy(gr);
      throw;
    }
    PF_CUDA_OK(cudaStreamEndCapture(s->st, &gr));
    s->sor_graph_nodes = pf_launch_count() - before;
    PF_CUDA_OK(cudaGraphInstantiate(&s->sor_graph, gr, 0));
    cudaGraphDestroy(gr);
    s->sor_graph_iters = iters;
  } else {
    for (long long q = 0; q < s->sor_graph_nodes; ++q) pf_count_launch();
  }
  PF_CUDA_OK(cudaGraphLaunch(s->sor_graph, s->st));
}

void sor_iterations(pf_solver *s, int iters) {
  const Geo &g = s->g;
  if (s->fused.enabled) {
    // variants 3/4/6: one fused red+black launch per iteration on the depth-2-ghost arrays (pf_sor_fused.cu,
    // pf_sor_tma.cu); on z-slab ranks the boundary planes are handed to the neighbours after every launch
    FusedArrays &A = s->fused;
    if (!A.slab) {
      // one rank: the natural pressure goes straight into both ping-pong buffers (do_sor skips the split layout)
      k_fused_gather_nat(g, A, s->f.p, s->st);
    } else {
      k_fused_gather(g, A, s->S[0].p, s->S[1].p, A.p[0][0], A.p[0][1], s->st);
      k_fused_gather(g, A, s->S[0].p, s->S[1].p, A.p[1][0], A.p[1][1], s->st);
    }
    k_fused_gather(g, A, s->S[0].bb, s->S[1].bb, A.bb[0], A.bb[1], s->st);
    if (A.slab) {
      // the ghost planes hold the neighbours' cells (x-halo slots included), not this slab's periodic images
      pf_comm_group_begin(s->comm);
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) exchange_split2(s, A.p[b][c]);
      for (int c = 0; c < 2; ++c) exchange_split2(s, A.bb[c]);
      pf_comm_group_end(s->comm);
      if (s->peer) slab_barrier(s);   // nobody stores into a neighbour before that neighbour's ghosts are set up
    }
    sor_fused_loop(s, iters);
    const int fin = iters & 1;
    if (!A.slab) {
      // ... and back: interior + the periodic rows and planes of the closing halo refresh (:588-605)
      k_fused_scatter_nat(g, A, fin, s->f.p, s->st);
      return;
    }
    k_fused_scatter(g, A, A.p[fin][0], A.p[fin][1], s->S[0].p, s->S[1].p, s->st);
    sor_refresh(s, 3);  // :588-605
    return;
  }
  // colour order: 3D (i+j+k) even first (:492-517); 2D (i+j) odd first (ibm_2d_uniform_omp_cpu.f90:339-352)
  const int order[2] = {g.dim == 3 ? 0 : 1, g.dim == 3 ? 1 : 0};
  if (s->nranks > 1 && s->cfg.sor_variant != 5) {
    // z-slab ranks: sweep the two boundary planes first, ship them to the neighbours on a
    // high-priority stream, and sweep the interior planes while the planes are in flight.
    sor_refresh(s, 1 << (order[0] ^ 1));
    const int flip = g.l & 1;
    for (int it = 0; it < iters; ++it)
      for (int half = 0; half < 2; ++half) {
        const int c = order[half];
        const int with_err = (g.dim == 2) || half == 1;
        k_sor_sweep(g, s->ph, s->S, c, with_err, s->err_bits, s->cfg.sor_variant, s->st, 1, g.lz - 1, 2);
        PF_CUDA_OK(cudaEventRecord(s->ev_edge, s->st));
        PF_CUDA_OK(cudaStreamWaitEvent(s->comm_st, s->ev_edge, 0));
        {
          const int cs_lo = (s->rank == 0 && flip) ? c ^ 1 : c;
          const int cs_hi = (s->rank == s->nranks - 1 && flip) ? c ^ 1 : c;
          // with an odd l the seam ranks send the OTHER colour's array (the seam flips the colour).  No in-order
          // fall-back is needed: they send its planes 1 / lz, which the half-sweep of colour c never writes
          double *dst = s->S[c].p;
          pf_comm_exchange(s->comm, s->S[cs_lo].p + g.hplane * 1, s->S[cs_hi].p + g.hplane * g.lz, dst,
                           dst + g.hplane * (g.lz + 1), (size_t)g.hplane, s->air ? 0 : 1, s->comm_st);
        }
        PF_CUDA_OK(cudaEventRecord(s->ev_comm, s->comm_st));
        k_sor_sweep(g, s->ph, s->S, c, with_err, s->err_bits, s->cfg.sor_variant, s->st, 2, 1, g.lz - 2);
        if (!s->air) k_sor_halo_y(g, s->S[0].p, s->S[1].p, 1 << c, s->st);
        PF_CUDA_OK(cudaStreamWaitEvent(s->st, s->ev_comm, 0));
      }
    sor_refresh(s, 3);  // :588-605
    return;
  }
  // 2D: the sweep kernel keeps the periodic y-halo rows itself (pf_sor.cu, YIMG); one refresh up front suffices
  const bool self_halo = s->nranks == 1 && pf_sor_stores_y_images(g);
  if (self_halo) sor_refresh(s, 3);
  if (s->cfg.sor_variant == 8) {
    // temporally blocked: T iterations per launch in shared memory (pf_sor_tb2d.cu); the tiles wrap the periodic y
    // direction themselves and never read the halo rows
    k_sor_tb2d(g, s->ph, s->S, s->p_alt, iters, s->err_bits, s->st);
    sor_refresh(s, 3);  // :588-605
    return;
  }
  if (s->cfg.sor_variant == 7 && pf_persistent_applicable(g, s->air, s->nranks)) {
    // opt-in: the same half-sweeps with the iteration loop on the device (pf_sor_persistent.cu)
    k_sor_persistent(g, s->ph, s->S, iters, s->err_bits, s->st);
    sor_refresh(s, 3);  // :588-605
    return;
  }
  // where a solve is nothing but half-sweeps (2D with self-kept halo rows; air-condition: no refresh inside the solve,
  // ibm_3d_air_condition_omp_cpu.f90:509-527) every launch after the first overlaps the tail of its predecessor
  // (programmatic dependent launch, pf_sor.cu)
  const bool chain = s->nranks == 1 && (self_halo || s->air) && s->cfg.sor_variant == 1 && pf_pdl_enabled();
  for (int it = 0; it < iters; ++it)
    for (int half = 0; half < 2; ++half) {
      const int c = order[half];
      if (!self_halo) sor_refresh(s, 1 << (c ^ 1));  // only the colour about to be read
      // error: 3D only after the second half-sweep (:575-583); 2D in both (:351,:385)
      const int with_err = (g.dim == 2) || half == 1;
      k_sor_sweep(g, s->ph, s->S, c, with_err, s->err_bits, s->cfg.sor_variant, s->st, 1, 1, -1,
                  chain && (it > 0 || half > 0));
    }
  sor_refresh(s, 3);  // :588-605
}

void do_sor(pf_solver *s, int iters, double *err_slot_dev) {
  const Geo &g = s->g;
  PF_CUDA_OK(cudaMemsetAsync(s->err_bits, 0, sizeof(unsigned long long), s->st));
  // the fused kernels on one rank read and write the natural pressure themselves (k_fused_gather_nat / _scatter_nat)
  const bool direct = s->fused.enabled && !s->fused.slab;
  if (!direct) k_nat_to_split(g, s->f.p, s->S[0].p, s->S[1].p, s->st);
  // (variant 7 is a single cooperative launch: nothing to replay)
  const bool graph = s->cfg.use_graph != 0 && s->nranks == 1 && iters > 0 && s->cfg.sor_variant != 7;
  if (!graph) {
    sor_iterations(s, iters);
  } else {
    if (s->sor_graph_iters != iters) {
      if (s->sor_graph) { cudaGraphExecDestroy(s->sor_graph); s->sor_graph = nullptr; }
      cudaGraph_t gr = nullptr;
      const long long before = pf_launch_count();
      PF_CUDA_OK(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
      try {
        sor_iterations(s, iters);
      } catch (...) {
        cudaStreamEndCapture(s->st, &gr);
        if (gr) cudaGraphDestroy(gr);
        throw;
      }
      PF_CUDA_OK(cudaStreamEndCapture(s->st, &gr));
      s->sor_graph_nodes = pf_launch_count() - before;
      PF_CUDA_OK(cudaGraphInstantiate(&s->sor_graph, gr, 0));
      cudaGraphDestroy(gr);
      s->sor_graph_iters = iters;
    } else {
      for (long long q = 0; q < s->sor_graph_nodes; ++q) pf_count_launch();
    }
    PF_CUDA_OK(cudaGraphLaunch(s->sor_graph, s->st));
  }
  if (!direct) k_split_to_nat(g, s->S[0].p, s->S[1].p, s->f.p, s->st);
  if (err_slot_dev)
    PF_CUDA_OK(cudaMemcpyAsync(err_slot_dev, s->err_bits, sizeof(double), cudaMemcpyDeviceToDevice, s->st));
}

void do_project(pf_solver *s) { k_project(s->g, s->ph, s->f, s->st); }

void do_boundary(pf_solver *s) {
  const Geo &g = s->g;
  k_boundary_local(g, s->ph, s->f, s->rank, s->nranks, s->st);
  if (g.dim != 3) return;
  double *arr[4] = {s->f.u, s->f.v, s->f.w, s->f.p};
  if (s->nranks == 1) {
    if (!s->air)                                              // :735-748
      for (double *a : arr) {
        k_plane_copy_full(g, a, 0, g.l, s->st);
        k_plane_copy_full(g, a, g.l + 1, 1, s->st);
      }
  } else {
    for (double *a : arr) exchange_nat(s, a, s->air ? 0 : 1);
  }
}

void ensure_errs(pf_solver *s, int n) {
  if (n <= s->errs_cap) return;
  void *p = nullptr;
  PF_CUDA_OK(cudaMalloc(&p, (size_t)n * sizeof(double)));
  s->allocs.push_back(p);
  s->errs_dev = static_cast<double *>(p);
  s->errs_cap = n;
}

void ensure_events(pf_solver *s, size_t n) {
  while (s->events.size() < n) {
    cudaEvent_t e;
    PF_CUDA_OK(cudaEventCreate(&e));
    s->events.push_back(e);
  }
}

void fetch_errors(pf_solver *s, int n, double *host) {
  if (s->nranks > 1) pf_comm_allreduce_max(s->comm, s->errs_dev, (size_t)n);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  if (host) PF_CUDA_OK(cudaMemcpy(host, s->errs_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
}

void run_steps(pf_solver *s, int nsteps, double *p_error) {
  if (!s->porosity_set) throw std::string("pf_set_porosity must be called before stepping");
  if (nsteps <= 0) return;
  ensure_errs(s, nsteps);
  ensure_events(s, 2 + 2 * (size_t)nsteps);
  pf_launch_count_reset();
  PF_CUDA_OK(cudaEventRecord(s->events[0], s->st));
  for (int it = 0; it < nsteps; ++it) {
    do_copy_old_by_swap(s);
    do_divergence(s);
    do_predictor(s);
    do_rhs(s);
    PF_CUDA_OK(cudaEventRecord(s->events[2 + 2 * it], s->st));
    do_sor(s, s->cfg.iter_max, s->errs_dev + it);
    PF_CUDA_OK(cudaEventRecord(s->events[3 + 2 * it], s->st));
    do_project(s);
    do_boundary(s);
  }
  PF_CUDA_OK(cudaEventRecord(s->events[1], s->st));
  fetch_errors(s, nsteps, p_error);
  PF_CUDA_OK(cudaEventSynchronize(s->events[1]));
  float ms = 0;
  PF_CUDA_OK(cudaEventElapsedTime(&ms, s->events[0], s->events[1]));
  s->ms_total = ms;
  s->ms_sor = 0;
  for (int it = 0; it < nsteps; ++it) {
    PF_CUDA_OK(cudaEventElapsedTime(&ms, s->events[2 + 2 * it], s->events[3 + 2 * it]));
    s->ms_sor += ms;
  }
  s->launches = pf_launch_count();
}

void validate(const pf_config *c) {
  if (!c) throw std::string("null config");
  if (c->struct_size != (int)sizeof(pf_config)) throw std::string("pf_config.struct_size mismatch (ABI)");
  if (c->solver_case < PF_IBM2_UNIFORM || c->solver_case > PF_IBM3_AIRCOND) throw std::string("bad solver_case");
  const bool d3 = c->solver_case >= PF_IBM3_UNIFORM;
  if (c->m < 2 || c->n < 2) throw std::string("m and n must be >= 2");
  if (d3 && c->l < 2) throw std::string("l must be >= 2 for the 3D cases");
  if (!(c->dx > 0) || !(c->dy > 0) || (d3 && !(c->dz > 0)) || !(c->dt > 0)) throw std::string("dx,dy,dz,dt must be > 0");
  if (c->iter_max < 0) throw std::string("iter_max < 0");
  if (c->nranks < 1 || c->rank < 0 || c->rank >= c->nranks) throw std::string("bad rank/nranks");
  if (c->nranks > 1) {
    if (!d3) throw std::string("the 2D cases run on one GPU (nothing to decompose along z)");
    if (c->l / c->nranks < 2) throw std::string("need at least 2 planes per rank");
    if (c->halo_transport < 0 || c->halo_transport > 3) throw std::string("halo_transport must be 0, 1, 2 or 3");
  }
  if (c->host_ldx && c->host_ldx < c->m + 2) throw std::string("host_ldx < m+2");
  if (c->host_ldy && c->host_ldy < c->n + 2) throw std::string("host_ldy < n+2");
  if (c->solver_case == PF_IBM3_AIRCOND)
    for (int i = 0; i < 6; ++i)
      if (c->wall[i] < 0 || c->wall[i] > 2) throw std::string("wall code must be 0, 1 or 2");
}

void build(pf_solver *s) {
  const pf_config &c = s->cfg;
  Geo &g = s->g;
  const bool d3 = c.solver_case >= PF_IBM3_UNIFORM;
  s->air = c.solver_case == PF_IBM3_AIRCOND;
  s->uniform3 = c.solver_case == PF_IBM3_UNIFORM;
  s->rank = c.rank;
  s->nranks = c.nranks;
  g.dim = d3 ? 3 : 2;
  g.m = c.m;
  g.n = c.n;
  g.l = d3 ? c.l : 1;
  if (d3) {
    const int base = c.l / c.nranks, rem = c.l % c.nranks;
    g.lz = base + (c.rank < rem ? 1 : 0);
    g.koff = c.rank * base + std::min(c.rank, rem);
  } else {
    g.lz = 1;
    g.koff = 0;
  }
  g.kin0 = d3 ? 1 : 0;
  g.X0 = 15;
  g.NX = round_up(g.X0 + g.m + 2, 16);
  g.NY = g.n + 2;
  g.NZ = d3 ? g.lz + 2 : 1;
  g.H0 = 4;
  g.HX = round_up(g.H0 + (g.m + 1) / 2 + 3, 4);
  g.plane = (long long)g.NX * g.NY;
  g.hplane = (long long)g.HX * g.NY;
  g.nat_elems = g.plane * g.NZ;
  g.split_elems = g.hplane * g.NZ;
  s->host_ldx = c.host_ldx ? c.host_ldx : c.m + 2;
  s->host_ldy = c.host_ldy ? c.host_ldy : c.n + 2;

  Phys &ph = s->ph;
  ph.dx = c.dx; ph.dy = c.dy; ph.dz = c.dz; ph.dt = c.dt;
  ph.xnue = c.xnue; ph.xlambda = c.xlambda; ph.density = c.density; ph.thickness = c.thickness;
  auto mkinv = [](double d) {
    Inv v;
    v.d = d;
    v.r = 1.0 / d;
    v.fast = (std::isfinite(d) && fabs(d) > 1e-100 && fabs(d) < 1e100) ? 1 : 0;
    return v;
  };
  ph.ix = mkinv(c.dx); ph.iy = mkinv(c.dy); ph.iz = mkinv(c.dz);
  // (thickness*dx)**2 -> (thickness*dx)*(thickness*dx), as gfortran expands the integer power
  ph.itx2 = mkinv((c.thickness * c.dx) * (c.thickness * c.dx));
  ph.ity2 = mkinv((c.thickness * c.dy) * (c.thickness * c.dy));
  ph.itz2 = mkinv((c.thickness * c.dz) * (c.thickness * c.dz));
  ph.dtrho = c.dt / c.density;
  {  // the check-free reciprocal division of the predictor is used only with moderate parameters
    const double par[] = {c.dt, c.xnue, c.dx, c.dy, c.dz, c.thickness};
    bool ok = std::isfinite(c.xlambda) && fabs(c.xlambda) < 1e30;
    for (double v : par) ok = ok && std::isfinite(v) && fabs(v) > 1e-30 && fabs(v) < 1e30;
    if (!ok) ph.ix.fast = 0;
  }
  ph.relux = c.relux_factor;
  ph.nonslip = c.nonslip;
  ph.scase = c.solver_case;
  ph.inlet_velocity = c.inlet_velocity;
  ph.outlet_pressure = c.outlet_pressure;
  for (int i = 0; i < 6; ++i) ph.wall[i] = c.wall[i];
  // loop constants evaluated on the host with libm, exactly as the reference's scalar code does
  const double pi = atan(1.) * 4.;
  if (d3) {
    ph.uin = c.inlet_velocity * cos(c.AoA / 1300. * pi);  // sic, ibm_3d_uniform_omp_cpu.f90:695-696
    ph.vin = c.inlet_velocity * sin(c.AoA / 1300. * pi);
    ph.u0 = s->air ? 0. : c.inlet_velocity * cos(c.AoA / 360 * pi);  // :782-783 ; air :1197-1198
    ph.v0 = s->air ? 0. : c.inlet_velocity * sin(c.AoA / 360 * pi);
  } else {
    ph.uin = c.inlet_velocity * cos(c.AoA / 180. * pi);   // ibm_2d_uniform_omp_cpu.f90:478-479
    ph.vin = c.inlet_velocity * sin(c.AoA / 180. * pi);
    ph.u0 = c.inlet_velocity * cos(c.AoA / 180 * pi);     // :560-561
    ph.v0 = c.inlet_velocity * sin(c.AoA / 180 * pi);
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    throw std::string("no CUDA device (there is no CPU fallback)");
  }
  if (c.device >= 0) PF_CUDA_OK(cudaSetDevice(c.device));
  PF_CUDA_OK(cudaGetDevice(&s->device));
  PF_CUDA_OK(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));

  double **nat[] = {&s->f.u, &s->f.v, &s->f.w, &s->f.p, &s->f.uo, &s->f.vo, &s->f.wo, &s->f.eps, &s->f.div, &s->tmp};
  for (double **a : nat) *a = dalloc(s, g.nat_elems);
  // air-condition on z-slabs: the two places where the reference reads the opposite z face (pf_internal.cuh, Fields)
  s->opp_top = s->air && c.nranks > 1 && c.wall[PF_TOP] == 2;
  s->opp_bottom = s->air && c.nranks > 1 && c.wall[PF_BOTTOM] == 1;
  if (s->opp_top && c.rank == c.nranks - 1) s->f.bb1 = dalloc(s, g.plane);
  if (s->opp_bottom && c.rank == 0) s->f.eps_top = dalloc(s, g.plane);
  for (int cidx = 0; cidx < 2; ++cidx) {
    SplitSet &S = s->S[cidx];
    double **sp[] = {&S.ap, &S.ae, &S.aw, &S.an, &S.as, &S.at, &S.ab, &S.bb, &S.p, &S.eps};
    for (double **a : sp) {
      const bool zcoef = (a == &S.at || a == &S.ab);
      *a = (zcoef && !d3) ? nullptr : dalloc(s, g.split_elems);
    }
  }
  void *eb = nullptr;
  PF_CUDA_OK(cudaMalloc(&eb, sizeof(unsigned long long)));
  s->allocs.push_back(eb);
  s->err_bits = static_cast<unsigned long long *>(eb);
  // SOR kernel selection.  1 = colour half-sweeps (works everywhere); 3/4 = fused red+black pass with
  // register prefetch (32x16 / 32x8 tiles); 6 = fused pass with the TMA pipeline; 2 = coefficients from
  // porosity (measured slower, kept for the record).  0 = auto: on one GPU, where the fused pass applies
  // (3D uniform, even n and l), take the TMA pipeline unless the rows are too short for its 30-column
  // tiles (then the register-prefetch kernel) -- the faster one in each regime on B200
  // (profiles/r01_fused_summary.md).  z-slab ranks take the same fused kernels where they apply: the
  // boundary planes go straight into the neighbours' ghost planes (peer stores over NVLink, or one NCCL
  // group per iteration); otherwise the half-sweeps, whose boundary-plane exchange overlaps the interior sweep.
  int variant = c.sor_variant;
  const bool fused_ok = pf_fused_applicable(g, s->ph, c.nranks), tma_ok = pf_tma_applicable(g, s->ph, c.nranks);
  if (variant == 0) {
    variant = 1;
    if (fused_ok) {
      const int cols = ((g.m + 1) >> 1) + 2;
      variant = (cols >= 60 && tma_ok) ? 6 : 3;
    }
  }
  // A requested kernel that does not apply to this case is replaced by the one that does, and pf_get_sor_variant
  // reports THAT one: 6 -> 3 without the TMA preconditions, 3/4/6 -> 1 where the fused pass does not apply (2D,
  // air-condition, odd n or l, thin slabs), 7 -> 1 outside single-rank 2D / air-condition; unknown numbers -> 1.
  if (variant == 6 && !tma_ok) variant = 3;
  if ((variant == 3 || variant == 4) && !fused_ok) variant = 1;
  if (variant == 7 && !pf_persistent_applicable(g, s->air, c.nranks)) variant = 1;
  if (variant == 8 && !pf_tb2d_applicable(g, c.nranks)) variant = 1;
  if (variant < 1 || variant > 8) variant = 1;
  s->cfg.sor_variant = variant;
  if (variant == 8)
    for (int cc = 0; cc < 2; ++cc) s->p_alt[cc] = dalloc(s, g.split_elems);
  s->fused.enabled = variant == 3 || variant == 4 || variant == 6;
  s->fused.tma = variant == 6;
  double *block = nullptr;
  if (s->fused.enabled) {
    FusedArrays &A = s->fused;
    A.rpt = (variant == 4) ? 1 : 2;
    A.slab = c.nranks > 1;
    const long long ne = pf_fused_elems(g);
    for (int cc = 0; cc < 2; ++cc) {
      A.cx[cc] = dalloc(s, ne); A.cy[cc] = dalloc(s, ne); A.cz[cc] = dalloc(s, ne); A.bb[cc] = dalloc(s, ne);
    }
    // one allocation for what a neighbour rank may write: 32 doubles of flags, then the four pressure buffers
    block = dalloc(s, 32 + 4 * ne);
    s->flags = reinterpret_cast<unsigned long long *>(block);
    const long long hp2 = (long long)g.HX * (g.n + 4);
    for (int b = 0; b < 2; ++b)
      for (int cc = 0; cc < 2; ++cc) {
        A.p[b][cc] = block + 32 + (2 * b + cc) * ne;
        A.img_lo[b][cc] = A.slab ? nullptr : A.p[b][cc];   // one rank: the periodic images in the same array
        A.img_hi[b][cc] = A.slab ? nullptr : A.p[b][cc];
      }
    A.dk_lo = (long long)g.lz * hp2;
    A.dk_hi = -(long long)g.lz * hp2;
  }
  PF_CUDA_OK(cudaMemsetAsync(eb, 0, sizeof(unsigned long long), s->st));
  if (c.nranks > 1) {
    s->comm = pf_comm_create(c.rank, c.nranks, c.nccl_unique_id, s->st);
    if (s->fused.enabled) {
      FusedArrays &A = s->fused;
      PF_CUDA_OK(cudaStreamSynchronize(s->st));   // the flags are zero before any neighbour can see them
      if (c.halo_transport != 1) s->peer = pf_peer_open(s->comm, block, s->peer_why);
      else s->peer_why = "halo_transport = 1";
      if (!s->peer && c.halo_transport >= 2)
        throw std::string("halo_transport = 2 / 3 (peer stores) is not available: ") + s->peer_why;
      if (s->peer) {
        // the neighbours' blocks have the same layout; their slabs may be one plane thicker or thinner
        const long long ne = pf_fused_elems(g);
        const long long hp2 = (long long)g.HX * (g.n + 4);
        const int base = c.l / c.nranks, rem = c.l % c.nranks;
        const int prev = (c.rank + c.nranks - 1) % c.nranks, next = (c.rank + 1) % c.nranks;
        const int lz_prev = base + (prev < rem ? 1 : 0), lz_next = base + (next < rem ? 1 : 0);
        const long long ne_prev = ((long long)hp2 * (lz_prev + 4) + 31) / 32 * 32;
        const long long ne_next = ((long long)hp2 * (lz_next + 4) + 31) / 32 * 32;
        (void)ne;
        for (int b = 0; b < 2; ++b)
          for (int cc = 0; cc < 2; ++cc) {
            A.img_lo[b][cc] = static_cast<double *>(s->peer->prev) + 32 + (2 * b + cc) * ne_prev;
            A.img_hi[b][cc] = static_cast<double *>(s->peer->next) + 32 + (2 * b + cc) * ne_next;
          }
        A.dk_lo = (long long)lz_prev * hp2;   // my plane k (1,2)      -> the previous rank's plane lz_prev + k
        A.dk_hi = -(long long)g.lz * hp2;     // my plane k (lz-1, lz) -> the next rank's plane k - lz
        if (A.tma && c.halo_transport == 3) {
          // the TMA kernel meets its neighbours itself (pf_sor_tma.cu, slab_sync): I am the previous rank's "next"
          A.sync = s->flags;
          A.sync_to_prev = static_cast<unsigned long long *>(s->peer->prev) + PF_SY_FROM_NEXT;
          A.sync_to_next = static_cast<unsigned long long *>(s->peer->next) + PF_SY_FROM_PREV;
        }
      }
    }
    int lo = 0, hi = 0;
    PF_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PF_CUDA_OK(cudaStreamCreateWithPriority(&s->comm_st, cudaStreamNonBlocking, hi));
    PF_CUDA_OK(cudaEventCreateWithFlags(&s->ev_edge, cudaEventDisableTiming));
    PF_CUDA_OK(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
  }
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
}

double *field_ptr(pf_solver *s, int field) {
  switch (field) {
    case PF_F_U: return s->f.u;   case PF_F_V: return s->f.v;   case PF_F_W: return s->f.w;
    case PF_F_P: return s->f.p;   case PF_F_UOLD: return s->f.uo; case PF_F_VOLD: return s->f.vo;
    case PF_F_WOLD: return s->f.wo; case PF_F_POROSITY: return s->f.eps; case PF_F_DIV: return s->f.div;
  }
  return nullptr;
}

double *split_ptr(SplitSet &S, int field) {
  switch (field) {
    case PF_F_AP: return S.ap; case PF_F_AE: return S.ae; case PF_F_AW: return S.aw;
    case PF_F_AN: return S.an; case PF_F_AS: return S.as; case PF_F_AT: return S.at;
    case PF_F_AB: return S.ab; case PF_F_BB: return S.bb;
  }
  return nullptr;
}

void make_current(pf_solver *s) {
  int cur = -1;
  if (s->device >= 0 && (cudaGetDevice(&cur) != cudaSuccess || cur != s->device)) PF_CUDA_OK(cudaSetDevice(s->device));
}

}  // namespace

// every entry point makes the solver's device current first: a host thread that drives several solvers (one per
// GPU) or that changed its device between calls must not launch this solver's kernels elsewhere
#define PF_API_BEGIN(s)                         \
  if (!(s)) return 1;                           \
  try {                                         \
    make_current(s);
#define PF_API_END(s)                           \
    return 0;                                   \
  } catch (const std::string &e) {              \
    (s)->err = e;                               \
    return 1;                                   \
  } catch (const std::exception &e) {           \
    (s)->err = e.what();                        \
    return 1;                                   \
  }

extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }

void pf_config_init(pf_config *c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
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
  // shipped wall_conditions (ibm_3d_air_condition_omp_cpu.f90:10-15): top inlet, south outlet
  c->wall[PF_TOP] = 1;
  c->wall[PF_SOUTH] = 2;
  c->device = -1;
  c->nranks = 1;
  c->use_graph = 1;
}

int pf_create(pf_solver **out, const pf_config *cfg) {
  if (!out) { g_create_error = "null out pointer"; return 1; }
  *out = nullptr;
  pf_solver *s = nullptr;
  try {
    validate(cfg);
    s = new pf_solver;
    s->cfg = *cfg;
    build(s);
    *out = s;
    return 0;
  } catch (const std::string &e) {
    g_create_error = e;
  } catch (const std::exception &e) {
    g_create_error = e.what();
  }
  if (s) pf_destroy(s);
  return 1;
}

void pf_destroy(pf_solver *s) {
  if (!s) return;
  if (s->device >= 0) cudaSetDevice(s->device);
  if (s->st) cudaStreamSynchronize(s->st);
  if (s->sor_graph) cudaGraphExecDestroy(s->sor_graph);
  pf_tma_release(s->fused);
  if (s->peer) pf_peer_close(s->comm, s->peer);
  pf_comm_destroy(s->comm);
  if (s->ev_edge) cudaEventDestroy(s->ev_edge);
  if (s->ev_comm) cudaEventDestroy(s->ev_comm);
  if (s->comm_st) cudaStreamDestroy(s->comm_st);
  for (cudaEvent_t e : s->events) cudaEventDestroy(e);
  if (s->text_dev) cudaFree(s->text_dev);
  for (void *p : s->allocs) cudaFree(p);
  if (s->st) cudaStreamDestroy(s->st);
  delete s;
}

const char *pf_last_error(const pf_solver *s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int pf_comm_unique_id(void *out128) {
  std::string e;
  const int rc = pf_comm_get_unique_id(out128, e);
  if (rc) g_create_error = e;
  return rc;
}

int pf_local_slab(const pf_solver *s, int *k_first, int *k_count) {
  if (!s) return 1;
  if (k_first) *k_first = s->g.koff + 1;
  if (k_count) *k_count = s->g.lz;
  return 0;
}

int pf_set_porosity(pf_solver *s, const double *porosity) {
  PF_API_BEGIN(s)
  upload_field(s, s->f.eps, porosity);
  if (s->opp_bottom) {
    // air-condition slabs with a bottom inlet: its fluid test reads porosity(i,j,l) (sic, :948) -- the plane of the last rank
    if (s->rank == s->nranks - 1) pf_comm_send(s->comm, s->f.eps + s->g.plane * s->g.lz, (size_t)s->g.plane, 0);
    else if (s->rank == 0)        pf_comm_recv(s->comm, s->f.eps_top, (size_t)s->g.plane, s->nranks - 1);
  }
  k_coefficients(s->g, s->ph, s->f, s->S, s->st);
  k_nat_to_split(s->g, s->f.eps, s->S[0].eps, s->S[1].eps, s->st);
  if (s->fused.enabled) {
    k_fused_build_faces(s->g, s->ph, s->f.eps, s->fused, s->st);
    if (s->fused.tma) pf_tma_schedule(s->g, s->fused);
    if (s->fused.slab) {   // ghost planes of the face coefficients = the neighbours' planes
      FusedArrays &A = s->fused;
      pf_comm_group_begin(s->comm);
      for (int cc = 0; cc < 2; ++cc) {
        exchange_split2(s, A.cx[cc]);
        exchange_split2(s, A.cy[cc]);
        exchange_split2(s, A.cz[cc]);
      }
      pf_comm_group_end(s->comm);
    }
  }
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  s->porosity_set = true;
  PF_API_END(s)
}

int pf_upload(pf_solver *s, const double *u, const double *v, const double *w, const double *p) {
  PF_API_BEGIN(s)
  upload_field(s, s->f.u, u);
  upload_field(s, s->f.v, v);
  if (s->g.dim == 3) upload_field(s, s->f.w, w);
  upload_field(s, s->f.p, p);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_download(pf_solver *s, double *u, double *v, double *w, double *p) {
  PF_API_BEGIN(s)
  download_field(s, s->f.u, u);
  download_field(s, s->f.v, v);
  if (s->g.dim == 3) download_field(s, s->f.w, w);
  download_field(s, s->f.p, p);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

// Collective over the slab ranks: afterwards rank 0's host arrays (GLOBAL shape, whatever host_is_slab says) hold the
// whole fields -- what the reference's output routines (lib/output.f90) want to see.  The other ranks pass their
// planes through the rank-0 GPU (NCCL send/recv into its scratch array, then one strided copy to the host); their
// host pointers are ignored and may be NULL.  On one rank this is pf_download.
int pf_gather(pf_solver *s, double *u, double *v, double *w, double *p) {
  PF_API_BEGIN(s)
  const Geo &g = s->g;
  double *dev[4] = {s->f.u, s->f.v, g.dim == 3 ? s->f.w : nullptr, s->f.p};
  double *host[4] = {u, v, w, p};
  if (s->nranks == 1) {
    for (int q = 0; q < 4; ++q)
      if (dev[q]) download_field(s, dev[q], host[q]);
    PF_CUDA_OK(cudaStreamSynchronize(s->st));
  } else {
    const int base = s->cfg.l / s->nranks, rem = s->cfg.l % s->nranks;
    for (int q = 0; q < 4; ++q) {
      if (!dev[q]) continue;
      if (s->rank == 0) {
        if (!host[q]) throw std::string("pf_gather: null host array on rank 0");
        // own planes 0 .. lz (the global ghost plane 0 included)
        xfer(s, dev[q], host[q], false, 0, 0, g.lz + 1);
        for (int r = 1; r < s->nranks; ++r) {
          const int lz_r = base + (r < rem ? 1 : 0), koff_r = r * base + std::min(r, rem);
          const int np = lz_r + (r == s->nranks - 1 ? 1 : 0);       // + the global ghost plane l+1 from the last rank
          pf_comm_recv(s->comm, s->tmp + g.plane, (size_t)g.plane * np, r);
          xfer(s, s->tmp, host[q], false, 1, koff_r + 1, np);
          PF_CUDA_OK(cudaStreamSynchronize(s->st));                 // tmp is reused by the next rank
        }
      } else {
        const int np = g.lz + (s->rank == s->nranks - 1 ? 1 : 0);
        pf_comm_send(s->comm, dev[q] + g.plane, (size_t)g.plane * np, 0);
      }
    }
    PF_CUDA_OK(cudaStreamSynchronize(s->st));
  }
  PF_API_END(s)
}

int pf_get_field(pf_solver *s, int field, double *host) {
  PF_API_BEGIN(s)
  double *d = field_ptr(s, field);
  if (!d) {
    double *a0 = split_ptr(s->S[0], field), *a1 = split_ptr(s->S[1], field);
    if (!a0 || !a1) throw std::string("pf_get_field: unknown field (or a z coefficient in 2D)");
    k_split_to_nat(s->g, a0, a1, s->tmp, s->st);
    d = s->tmp;
  }
  download_field(s, d, host);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_set_field(pf_solver *s, int field, const double *host) {
  PF_API_BEGIN(s)
  double *d = field_ptr(s, field);
  if (d) {
    upload_field(s, d, host);
  } else {
    double *a0 = split_ptr(s->S[0], field), *a1 = split_ptr(s->S[1], field);
    if (!a0 || !a1) throw std::string("pf_set_field: unknown field (or a z coefficient in 2D)");
    upload_field(s, s->tmp, host);
    k_nat_to_split(s->g, s->tmp, a0, a1, s->st);
  }
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_step(pf_solver *s, int nsteps, double *p_error) {
  PF_API_BEGIN(s)
  run_steps(s, nsteps, p_error);
  PF_API_END(s)
}

int pf_step_host(pf_solver *s, int nsteps, double *u, double *v, double *w, double *p, double *p_error) {
  PF_API_BEGIN(s)
  upload_field(s, s->f.u, u);
  upload_field(s, s->f.v, v);
  if (s->g.dim == 3) upload_field(s, s->f.w, w);
  upload_field(s, s->f.p, p);
  run_steps(s, nsteps, p_error);
  download_field(s, s->f.u, u);
  download_field(s, s->f.v, v);
  if (s->g.dim == 3) download_field(s, s->f.w, w);
  download_field(s, s->f.p, p);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_initial_conditions(pf_solver *s) {
  PF_API_BEGIN(s)
  if (!s->porosity_set) throw std::string("pf_set_porosity must be called first");
  k_initial(s->g, s->ph, s->f, s->st);
  do_boundary(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_copy_old(pf_solver *s) {
  PF_API_BEGIN(s)
  do_copy_old(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_divergence(pf_solver *s) {
  PF_API_BEGIN(s)
  do_divergence(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_predictor(pf_solver *s) {
  PF_API_BEGIN(s)
  do_predictor(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_build_poisson(pf_solver *s) {
  PF_API_BEGIN(s)
  if (!s->porosity_set) throw std::string("pf_set_porosity must be called first");
  do_rhs(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_sor(pf_solver *s, int iters, double *p_error) {
  PF_API_BEGIN(s)
  ensure_errs(s, 1);
  do_sor(s, iters, s->errs_dev);
  fetch_errors(s, 1, p_error);
  PF_API_END(s)
}

int pf_project(pf_solver *s) {
  PF_API_BEGIN(s)
  do_project(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_boundary(pf_solver *s) {
  PF_API_BEGIN(s)
  do_boundary(s);
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_force_log_2d(pf_solver *s, double radius, double *out8) {
  PF_API_BEGIN(s)
  if (s->g.dim != 2) throw std::string("pf_force_log_2d is for the 2D cases (lib/output.f90:244-305)");
  if (!out8) throw std::string("null output");
  const int blocks = pf_sm_count() * 4;
  if (!s->force_scratch) s->force_scratch = dalloc(s, 4 * blocks + 4);
  k_force2d(s->g, s->ph, s->f, s->force_scratch, blocks, s->force_scratch + 4 * blocks, s->st);
  double h[4];
  PF_CUDA_OK(cudaMemcpyAsync(h, s->force_scratch + 4 * blocks, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  out8[0] = h[0]; out8[1] = h[1]; out8[2] = h[2]; out8[3] = h[3];
  out8[4] = h[0] + h[2];                                   // force_x = force_px + force_vx (:293)
  out8[5] = h[1] + h[3];
  const double den = s->cfg.density * (s->cfg.inlet_velocity * s->cfg.inlet_velocity) * radius;
  out8[6] = out8[4] / den;                                 // cd (:296)
  out8[7] = out8[5] / den;                                 // cl (:297)
  PF_API_END(s)
}

int pf_force_log_3d(pf_solver *s, double radius, double *out12) {
  PF_API_BEGIN(s)
  if (s->g.dim != 3) throw std::string("pf_force_log_3d is for the 3D cases (lib/output.f90:1090-1165)");
  if (!out12) throw std::string("null output");
  const int blocks = pf_sm_count() * 4;
  if (!s->force_scratch) s->force_scratch = dalloc(s, 6 * blocks + 8);
  double *sums = s->force_scratch + 6 * blocks;
  k_force3d(s->g, s->ph, s->f, s->force_scratch, blocks, sums, s->st);
  if (s->nranks > 1) pf_comm_allreduce_sum(s->comm, sums, 6);   // the slabs' partial sums
  double h[6];
  PF_CUDA_OK(cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  for (int q = 0; q < 6; ++q) out12[q] = h[q];
  for (int q = 0; q < 3; ++q) out12[6 + q] = h[q] + h[3 + q];           // force = pressure + viscous part (:1151-1153)
  const double den = s->cfg.density * (s->cfg.inlet_velocity * s->cfg.inlet_velocity) * radius;
  out12[9] = out12[6] / den;                                            // Cd(x) (:1155)
  out12[10] = out12[7] / den;                                           // Cl    (:1156)
  out12[11] = out12[8] / den;                                           // Cd(z) (:1157)
  PF_API_END(s)
}

size_t pf_vtk_section_bytes(const pf_solver *s, int section, int nplanes) {
  if (!s || !pf_vtk_section_valid(s->g, section)) return 0;
  const size_t planes = s->g.dim == 3 ? (size_t)std::max(nplanes, 0) : 1;
  return (size_t)s->g.m * s->g.n * planes * pf_vtk_record_bytes(section);
}

int pf_vtk_section(pf_solver *s, int section, int k_local0, int nplanes, const double *xp, const double *yp,
                   const double *zp, char *out) {
  PF_API_BEGIN(s)
  const Geo &g = s->g;
  if (!pf_vtk_section_valid(g, section)) throw std::string("pf_vtk_section: no such section for this case");
  if (!xp || !yp || !out || (g.dim == 3 && !zp)) throw std::string("pf_vtk_section: null argument");
  if (g.dim == 3 && (k_local0 < 1 || nplanes < 1 || k_local0 + nplanes - 1 > g.lz))
    throw std::string("pf_vtk_section: plane range outside this rank's slab");
  if (!s->porosity_set) throw std::string("pf_set_porosity must be called first");
  const size_t nx = (size_t)g.m + 2, ny = (size_t)g.n + 2, nz = g.dim == 3 ? (size_t)g.l + 2 : 0;
  if (!s->coord_dev) s->coord_dev = dalloc(s, (long long)(nx + ny + nz));
  PF_CUDA_OK(cudaMemcpyAsync(s->coord_dev, xp, nx * sizeof(double), cudaMemcpyHostToDevice, s->st));
  PF_CUDA_OK(cudaMemcpyAsync(s->coord_dev + nx, yp, ny * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (nz) PF_CUDA_OK(cudaMemcpyAsync(s->coord_dev + nx + ny, zp, nz * sizeof(double), cudaMemcpyHostToDevice, s->st));
  const size_t bytes = pf_vtk_section_bytes(s, section, nplanes);
  if (bytes > s->text_cap) {
    if (s->text_dev) { PF_CUDA_OK(cudaStreamSynchronize(s->st)); cudaFree(s->text_dev); s->text_dev = nullptr; s->text_cap = 0; }
    void *p = nullptr;
    PF_CUDA_OK(cudaMalloc(&p, bytes));
    s->text_dev = static_cast<char *>(p);
    s->text_cap = bytes;
  }
  k_vtk_section(g, s->f, section, k_local0, nplanes, s->coord_dev, s->coord_dev + nx,
                nz ? s->coord_dev + nx + ny : nullptr, s->cfg.inlet_velocity, s->text_dev, s->st);
  PF_CUDA_OK(cudaGetLastError());
  PF_CUDA_OK(cudaMemcpyAsync(out, s->text_dev, bytes, cudaMemcpyDeviceToHost, s->st));
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_sync(pf_solver *s) {
  PF_API_BEGIN(s)
  PF_CUDA_OK(cudaStreamSynchronize(s->st));
  PF_API_END(s)
}

int pf_last_timing(const pf_solver *s, double *ms_total, double *ms_sor, long long *launches) {
  if (!s) return 1;
  if (ms_total) *ms_total = s->ms_total;
  if (ms_sor) *ms_sor = s->ms_sor;
  if (launches) *launches = s->launches;
  return 0;
}

int pf_get_sor_variant(const pf_solver *s) { return s ? s->cfg.sor_variant : -1; }

int pf_get_halo_transport(const pf_solver *s) {
  if (!s) return -1;
  if (s->nranks == 1) return 0;
  if (!s->peer) return 1;
  return s->fused.sync ? 3 : 2;
}

void *pf_stream(const pf_solver *s) { return s ? (void *)s->st : nullptr; }

}  // extern "C"
