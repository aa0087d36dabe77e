// pf_sor_fused.cu -- SOR variant 3: one kernel launch per red-black ITERATION (both colours fused),
// face-shared coefficients, 48 instead of 88 bytes of HBM traffic per cell per sweep.
//
// Same arithmetic as pf_sor.cu (reference: solve_matrix_vec_omp, ibm_3d_uniform_omp_cpu.f90:433-614,
// update :510-515, coefficients :390-402, boundrary_matrix :636-658) -- bit-identical results.
//
// Where the bytes go away
//   * aw(i) == ae(i-1), as(j) == an(j-1), ab(k) == at(k-1) bit for bit (the sum of two porosities is
//     commutative), and ap is a fixed sum of the six raw values.  So three FACE arrays (cx, cy, cz)
//     replace seven per-cell coefficient arrays; the inlet/outlet folds depend only on i and are
//     applied in registers.
//   * red and black are updated in ONE pass: a block streams along z, updates the red cells of plane
//     k from the old pressure, parks them in shared memory, then updates the black cells of plane
//     k-1 from the new red values.  Every face coefficient, bb and p element is then read from HBM
//     once per iteration instead of once per half-sweep: 3 faces + bb + p read + p written = 48 B/cell.
//   * blocks are independent (no inter-block sync): each block recomputes the red values on a
//     one-cell ring around its tile (rows, columns and the two planes bounding its z-chunk).  That
//     needs old values two cells out, hence depth-2 ghost rows/planes ("split2" layout) and a
//     ping-pong pair of pressure buffers (read p_in, write p_out).  Periodic images of the written
//     cells are stored by the same threads, so no separate halo kernels run inside the solve.
//
// Applicability: 3D uniform case, n and l even (periodic images keep their colour), n >= 4, >= 4 planes per
// rank.  On z-slab ranks the plane images go to the NEIGHBOUR's ghost planes instead of this array's: the
// kernel stores them straight into the neighbour's memory over NVLink (CUDA IPC mapping, pf_comm.cu) and the
// ranks meet at a flag barrier between iterations; without peer mapping the planes travel by NCCL.
#include "pf_internal.cuh"

namespace {

// tile shape = thread block shape: FTX elements per tile row (incl. one overlap column), FTY rows
// (incl. the two ring rows).  Two shapes are instantiated: 32x8 (2 blocks/SM) and 32x16.
constexpr double SMALLC = 1.e-6;

struct Fused {
  int NY2, NZ2;            // rows n+4, planes lz+4
  int hplane2;             // HX*NY2  (all element indices fit 32 bits: checked in pf_fused_applicable)
  const double *cx[2], *cy[2], *cz[2], *bb[2];
  const double *pin[2];
  double *pout[2];
  double *ilo[2], *ihi[2]; // image destinations of planes 1,2 / lz-1,lz (FusedArrays::img_lo / img_hi), may be null
  int dk_lo, dk_hi;        // their element offsets
  int cz_planes;           // owned planes per z-chunk
};

// 32-bit element index: one IMAD.WIDE per address instead of a 64-bit add chain
__device__ __forceinline__ int row2(const Geo &g, const Fused &F, int j, int kl) {
  return g.H0 + g.HX * ((j + 1) + F.NY2 * (kl + 1));
}
__device__ __forceinline__ double ldg(const double *p, int idx) { return __ldg(p + idx); }

// one SOR update, the reference's expression order (:510-515)
__device__ __forceinline__ double sor_update(double bb, double ae, double aw, double an, double as, double at,
                                             double ab, double pE, double pW, double pN, double pS, double pT,
                                             double pB, double pold, double relux, double omr, int i, int m) {
  const double ap = -ae - aw - an - as - at - ab;   // :402, from the raw coefficients
  if (i == 1) { ae = ae + aw; aw = 0.; }            // :640-641
  if (i == m) { ae = aw = an = as = at = ab = 0.; } // :651-656
  const double r = bb - ae * pE - aw * pW - an * pN - as * pS - at * pT - ab * pB;
  return r / ap * relux + pold * omr;
}

// stores v at element c of `dst` and at its images: the periodic row image in the depth-2 ghost rows
// (dj = index offset, 0 = none) and the plane image in `img` (this array's ghost planes on one rank, the
// neighbour rank's ghost planes on a z-slab; null = none) at offset dk
__device__ __forceinline__ void store_with_images(double *dst, double *img, int c, int dj, int dk, double v) {
  dst[c] = v;
  if (dj) dst[c + dj] = v;
  if (img) {
    img[c + dk] = v;
    if (dj) img[c + dk + dj] = v;
  }
}

// operands of one red update / one black update that come from global memory; they are loaded ONE
// z-step ahead of their use (software pipelining: the kernel is a long dependent chain of tiny steps,
// and without the prefetch every step exposes a full DRAM latency twice)
struct RedIn { double pold, px, pN, pS, pT, ae, aw, an, as, at, ab, bb; };
struct BlkIn { double ae, aw, an, as, bb; };   // at(k-1) of the black cell == ab(k) of the red cell above it

__device__ __forceinline__ RedIn load_red(const Geo &g, const Fused &F, int c, int ih, int s, bool active) {
  RedIn q;
  q.pold = q.px = q.pN = q.pS = q.pT = q.ae = q.aw = q.an = q.as = q.at = q.ab = q.bb = 0.;
  if (!active) return q;
  q.pT = ldg(F.pin[1], c + F.hplane2);                       // old black, plane k+1
  q.pold = ldg(F.pin[0], c);
  const int i = 2 * ih + 2 - s;                              // s = parity of i in this red row
  if (i >= 1 && i <= g.m) {
    const int cw = c - s;                                    // west neighbour slot in the black array
    q.px = ldg(F.pin[1], c + 1 - 2 * s);                     // the x neighbour that is not (ih,j,k) itself
    q.pN = ldg(F.pin[1], c + g.HX);
    q.pS = ldg(F.pin[1], c - g.HX);
    q.ae = ldg(F.cx[0], c);
    q.aw = ldg(F.cx[1], cw);
    q.an = ldg(F.cy[0], c);
    q.as = ldg(F.cy[1], c - g.HX);
    q.at = ldg(F.cz[0], c);
    q.ab = ldg(F.cz[1], c - F.hplane2);
    q.bb = ldg(F.bb[0], c);
  }
  return q;
}

__device__ __forceinline__ BlkIn load_blk(const Geo &g, const Fused &F, int c, int ih, int s, bool active) {
  BlkIn q;
  q.ae = q.aw = q.an = q.as = q.bb = 0.;
  const int i = 2 * ih + 2 - s;                              // s = parity of i in this black row
  if (!active || i < 1 || i > g.m) return q;
  q.ae = ldg(F.cx[1], c);
  q.aw = ldg(F.cx[0], c - s);
  q.an = ldg(F.cy[1], c);
  q.as = ldg(F.cy[0], c - g.HX);
  q.bb = ldg(F.bb[1], c);
  return q;
}

template <int FTX, int FTY>
__global__ void __launch_bounds__(FTX *FTY, (FTX * FTY <= 256) ? 2 : 1) sor_fused_kernel(Geo g, Fused F, double relux,
                                                             unsigned long long *err_bits) {
  __shared__ double R[3][FTY][FTX + 1];              // new red values of planes k, k-1, k-2 (mod 3)
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int h0 = (int)blockIdx.x * (FTX - 1) - 1;    // first red element of the tile (stride FTX-1, overlap 1)
  const int j0 = (int)blockIdx.y * (FTY - 2);        // first ext row (ring row); owned rows j0+1 .. j0+FTY-2
  const int kc0 = (int)blockIdx.z * F.cz_planes + 1; // owned planes kc0 .. kc1
  const int kc1 = min(kc0 + F.cz_planes - 1, g.lz);
  const int ih = h0 + tx;
  const int ihmax = (g.m + 1) >> 1;                  // slot of the x-halo i=m+1 (or one past the last cell)
  const int j = j0 + ty;
  const bool act = ih >= -1 && ih <= ihmax && j <= g.n + 1;           // rows 0..n+1 carry red values
  const bool own_row = act && ty >= 1 && ty <= FTY - 2 && j >= 1 && j <= g.n;
  const double omr = 1. - relux;
  const int m = g.m;
  // periodic row image of this thread's row (index offset, 0 = none); plane images are per step
  const int dj = (j <= 2) ? g.n * g.HX : ((j >= g.n - 1) ? -g.n * g.HX : 0);
  const int sj = (j + g.koff) & 1;                   // row part of the parity
  // state carried along z (registers)
  double pb0 = 0., pb1 = 0., pb2 = 0.;   // old black p of this (ih,j) at planes k-1, k, k+1
  double rn0 = 0., rn1 = 0., rn2 = 0.;   // new red of this (ih,j) at planes k-2, k-1, k
  double cz0 = 0., cz1 = 0., cz2 = 0.;   // cz of the red array at planes k-2, k-1, k
  int c = row2(g, F, j, kc0 - 1) + ih;   // element index at plane k (advanced by hplane2 per step)
  if (act) {
    pb1 = ldg(F.pin[1], c - F.hplane2);
    pb2 = ldg(F.pin[1], c);
  }
  RedIn rnext = load_red(g, F, c, ih, (sj + kc0 - 1) & 1, act);
  BlkIn bnext = load_blk(g, F, c, ih, 0, false);
  double emax = 0.0;
#pragma unroll 1
  for (int k = kc0 - 1; k <= kc1 + 1; ++k, c += F.hplane2) {
    const int slot = (k + 3) % 3;
    const int s = (sj + k) & 1;                      // parity of i: red row at plane k == black row at plane k-1
    const RedIn rc = rnext;
    const BlkIn bc = bnext;
    // prefetch: red operands of plane k+1, black operands of plane k (both used in the next step)
    rnext = load_red(g, F, c + F.hplane2, ih, s ^ 1, act && k + 1 <= kc1 + 1);
    bnext = load_blk(g, F, c, ih, s ^ 1, own_row && k >= kc0 && k <= kc1);
    // ------------------------------ red stage, plane k ------------------------------
    pb0 = pb1; pb1 = pb2; pb2 = rc.pT;
    rn0 = rn1; rn1 = rn2;
    cz0 = cz1; cz1 = cz2; cz2 = rc.at;
    const int i = 2 * ih + 2 - s;
    const bool cell = i >= 1 && i <= m;
    double val = rc.pold;
    if (act && cell) {
      const double pW = s ? rc.px : pb1;
      const double pE = s ? pb1 : rc.px;
      val = sor_update(rc.bb, rc.ae, rc.aw, rc.an, rc.as, rc.at, rc.ab, pE, pW, rc.pN, rc.pS, pb2, pb0, rc.pold,
                       relux, omr, i, m);
      if (tx <= FTX - 2 && own_row && k >= kc0 && k <= kc1) {
        const bool lo = k <= 2, hi = k >= g.lz - 1;
        store_with_images(F.pout[0], lo ? F.ilo[0] : (hi ? F.ihi[0] : nullptr), c, dj, lo ? F.dk_lo : F.dk_hi, val);
      }
    }
    rn2 = val;
    R[slot][ty][tx] = val;
    __syncthreads();
    // ------------------------------ black stage, plane k-1 --------------------------
    const int kb = k - 1;
    // black ownership inside the tile: needs both x neighbours among the tile's red elements
    const bool own_col = s ? (tx >= 1) : (tx <= FTX - 2);
    if (kb >= kc0 && kb <= kc1 && own_row && own_col && cell) {
      const int sb = (kb + 3) % 3;
      const double pold = pb0;
      const double pW = R[sb][ty][tx - s];
      const double pE = R[sb][ty][tx + 1 - s];
      const double pN = R[sb][ty + 1][tx], pS = R[sb][ty - 1][tx];
      // at of the black cell (ih,j,kb) is the ab the red cell above it just used; ab is cz_red(k-2)
      const double v = sor_update(bc.bb, bc.ae, bc.aw, bc.an, bc.as, rc.ab, cz0, pE, pW, pN, pS, rn2, rn0, pold,
                                  relux, omr, i, m);
      const bool lo = kb <= 2, hi = kb >= g.lz - 1;
      store_with_images(F.pout[1], lo ? F.ilo[1] : (hi ? F.ihi[1] : nullptr), c - F.hplane2, dj,
                        lo ? F.dk_lo : F.dk_hi, v);
      emax = fmax(emax, fabs(v - pold));
    }
  }
  // running max of |p - p_old| over the black cells (:575-583)
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  __shared__ double wmax[FTX * FTY / 32];
  const int tid = ty * FTX + tx;
  __syncthreads();
  if ((tid & 31) == 0) wmax[tid >> 5] = emax;
  __syncthreads();
  if (tid < 32) {
    double v = (tid < FTX * FTY / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

// ---- set-up kernels ----------------------------------------------------------------------------------
__device__ __forceinline__ int wrap1(int a, int n) { return a < 1 ? a + n : (a > n ? a - n : a); }

// face coefficients from the natural porosity, on every slot of the split2 arrays (ghosts = periodic images)
__global__ void fused_faces_kernel(Geo g, Phys ph, Fused F, const double *eps, double *cx0, double *cx1,
                                   double *cy0, double *cy1, double *cz0, double *cz1) {
  const int ih = blockIdx.x * blockDim.x + threadIdx.x - 1;
  const int j = (int)blockIdx.y - 1;
  const int kl = (int)blockIdx.z - 1;
  if (ih > ((g.m + 1) >> 1)) return;
  const int jw = wrap1(j, g.n), kw = wrap1(kl, g.lz);
  // the plane above comes from the porosity's own ghost plane: the periodic image on one rank (lib/grid.f90:
  // 349-378 fills it), the neighbour slab's first plane on a z-slab
  const int jn = wrap1(jw + 1, g.n), kt = kw + 1;
  for (int c = 0; c < 2; ++c) {
    const int s = (c + j + kl + g.koff) & 1;
    const int i = 2 * ih + 2 - s;
    if (i < 0 || i > g.m) continue;
    const long long d = row2(g, F, j, kl) + ih;
    const double e0 = eps[nat_idx(g, i, jw, kw)];
    (c ? cx1 : cx0)[d] = ph.dt * fmax(SMALLC, (eps[nat_idx(g, i + 1, jw, kw)] + e0) * 0.5) / ph.ix / ph.ix;
    if (i >= 1) {
      (c ? cy1 : cy0)[d] = ph.dt * fmax(SMALLC, (eps[nat_idx(g, i, jn, kw)] + e0) * 0.5) / ph.iy / ph.iy;
      (c ? cz1 : cz0)[d] = ph.dt * fmax(SMALLC, (eps[nat_idx(g, i, jw, kt)] + e0) * 0.5) / ph.iz / ph.iz;
    }
  }
}

// split (ghost depth 1) -> split2 (ghost depth 2): every slot of every row/plane, ghosts taken from the
// periodic image of the interior; the whole row is copied, x-halo slots included
__global__ void fused_gather_kernel(Geo g, Fused F, const double *s0, const double *s1, double *d0, double *d1) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;   // raw column 0..HX-1
  const int j = (int)blockIdx.y - 1;
  const int kl = (int)blockIdx.z - 1;
  if (col >= g.HX) return;
  const int jw = wrap1(j, g.n), kw = wrap1(kl, g.lz);
  const long long src = (long long)g.HX * (jw + (long long)g.NY * kw) + col;
  const long long dst = (long long)g.HX * ((j + 1) + (long long)F.NY2 * (kl + 1)) + col;
  d0[dst] = s0[src];
  d1[dst] = s1[src];
}

// split2 interior -> split (ghost depth 1) interior rows/planes (whole rows)
__global__ void fused_scatter_kernel(Geo g, Fused F, const double *s0, const double *s1, double *d0, double *d1) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)blockIdx.y + 1;
  const int kl = (int)blockIdx.z + 1;
  if (col >= g.HX) return;
  // only the interior cells i=1..m of each colour: the x-halo slots of the destination keep their values
  const long long src = (long long)g.HX * ((j + 1) + (long long)F.NY2 * (kl + 1)) + col;
  const long long dst = (long long)g.HX * (j + (long long)g.NY * kl) + col;
  const int ih = col - g.H0;
  for (int c = 0; c < 2; ++c) {
    const int s = (c + j + kl + g.koff) & 1;
    const int i = 2 * ih + 2 - s;
    if (i >= 1 && i <= g.m) (c ? d1 : d0)[dst] = (c ? s1 : s0)[src];
  }
}

// One rank: natural pressure (with its x-halo columns) -> BOTH split2 pressure buffers, every slot of every row and
// plane, ghost rows / planes = the periodic images of the interior -- what convert (natural -> split) followed by two
// gathers wrote, in one pass over the natural array.
__global__ void fused_gather_nat_kernel(Geo g, Fused F, const double *__restrict__ nat, double *a0, double *a1,
                                        double *b0, double *b1) {
  const int ih = blockIdx.x * blockDim.x + threadIdx.x - 1;
  const int j = (int)blockIdx.y - 1;
  const int kl = (int)blockIdx.z - 1;
  if (ih > ((g.m + 1) >> 1)) return;
  const int jw = wrap1(j, g.n), kw = wrap1(kl, g.lz);
  const long long d = row2(g, F, j, kl) + ih;
  for (int c = 0; c < 2; ++c) {
    const int s = (c + j + kl + g.koff) & 1;
    const int i = 2 * ih + 2 - s;
    if (i < 0 || i > g.m + 1) continue;
    const double v = nat[nat_idx(g, i, jw, kw)];
    (c ? a1 : a0)[d] = v;
    (c ? b1 : b0)[d] = v;
  }
}

// One rank: the final split2 pressure -> natural: the interior, the periodic rows j = 0, n+1 (i = 1..m, k = 1..lz) and
// the periodic planes k = 0, lz+1 (i = 1..m, j = 1..n) -- exactly the cells that scatter + the closing halo refresh
// (:588-605) + convert (split -> natural) changed; x-halo columns and the edges of the halo shell keep their values.
__global__ void fused_scatter_nat_kernel(Geo g, Fused F, const double *__restrict__ s0, const double *__restrict__ s1,
                                         double *nat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = (int)blockIdx.y;
  const int k = (int)blockIdx.z;
  if (i > g.m) return;
  const bool jh = j == 0 || j == g.n + 1, kh = k == 0 || k == g.lz + 1;
  if (jh && kh) return;
  const int jw = wrap1(j, g.n), kw = wrap1(k, g.lz);
  const int c = (i + jw + kw + g.koff) & 1;
  const long long src = row2(g, F, jw, kw) + (((i + 1) >> 1) - 1);
  nat[nat_idx(g, i, j, k)] = (c ? s1 : s0)[src];
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
long long pf_fused_elems(const Geo &g);
// the same answer on every rank of a run: only global sizes enter (slabs differ by at most one plane)
bool pf_fused_applicable(const Geo &g, const Phys &ph, int nranks) {
  const int lz_min = g.l / nranks, lz_max = lz_min + (g.l % nranks ? 1 : 0);
  const long long elems_max = (long long)g.HX * (g.n + 4) * (lz_max + 4);
  return g.dim == 3 && ph.scase == PF_IBM3_UNIFORM && (g.n % 2 == 0) && (g.l % 2 == 0) && g.n >= 4 && lz_min >= 4 &&
         elems_max < (1ll << 31) - (1ll << 20);
}

// elements of one split2 array, padded so that arrays packed back to back stay 256-byte aligned
long long pf_fused_elems(const Geo &g) { return ((long long)g.HX * (g.n + 4) * (g.lz + 4) + 31) / 32 * 32; }

static Fused make_fused(const Geo &g, const FusedArrays &A, int in) {
  Fused F;
  F.NY2 = g.n + 4;
  F.NZ2 = g.lz + 4;
  F.hplane2 = g.HX * F.NY2;
  for (int c = 0; c < 2; ++c) {
    F.cx[c] = A.cx[c]; F.cy[c] = A.cy[c]; F.cz[c] = A.cz[c]; F.bb[c] = A.bb[c];
    F.pin[c] = A.p[in][c];
    F.pout[c] = A.p[in ^ 1][c];
    F.ilo[c] = A.img_lo[in ^ 1][c];
    F.ihi[c] = A.img_hi[in ^ 1][c];
  }
  F.dk_lo = (int)A.dk_lo;
  F.dk_hi = (int)A.dk_hi;
  F.cz_planes = A.cz_planes;
  return F;
}

void k_fused_build_faces(const Geo &g, const Phys &ph, const double *eps_nat, FusedArrays &A, cudaStream_t st) {
  // z-chunk size: whole waves of blocks (SMs x resident blocks), chunks no thinner than 16 planes
  const int FTX = 32, FTY = A.rpt == 2 ? 16 : 8, resident = A.rpt == 2 ? 1 : 2;
  const int xt = ((g.m + 1) / 2 + 2 + (FTX - 2)) / (FTX - 1);
  const int yt = (g.n + (FTY - 2) - 1) / (FTY - 2);
  int best = g.lz;
  double best_cost = 1e30;
  for (int cz = g.lz; cz >= 16 || cz == g.lz; --cz) {
    const long long blocks = (long long)xt * yt * ((g.lz + cz - 1) / cz);
    const long long slots = (long long)pf_sm_count() * resident, waves = (blocks + slots - 1) / slots;
    const double cost = (double)waves * (cz + 2);     // z-steps on the critical path
    if (cost < best_cost) { best_cost = cost; best = cz; }
    if (cz <= 16) break;
  }
  A.cz_planes = best;
  const Fused F = make_fused(g, A, 0);
  const int cols = (g.m + 1) / 2 + 2;
  fused_faces_kernel<<<dim3((cols + 63) / 64, g.n + 4, g.lz + 4), 64, 0, st>>>(g, ph, F, eps_nat, A.cx[0], A.cx[1],
                                                                               A.cy[0], A.cy[1], A.cz[0], A.cz[1]);
  pf_count_launch();
}

void k_fused_gather(const Geo &g, const FusedArrays &A, const double *s0, const double *s1, double *d0, double *d1,
                    cudaStream_t st) {
  const Fused F = make_fused(g, A, 0);
  fused_gather_kernel<<<dim3((g.HX + 127) / 128, g.n + 4, g.lz + 4), 128, 0, st>>>(g, F, s0, s1, d0, d1);
  pf_count_launch();
}

void k_fused_scatter(const Geo &g, const FusedArrays &A, const double *s0, const double *s1, double *d0, double *d1,
                     cudaStream_t st) {
  const Fused F = make_fused(g, A, 0);
  fused_scatter_kernel<<<dim3((g.HX + 127) / 128, g.n, g.lz), 128, 0, st>>>(g, F, s0, s1, d0, d1);
  pf_count_launch();
}

void k_fused_gather_nat(const Geo &g, const FusedArrays &A, const double *nat, cudaStream_t st) {
  const Fused F = make_fused(g, A, 0);
  const int cols = ((g.m + 1) >> 1) + 2;   // ih = -1 .. (m+1)/2
  fused_gather_nat_kernel<<<dim3((cols + 127) / 128, g.n + 4, g.lz + 4), 128, 0, st>>>(g, F, nat, A.p[0][0], A.p[0][1],
                                                                                      A.p[1][0], A.p[1][1]);
  pf_count_launch();
}

void k_fused_scatter_nat(const Geo &g, const FusedArrays &A, int fin, double *nat, cudaStream_t st) {
  const Fused F = make_fused(g, A, 0);
  fused_scatter_nat_kernel<<<dim3((g.m + 127) / 128, g.n + 2, g.lz + 2), 128, 0, st>>>(g, F, A.p[fin][0], A.p[fin][1], nat);
  pf_count_launch();
}

// one red-black iteration: reads A.p[in], writes A.p[in^1]
void k_fused_iteration(const Geo &g, const Phys &ph, const FusedArrays &A, int in, unsigned long long *err_bits,
                       cudaStream_t st) {
  const Fused F = make_fused(g, A, in);
  constexpr int FTX = 32;
  const int xt = ((g.m + 1) / 2 + 2 + (FTX - 2)) / (FTX - 1);
  const int zt = (g.lz + F.cz_planes - 1) / F.cz_planes;
  if (A.rpt == 2) {
    const int yt = (g.n + 14 - 1) / 14;
    sor_fused_kernel<FTX, 16><<<dim3(xt, yt, zt), dim3(FTX, 16, 1), 0, st>>>(g, F, ph.relux, err_bits);
  } else {
    const int yt = (g.n + 6 - 1) / 6;
    sor_fused_kernel<FTX, 8><<<dim3(xt, yt, zt), dim3(FTX, 8, 1), 0, st>>>(g, F, ph.relux, err_bits);
  }
  pf_count_launch();
}
