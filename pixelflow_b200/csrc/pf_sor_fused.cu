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
// Applicability: 3D uniform case, n and l even (periodic images keep their colour), n,l >= 4,
// single rank (the z-slab version exchanges two planes per iteration; not in this round).
#include "pf_internal.cuh"

namespace {

constexpr int FTX = 64;   // threads (= red/black elements) per tile row, incl. one overlap column
constexpr int FTY = 8;    // thread rows per block
constexpr double SMALLC = 1.e-6;

struct Fused {
  int NY2, NZ2;            // rows n+4, planes lz+4
  long long hplane2;       // HX*NY2
  const double *cx[2], *cy[2], *cz[2], *bb[2];
  const double *pin[2];
  double *pout[2];
  int cz_planes;           // owned planes per z-chunk
};

__device__ __forceinline__ long long row2(const Geo &g, const Fused &F, int j, int kl) {
  return (long long)g.H0 + (long long)g.HX * ((j + 1) + (long long)F.NY2 * (kl + 1));
}
__device__ __forceinline__ double ldg(const double *p) { return __ldg(p); }

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

// stores v at (ih, j, kl) of `dst` and at its periodic images in the depth-2 ghost rows / planes
__device__ __forceinline__ void store_with_images(const Geo &g, const Fused &F, double *dst, int ih, int j,
                                                  int kl, double v) {
  constexpr int NONE = -1000;
  const int n = g.n, lz = g.lz;
  const int j2 = (j <= 2) ? j + n : ((j >= n - 1) ? j - n : NONE);
  const int k2 = (kl <= 2) ? kl + lz : ((kl >= lz - 1) ? kl - lz : NONE);
  dst[row2(g, F, j, kl) + ih] = v;
  if (j2 != NONE) dst[row2(g, F, j2, kl) + ih] = v;
  if (k2 != NONE) {
    dst[row2(g, F, j, k2) + ih] = v;
    if (j2 != NONE) dst[row2(g, F, j2, k2) + ih] = v;
  }
}

template <int RPT>
__global__ void __launch_bounds__(FTX *FTY) sor_fused_kernel(Geo g, Fused F, double relux,
                                                             unsigned long long *err_bits) {
  constexpr int TJ = FTY * RPT;                      // ext rows per tile (ring included)
  __shared__ double R[3][TJ][FTX + 1];               // new red values of planes k, k-1, k-2 (mod 3)
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int h0 = (int)blockIdx.x * (FTX - 1) - 1;    // first red element of the tile (stride 63, overlap 1)
  const int j0 = (int)blockIdx.y * (TJ - 2);         // first ext row (ring row); owned rows j0+1 .. j0+TJ-2
  const int kc0 = (int)blockIdx.z * F.cz_planes + 1; // owned planes kc0 .. kc1
  const int kc1 = min(kc0 + F.cz_planes - 1, g.lz);
  const int ih = h0 + tx;
  const int ihmax = (g.m + 1) >> 1;                  // slot of the x-halo i=m+1 (or one past the last cell)
  const bool col_ok = ih >= -1 && ih <= ihmax;
  const double omr = 1. - relux;
  const int m = g.m;
  // per-row state carried along z (registers)
  double pbo[RPT][3];   // old black p of this (ih,j) at planes k-1, k, k+1
  double rn[RPT][3];    // new red of this (ih,j) at planes k-2, k-1, k
  double czr[RPT][3];   // cz of the red array at planes k-2, k-1, k
  int jrow[RPT];
  bool row_ok[RPT], own_row[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int tj = ty + r * FTY;
    jrow[r] = j0 + tj;
    row_ok[r] = jrow[r] <= g.n + 1;                          // rows 0..n+1 carry red values
    own_row[r] = tj >= 1 && tj <= TJ - 2 && jrow[r] >= 1 && jrow[r] <= g.n;
    for (int q = 0; q < 3; ++q) { pbo[r][q] = 0.; rn[r][q] = 0.; czr[r][q] = 0.; }
    if (col_ok && row_ok[r]) {
      pbo[r][1] = ldg(F.pin[1] + row2(g, F, jrow[r], kc0 - 2) + ih);
      pbo[r][2] = ldg(F.pin[1] + row2(g, F, jrow[r], kc0 - 1) + ih);
    }
  }
  double emax = 0.0;
  // owned columns of the tile: tx in [0, FTX-2] own the red stores; black ownership depends on parity
  for (int k = kc0 - 1; k <= kc1 + 1; ++k) {
    const int slot = (k + 3) % 3;
    // ------------------------------ red stage, plane k ------------------------------
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int tj = ty + r * FTY;
      const int j = jrow[r];
      pbo[r][0] = pbo[r][1]; pbo[r][1] = pbo[r][2];
      rn[r][0] = rn[r][1]; rn[r][1] = rn[r][2];
      czr[r][0] = czr[r][1]; czr[r][1] = czr[r][2];
      double val = 0.;
      if (col_ok && row_ok[r]) {
        const long long c = row2(g, F, j, k) + ih;
        pbo[r][2] = ldg(F.pin[1] + c + F.hplane2);           // old black, plane k+1
        const int s = (j + k + g.koff) & 1;                  // parity of i in the red row (colour 0)
        const int i = 2 * ih + 2 - s;
        const double pold = ldg(F.pin[0] + c);
        val = pold;
        if (i >= 1 && i <= m) {
          const long long cw = s ? c - 1 : c;                // west neighbour slot in the black array
          const long long ce = s ? c : c + 1;                // east neighbour slot
          const double pW = s ? ldg(F.pin[1] + cw) : pbo[r][1];
          const double pE = s ? pbo[r][1] : ldg(F.pin[1] + ce);
          const double ae = ldg(F.cx[0] + c), aw = ldg(F.cx[1] + cw);
          const double an = ldg(F.cy[0] + c), as = ldg(F.cy[1] + c - g.HX);
          const double at = ldg(F.cz[0] + c), ab = ldg(F.cz[1] + c - F.hplane2);
          czr[r][2] = at;
          const double pN = ldg(F.pin[1] + c + g.HX), pS = ldg(F.pin[1] + c - g.HX);
          val = sor_update(ldg(F.bb[0] + c), ae, aw, an, as, at, ab, pE, pW, pN, pS, pbo[r][2], pbo[r][0], pold,
                           relux, omr, i, m);
          if (tx <= FTX - 2 && own_row[r] && k >= kc0 && k <= kc1) store_with_images(g, F, F.pout[0], ih, j, k, val);
        } else {
          czr[r][2] = 0.;
        }
      }
      rn[r][2] = val;
      R[slot][tj][tx] = val;
    }
    __syncthreads();
    // ------------------------------ black stage, plane k-1 --------------------------
    const int kb = k - 1;
    if (kb >= kc0 && kb <= kc1) {
      const int sb_slot = (kb + 3) % 3;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int tj = ty + r * FTY;
        const int j = jrow[r];
        if (!(col_ok && own_row[r])) continue;
        const int s = (1 + j + kb + g.koff) & 1;             // parity of i in the black row (colour 1)
        const int i = 2 * ih + 2 - s;
        // black ownership inside the tile: needs both x neighbours among the tile's red elements
        const bool own_col = s ? (tx >= 1) : (tx <= FTX - 2);
        if (!own_col || i < 1 || i > m) continue;
        const long long c = row2(g, F, j, kb) + ih;
        const double pold = pbo[r][0];
        const double pW = s ? R[sb_slot][tj][tx - 1] : R[sb_slot][tj][tx];
        const double pE = s ? R[sb_slot][tj][tx] : R[sb_slot][tj][tx + 1];
        const double pN = R[sb_slot][tj + 1][tx], pS = R[sb_slot][tj - 1][tx];
        const long long cw = s ? c - 1 : c;
        const double ae = ldg(F.cx[1] + c), aw = ldg(F.cx[0] + cw);
        const double an = ldg(F.cy[1] + c), as = ldg(F.cy[0] + c - g.HX);
        const double at = ldg(F.cz[1] + c), ab = czr[r][0];
        const double v = sor_update(ldg(F.bb[1] + c), ae, aw, an, as, at, ab, pE, pW, pN, pS, rn[r][2], rn[r][0],
                                    pold, relux, omr, i, m);
        store_with_images(g, F, F.pout[1], ih, j, kb, v);
        emax = fmax(emax, fabs(v - pold));
      }
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
  const int jn = wrap1(jw + 1, g.n), kt = wrap1(kw + 1, g.lz);
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

}  // namespace

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
bool pf_fused_applicable(const Geo &g, const Phys &ph, int nranks) {
  return g.dim == 3 && ph.scase == PF_IBM3_UNIFORM && nranks == 1 && (g.n % 2 == 0) && (g.l % 2 == 0) &&
         g.n >= 4 && g.l >= 4 && g.lz == g.l;
}

long long pf_fused_elems(const Geo &g) { return (long long)g.HX * (g.n + 4) * (g.lz + 4); }

static Fused make_fused(const Geo &g, const FusedArrays &A, int in) {
  Fused F;
  F.NY2 = g.n + 4;
  F.NZ2 = g.lz + 4;
  F.hplane2 = (long long)g.HX * F.NY2;
  for (int c = 0; c < 2; ++c) {
    F.cx[c] = A.cx[c]; F.cy[c] = A.cy[c]; F.cz[c] = A.cz[c]; F.bb[c] = A.bb[c];
    F.pin[c] = A.p[in][c];
    F.pout[c] = A.p[in ^ 1][c];
  }
  F.cz_planes = A.cz_planes;
  return F;
}

void k_fused_build_faces(const Geo &g, const Phys &ph, const double *eps_nat, FusedArrays &A, cudaStream_t st) {
  // z-chunk size: enough blocks for >= ~3 waves of 148 SMs, chunks no thinner than 8 planes
  const int xt = ((g.m + 1) / 2 + 2 + (FTX - 2)) / (FTX - 1);
  const int yt = (g.n + (FTY * A.rpt - 2) - 1) / (FTY * A.rpt - 2);
  int cz = g.lz;
  while (cz > 8 && (long long)xt * yt * ((g.lz + cz - 1) / cz) < 3 * 148) cz = (cz + 1) / 2;
  A.cz_planes = cz;
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

// one red-black iteration: reads A.p[in], writes A.p[in^1]
void k_fused_iteration(const Geo &g, const Phys &ph, const FusedArrays &A, int in, unsigned long long *err_bits,
                       cudaStream_t st) {
  const Fused F = make_fused(g, A, in);
  const int xt = ((g.m + 1) / 2 + 2 + (FTX - 2)) / (FTX - 1);
  const dim3 block(FTX, FTY, 1);
  if (A.rpt == 2) {
    const int yt = (g.n + (FTY * 2 - 2) - 1) / (FTY * 2 - 2);
    sor_fused_kernel<2><<<dim3(xt, yt, (g.lz + F.cz_planes - 1) / F.cz_planes), block, 0, st>>>(g, F, ph.relux, err_bits);
  } else {
    const int yt = (g.n + (FTY - 2) - 1) / (FTY - 2);
    sor_fused_kernel<1><<<dim3(xt, yt, (g.lz + F.cz_planes - 1) / F.cz_planes), block, 0, st>>>(g, F, ph.relux, err_bits);
  }
  pf_count_launch();
}
