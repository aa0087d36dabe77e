// pf_sor.cu -- red-black SOR half-sweeps on the checkerboard layout.
//
// Reference: solve_matrix_vec_omp, src/omp_parallel/ibm_3d_uniform_omp_cpu.f90:433-614 (2D:
// ibm_2d_uniform_omp_cpu.f90:293-406, air-condition :480-661).  Per iteration the reference does
// {halo refresh; p_old=p; colour-1 sweep; halo refresh; p_old=p; colour-2 sweep; error pass}.  All six
// neighbours of a cell have the other colour and halos are only rewritten between half-sweeps, so
// an in-place half-sweep that reads separately stored halos is bit-identical to the p_old form
// (SURVEY.md 8a); the two full copies and the separate error pass are not executed here.
//
// Update (:510-515), evaluated left to right, no FMA:
//   p = (bb - ae*pE - aw*pW - an*pN - as*pS - at*pT - ab*pB) / ap * relux + p*(1 - relux)
//
// Every operand of a half-sweep is a unit-stride stream in the checkerboard layout (pf_internal.cuh):
// each thread updates two adjacent same-colour cells with 16-byte (double2) loads/stores.
// Algorithmic traffic: 8 coefficient doubles + own p (R+W) + the other colour's p (R, reused 6x
// through L1/L2) = 88 B per cell per full sweep.
#include "pf_internal.cuh"

namespace {

constexpr int SBX = 64, SBY = 4;   // 64 pair-threads cover 256 cells of a row (short rows: 32 x 8, 16 x 16, 8 x 32 -- always 256 threads)

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
// streaming (read-once) operands: bypass L1 allocation so the re-used p lines stay resident
__device__ __forceinline__ double2 ld2_stream(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ void block_max_to_global(double v, unsigned long long *err_bits) {
  // non-negative doubles order like their bit patterns -> integer atomicMax
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __shared__ double wmax[SBX * SBY / 32];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0) wmax[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    v = (tid < SBX * SBY / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

// variant 1: one thread = one (pair, j, k); grid covers the slab.
// YIMG (2D): the half-sweep keeps the periodic y-halo rows itself, so no halo kernel runs between half-sweeps.
//   1 (even n): the halo cell has the colour of its source, so the thread that updates a cell of row 1 / row n
//               also stores the new value in row n+1 / row 0 of the same array;
//   2 (odd n) : the halo rows of THIS colour's array are images of the OTHER colour's rows 1 and n, which do not
//               change during this launch and are not read through these halo rows until the other colour's
//               next half-sweep: the threads of rows 1 and n copy them across (the reference's refresh, :323-330,
//               moved from "before the reader" to "after the writer").
// EARLY (the launches of a dependent-launch chain): operands the previous launch does not write are loaded before the
//   grid dependency is awaited.  Costs registers (66 instead of 44 in 3D), so the bandwidth-bound launches on large
//   grids, which are not chained, keep the interleaved order.
template <int DIM, bool ERR, int YIMG = 0, bool EARLY = false>
__global__ void __launch_bounds__(SBX *SBY, EARLY ? 1 : 5) sor_sweep_kernel(Geo g, SplitSet S, const double *__restrict__ po,
                                                             int colour, double relux,
                                                             unsigned long long *err_bits, int k0, int kstride) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = (DIM == 3) ? (int)blockIdx.z * kstride + k0 : 0;
  double emax = 0.0;
  const int s = (colour + j + k + g.koff) & 1;        // parity of i in this row of this colour
  const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;      // interior cells of that parity
  const int ih = 2 * q;
  const bool active = j <= g.n && ih < cnt;
  const long long r = split_row(g, j, k) + ih;
  // What the previous launch on this stream does not write -- the coefficients, bb and this colour's own pressure
  // (last written two launches ago) -- is loaded BEFORE the grid dependency is awaited: launched with programmatic
  // stream serialization (k_sor_sweep, `pdl`), this kernel's blocks start while the previous half-sweep drains, and
  // only the other colour's pressure waits for it.  Without that launch attribute the two instructions do nothing.
  double2 ap, bb, ae, aw, an, as, at, ab, pc;
  ap = bb = ae = aw = an = as = at = ab = pc = make_double2(0., 0.);
  auto load_own = [&]() {
    ap = ld2_stream(S.ap + r); bb = ld2_stream(S.bb + r);
    ae = ld2_stream(S.ae + r); aw = ld2_stream(S.aw + r);
    an = ld2_stream(S.an + r); as = ld2_stream(S.as + r);
    if (DIM == 3) { at = ld2_stream(S.at + r); ab = ld2_stream(S.ab + r); }
    pc = ld2(S.p + r);
  };
  if (EARLY && active) load_own();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (j <= g.n) {
    if (active) {
      if (!EARLY) load_own();
      const double2 px = ld2(po + r);
      const double xtra = s ? po[r - 1] : po[r + 2];
      const double2 pn = ld2(po + r + g.HX), ps = ld2(po + r - g.HX);
      // west/east neighbours: s=1 -> {ih-1, ih}, s=0 -> {ih, ih+1}
      const double wa = s ? xtra : px.x, ea = s ? px.x : px.y;
      const double wb = s ? px.x : px.y, eb = s ? px.y : xtra;
      double ra = bb.x - ae.x * ea - aw.x * wa - an.x * pn.x - as.x * ps.x;
      double rb = bb.y - ae.y * eb - aw.y * wb - an.y * pn.y - as.y * ps.y;
      if (DIM == 3) {
        const double2 pt = ld2(po + r + g.hplane), pb = ld2(po + r - g.hplane);
        ra = ra - at.x * pt.x - ab.x * pb.x;
        rb = rb - at.y * pt.y - ab.y * pb.y;
      }
      const double omr = 1. - relux;
      double2 out;
      out.x = ra / ap.x * relux + pc.x * omr;
      out.y = (ih + 1 < cnt) ? rb / ap.y * relux + pc.y * omr : pc.y;
      *reinterpret_cast<double2 *>(S.p + r) = out;
      if (YIMG == 1 && (j == 1 || j == g.n)) {           // interior cells only, like the halo refresh (:323-330)
        double *img = S.p + r + (j == 1 ? (long long)g.n * g.HX : -(long long)g.n * g.HX);
        img[0] = out.x;
        if (ih + 1 < cnt) img[1] = out.y;
      }
      if (ERR) emax = fmax(fabs(out.x - pc.x), fabs(out.y - pc.y));
    }
    if (YIMG == 2 && (j == 1 || j == g.n)) {
      const int jh = (j == 1) ? g.n + 1 : 0;             // row n+1 <- row 1, row 0 <- row n
      const int sh = (colour + jh + k + g.koff) & 1;
      const int cnth = sh ? (g.m + 1) >> 1 : g.m >> 1;
      const long long dst = split_row(g, jh, k), src = split_row(g, j, k);
      if (ih < cnth) S.p[dst + ih] = po[src + ih];
      if (ih + 1 < cnth) S.p[dst + ih + 1] = po[src + ih + 1];
    }
  }
  if (ERR) block_max_to_global(emax, err_bits);
}

// variant 2 (3D uniform case): the seven Poisson coefficients are not streamed from HBM but
// recomputed from the checkerboard porosity with the reference's own expressions (:390-402) and
// the boundrary_matrix fold (:636-658) -- bit-identical values, 48 instead of 88 bytes per cell per
// sweep (own/other eps 16 B, bb 8 B, p 24 B), paid for with 12 exact reciprocal divisions per update.
template <bool ERR>
__global__ void __launch_bounds__(SBX *SBY) sor_sweep_eps_kernel(Geo g, Phys ph, SplitSet S,
                                                                 const double *__restrict__ po,
                                                                 const double *__restrict__ eo, int colour,
                                                                 unsigned long long *err_bits, int k0, int kstride) {
  constexpr double SMALL = 1.e-6;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = (int)blockIdx.z * kstride + k0;
  double emax = 0.0;
  if (j <= g.n) {
    const int s = (colour + j + k + g.koff) & 1;
    const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;
    const int ih = 2 * q;
    if (ih < cnt) {
      const long long r = split_row(g, j, k) + ih;
      const double2 bb = ld2_stream(S.bb + r);
      const double2 ec = ld2(S.eps + r);
      const double2 pc = ld2(S.p + r);
      const double2 px = ld2(po + r), ex = ld2(eo + r);
      const double pxtra = s ? po[r - 1] : po[r + 2];
      const double extra = s ? eo[r - 1] : eo[r + 2];
      const double2 pn = ld2(po + r + g.HX), ps = ld2(po + r - g.HX);
      const double2 en = ld2(eo + r + g.HX), es = ld2(eo + r - g.HX);
      const double2 pt = ld2(po + r + g.hplane), pb = ld2(po + r - g.hplane);
      const double2 et = ld2(eo + r + g.hplane), eb = ld2(eo + r - g.hplane);
      const double relux = ph.relux, omr = 1. - relux, dt = ph.dt;
      const Inv dx = ph.ix, dy = ph.iy, dz = ph.iz;
      double out[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const double e0 = c ? ec.y : ec.x, p0 = c ? pc.y : pc.x;
        // west/east neighbours: s=1 -> {ih-1, ih}, s=0 -> {ih, ih+1}
        const double eW = c ? (s ? ex.x : ex.y) : (s ? extra : ex.x);
        const double eE = c ? (s ? ex.y : extra) : (s ? ex.x : ex.y);
        const double pW = c ? (s ? px.x : px.y) : (s ? pxtra : px.x);
        const double pE = c ? (s ? px.y : pxtra) : (s ? px.x : px.y);
        const double eN = c ? en.y : en.x, eS = c ? es.y : es.x, eT = c ? et.y : et.x, eB = c ? eb.y : eb.x;
        const double pN = c ? pn.y : pn.x, pS = c ? ps.y : ps.x, pT = c ? pt.y : pt.x, pB = c ? pb.y : pb.x;
        double ae = dt * fmax(SMALL, (eE + e0) * 0.5) / dx / dx;
        double aw = dt * fmax(SMALL, (e0 + eW) * 0.5) / dx / dx;
        double an = dt * fmax(SMALL, (eN + e0) * 0.5) / dy / dy;
        double as = dt * fmax(SMALL, (e0 + eS) * 0.5) / dy / dy;
        double at = dt * fmax(SMALL, (eT + e0) * 0.5) / dz / dz;
        double ab = dt * fmax(SMALL, (e0 + eB) * 0.5) / dz / dz;
        const double ap = -ae - aw - an - as - at - ab;
        const int i = 2 * (ih + c) + 2 - s;
        if (i == 1) { ae = ae + aw; aw = 0.; }
        if (i == g.m) { ae = aw = an = as = at = ab = 0.; }
        const double bbv = c ? bb.y : bb.x;
        const double rr = bbv - ae * pE - aw * pW - an * pN - as * pS - at * pT - ab * pB;
        out[c] = rr / ap * relux + p0 * omr;
      }
      double2 o;
      o.x = out[0];
      o.y = (ih + 1 < cnt) ? out[1] : pc.y;
      *reinterpret_cast<double2 *>(S.p + r) = o;
      if (ERR) emax = fmax(fabs(o.x - pc.x), fabs(o.y - pc.y));
    }
  }
  if (ERR) block_max_to_global(emax, err_bits);
}

// periodic-y halo rows of the checkerboard p, i=1..m only (:463-470).  Cell (i,0,k) has colour
// (i+k)&1 and copies cell (i,n,k) of colour (i+n+k)&1: same column ih, colour flipped iff n is odd.
__global__ void sor_halo_y_kernel(Geo g, double *p0, double *p1, int colour_mask) {
  const int ih = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = (int)blockIdx.y + g.kin0;
  const int flip = g.n & 1;
#pragma unroll
  for (int cd = 0; cd < 2; ++cd) {
    if (!((colour_mask >> cd) & 1)) continue;
    double *dst = cd ? p1 : p0;
    const double *src = (cd ^ flip) ? p1 : p0;
    {  // row 0 <- row n
      const int s = (cd + 0 + k + g.koff) & 1;
      const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;
      if (ih < cnt) dst[split_row(g, 0, k) + ih] = src[split_row(g, g.n, k) + ih];
    }
    {  // row n+1 <- row 1
      const int s = (cd + g.n + 1 + k + g.koff) & 1;
      const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;
      if (ih < cnt) dst[split_row(g, g.n + 1, k) + ih] = src[split_row(g, 1, k) + ih];
    }
  }
}

// periodic-z halo planes on a single rank, i=1..m, j=1..n (:473-480): plane 0 <- plane l,
// plane l+1 <- plane 1; colour flipped iff l is odd.
__global__ void sor_halo_z_kernel(Geo g, double *p0, double *p1, int colour_mask) {
  const int ih = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)blockIdx.y + 1;
  const int flip = g.l & 1;
#pragma unroll
  for (int cd = 0; cd < 2; ++cd) {
    if (!((colour_mask >> cd) & 1)) continue;
    double *dst = cd ? p1 : p0;
    const double *src = (cd ^ flip) ? p1 : p0;
    {
      const int s = (cd + j + 0) & 1;
      const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;
      if (ih < cnt) dst[split_row(g, j, 0) + ih] = src[split_row(g, j, g.l) + ih];
    }
    {
      const int s = (cd + j + g.l + 1) & 1;
      const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;
      if (ih < cnt) dst[split_row(g, j, g.l + 1) + ih] = src[split_row(g, j, 1) + ih];
    }
  }
}

}  // namespace

static inline void launched() { pf_count_launch(); }

// 2D: the half-sweep kernel keeps the periodic y-halo rows itself (YIMG 1 / 2 above)
bool pf_sor_stores_y_images(const Geo &g) { return g.dim == 2 && g.n >= 2; }

// pdl = 1: the previous launch on `st` is a half-sweep of the same solve -- launch with programmatic stream
// serialization, so this kernel's blocks are scheduled, and load their coefficients, while that one drains (the kernel
// awaits the dependency before it touches the other colour's pressure).  Captured into a CUDA graph like any launch.
template <class K>
static void launch_sweep(K kernel, dim3 grid, dim3 block, cudaStream_t st, int pdl, Geo g, SplitSet own, const double *po,
                         int colour, double relux, unsigned long long *err_bits, int k0, int kstride) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  PF_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, g, own, po, colour, relux, err_bits, k0, kstride));
}

void k_sor_sweep(const Geo &g, const Phys &ph, const SplitSet S[2], int colour, int with_error,
                 unsigned long long *err_bits, int variant, cudaStream_t st, int k0, int kstride, int nplanes, int pdl) {
  const double relux = ph.relux;
  if (nplanes < 0) nplanes = g.lz;
  if (nplanes == 0) return;
  const int pairs = ((g.m + 1) / 2 + 1) / 2;  // ceil(ceil(m/2)/2)
  // rows shorter than 64 pairs (the 64^3 room deck: 16) get a narrower, taller block: no idle lanes
  int bx = SBX;
  while (bx > 8 && bx / 2 >= pairs) bx /= 2;
  const int by = SBX * SBY / bx;
  const dim3 block(bx, by, 1);
  const dim3 grid((pairs + bx - 1) / bx, (g.n + by - 1) / by, g.dim == 3 ? nplanes : 1);
  const SplitSet &own = S[colour];
  const double *po = S[colour ^ 1].p;
  auto go = [&](auto kernel) { launch_sweep(kernel, grid, block, st, pdl, g, own, po, colour, relux, err_bits, k0, kstride); };
  if (variant == 2 && g.dim == 3 && ph.scase == PF_IBM3_UNIFORM) {
    const double *eo = S[colour ^ 1].eps;
    if (with_error) sor_sweep_eps_kernel<true><<<grid, block, 0, st>>>(g, ph, own, po, eo, colour, err_bits, k0, kstride);
    else            sor_sweep_eps_kernel<false><<<grid, block, 0, st>>>(g, ph, own, po, eo, colour, err_bits, k0, kstride);
  } else if (g.dim == 3) {
    if (pdl) { if (with_error) go(sor_sweep_kernel<3, true, 0, true>); else go(sor_sweep_kernel<3, false, 0, true>); }
    else     { if (with_error) go(sor_sweep_kernel<3, true>);          else go(sor_sweep_kernel<3, false>); }
  } else if (pf_sor_stores_y_images(g) && g.n % 2 == 0) {
    if (pdl) { if (with_error) go(sor_sweep_kernel<2, true, 1, true>); else go(sor_sweep_kernel<2, false, 1, true>); }
    else     { if (with_error) go(sor_sweep_kernel<2, true, 1>);       else go(sor_sweep_kernel<2, false, 1>); }
  } else if (pf_sor_stores_y_images(g)) {
    if (pdl) { if (with_error) go(sor_sweep_kernel<2, true, 2, true>); else go(sor_sweep_kernel<2, false, 2, true>); }
    else     { if (with_error) go(sor_sweep_kernel<2, true, 2>);       else go(sor_sweep_kernel<2, false, 2>); }
  } else {
    if (with_error) go(sor_sweep_kernel<2, true>);
    else            go(sor_sweep_kernel<2, false>);
  }
  launched();
}

void k_sor_halo_y(const Geo &g, double *p0, double *p1, int colour_mask, cudaStream_t st) {
  const int cnt = (g.m + 1) / 2;
  sor_halo_y_kernel<<<dim3((cnt + 127) / 128, g.dim == 3 ? g.lz : 1), 128, 0, st>>>(g, p0, p1, colour_mask);
  launched();
}

void k_sor_halo_z_local(const Geo &g, double *p0, double *p1, int colour_mask, cudaStream_t st) {
  const int cnt = (g.m + 1) / 2;
  sor_halo_z_kernel<<<dim3((cnt + 127) / 128, g.n), 128, 0, st>>>(g, p0, p1, colour_mask);
  launched();
}
