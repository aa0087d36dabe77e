// pf_sor_tma.cu -- SOR variant 6: the fused red+black pass of pf_sor_fused.cu with its operands staged
// by TMA (cp.async.bulk.tensor, 3-D tensor maps over the depth-2-ghost checkerboard arrays) into a
// multi-stage shared-memory pipeline guarded by mbarriers.
//
// Why: the register-prefetch version (variant 3) moves the right bytes (48 B/cell/sweep) but is bound by
// latency and instruction issue -- every z-step is a short dependent chain of scalar 8-byte loads.  Here
// lane 0 of a dedicated producer warp issues ten box copies per plane three steps ahead of their use; the
// 512 compute threads (one tile element each, both colours) only touch shared memory.
//
// Tile: 32 columns x 16 rows of checkerboard elements (ring included: owned columns 1..30, owned rows
// 1..14 -> 82 % of the tile; 64 x 8 measured slower: 73 %), streamed along a z-chunk.  Per plane p the pipeline holds
//   group(p) : P0 BB0 CX0 CX1 CY0 CY1 CZ0 CZ1 BB1          (NG = 5 slots, issued NG-2 planes ahead; used by red(p), black(p))
//   P1(p)    : the old black pressure, box widened by the ring   (5 slots; used by red(p-1), red(p), red(p+1),
//              black(p))
// and a 3-slot ring R of the new red values (231 KB of shared memory in total, one block per SM).  Arithmetic is
// the same sor_update() as everywhere else.  Plane images (periodic wrap on one rank, the neighbour rank's ghost
// planes on a z-slab) are stored by the thread that owns the cell: store_with_images().
#include <stdlib.h>

#include <algorithm>

#include "pf_tma_common.cuh"

namespace {

using namespace pf_tma;

#ifndef PF_TMA_TW
#define PF_TMA_TW 32
#define PF_TMA_TR 16
#endif
constexpr int TW = PF_TMA_TW;   // tile columns (elements), ring columns 0 and TW-1
constexpr int TWP = TW + 4;     // widened boxes: columns -2 .. TW+1
constexpr int TR = PF_TMA_TR;   // tile rows, ring rows 0 and TR-1   (TW*TR == 512 compute threads)
#ifndef PF_TMA_NG
#define PF_TMA_NG 5
#define PF_TMA_NR 3
#endif
constexpr int NG = PF_TMA_NG;   // group slots: planes k-1, k in use, the rest landed / in flight (lead = NG-2 steps)
constexpr int NP = 5;         // P1 slots
constexpr int NR = PF_TMA_NR;   // R slots (3 suffice: there are two block barriers per z-step)
#ifndef PF_TMA_MINB
#define PF_TMA_MINB 1           // resident blocks per SM the kernel is compiled for
#endif
constexpr int GLEAD = NG - 2;   // group(k+GLEAD) is issued at step k
constexpr int NCOMPUTE = TW * TR;   // compute threads: one per tile element (512)
static_assert(TW * TR == NCOMPUTE, "one compute thread per tile element");
constexpr int NTHREADS = NCOMPUTE + 32;   // + one producer warp

// byte sizes of the staged boxes (all multiples of 128)
constexpr int SZ_N = TW * TR * 8;             // narrow box            4096
constexpr int SZ_W = TWP * TR * 8;            // wide box              4608
constexpr int SZ_CY1 = TW * (TR + 1) * 8;     // rows -1 .. TR-1       4352
constexpr int SZ_P1 = (TWP * (TR + 2) * 8 + 127) / 128 * 128;   // padded to a multiple of 128
constexpr int P1_BYTES = TWP * (TR + 2) * 8;  // bytes actually copied 5184
// group layout: P0, BB0, CY0, CZ0, CZ1, BB1 (narrow) ; CX0, CX1 (wide) ; CY1
constexpr int OFF_P0 = 0, OFF_BB0 = SZ_N, OFF_CY0 = 2 * SZ_N, OFF_CZ0 = 3 * SZ_N, OFF_CZ1 = 4 * SZ_N,
              OFF_BB1 = 5 * SZ_N, OFF_CX0 = 6 * SZ_N, OFF_CX1 = 6 * SZ_N + SZ_W, OFF_CY1 = 6 * SZ_N + 2 * SZ_W;
constexpr int SZ_GROUP = 6 * SZ_N + 2 * SZ_W + SZ_CY1;   // 38144
constexpr int GROUP_BYTES = SZ_GROUP;
constexpr int SZ_R = TWP * TR * 8;                        // 4608 per slot
constexpr int SMEM_BYTES = NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R + 256 + 128;   // + mbarriers, block-max scratch, alignment slack

struct TmaMaps {
  CUtensorMap p0, p1, cx0, cx1, cy0, cy1, cz0, cz1, bb0, bb1;
};

struct TmaArgs {
  int NY2, hplane2, cz_planes;
  double *pout0, *pout1;
  double *ilo0, *ilo1, *ihi0, *ihi1;   // image destinations of planes 1,2 / lz-1,lz (FusedArrays::img_lo / img_hi)
  int dk_lo, dk_hi;
};

// 512 compute threads (one checkerboard element of the 32x16 tile each, 16 warps to hide the fp64
// dependency chains) + one producer warp whose lane 0 issues the TMA copies NG-2 planes ahead.
// Barriers: id 1 = compute threads only (red values visible before the black stage),
//           id 2 = everybody (step finished: the slots of planes k-2 may be overwritten).
__global__ void __launch_bounds__(NTHREADS, PF_TMA_MINB) sor_tma_kernel(const __grid_constant__ TmaMaps M, Geo g, TmaArgs A,
                                                              double relux, unsigned long long *err_bits) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  unsigned char *grp = smem;                                  // NG x SZ_GROUP
  unsigned char *p1s = smem + NG * SZ_GROUP;                  // NP x SZ_P1
  double *Rs = reinterpret_cast<double *>(smem + NG * SZ_GROUP + NP * SZ_P1);   // NR x TR x TWP
  uint64_t *gbar = reinterpret_cast<uint64_t *>(smem + NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R);   // NG
  uint64_t *pbar = gbar + NG;                                                                     // NP
  double *wmax = reinterpret_cast<double *>(pbar + NP + 1);

  const int tid = threadIdx.x;
  const int h0 = (int)blockIdx.x * (TW - 2) - 2;     // element index of tile column 0 (stride TW-2, even)
  const int j0 = (int)blockIdx.y * (TR - 2);         // ext row 0 of the tile; owned rows j0+1 .. j0+6
  const int kc0 = (int)blockIdx.z * A.cz_planes + 1;
  const int kc1 = min(kc0 + A.cz_planes - 1, g.lz);
  const int kfirst = kc0 - 1, klast = kc1 + 1;       // red planes
  // array coordinates of the boxes
  const int xn = g.H0 + h0, xw = xn - 2;             // narrow / wide box column origin
  const int yn = j0 + 1, ym = j0;                    // box row origin: rows j0.. / rows j0-1..
  if (tid == 0) {
    for (int q = 0; q < NG; ++q) mbar_init(&gbar[q], 1);
    for (int q = 0; q < NP; ++q) mbar_init(&pbar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // roles: warps 0..15 compute, warp 16 produces (its lane 0 issues the TMA copies).  Both roles run the
  // SAME z-loop and meet at ONE bar.sync per step (a single call site for the whole block).
  const bool is_producer = tid >= NCOMPUTE;
  const bool lead = tid == NCOMPUTE;
  {
    auto issue_group = [&](int p) {
      const int q = p - kfirst;
      unsigned char *b = grp + (q % NG) * SZ_GROUP;
      uint64_t *bar = &gbar[q % NG];
      const int z = p + 1;
      mbar_expect_tx(bar, GROUP_BYTES);
      tma_load_3d(b + OFF_P0, &M.p0, bar, xn, yn, z);
      tma_load_3d(b + OFF_BB0, &M.bb0, bar, xn, yn, z);
      tma_load_3d(b + OFF_CY0, &M.cy0, bar, xn, yn, z);
      tma_load_3d(b + OFF_CZ0, &M.cz0, bar, xn, yn, z);
      tma_load_3d(b + OFF_CZ1, &M.cz1, bar, xn, yn, z - 1);   // cz1 of plane p-1: ab of red(p) == at of black(p-1)
      tma_load_3d(b + OFF_BB1, &M.bb1, bar, xn, yn, z);
      tma_load_3d(b + OFF_CX0, &M.cx0, bar, xw, yn, z);
      tma_load_3d(b + OFF_CX1, &M.cx1, bar, xw, yn, z);
      tma_load_3d(b + OFF_CY1, &M.cy1, bar, xn, ym, z);
    };
    auto issue_p1 = [&](int p) {
      const int q = p - (kfirst - 1);
      uint64_t *bar = &pbar[q % NP];
      mbar_expect_tx(bar, P1_BYTES);
      tma_load_3d(p1s + (q % NP) * SZ_P1, &M.p1, bar, xw, ym, p + 1);
    };
    if (lead) {   // prologue: P1 planes kfirst-1 .. kfirst+2 ; group planes kfirst, kfirst+1
      issue_p1(kfirst - 1);
      issue_p1(kfirst);
      issue_group(kfirst);
      issue_p1(kfirst + 1);
      if (kfirst + 1 <= klast) issue_group(kfirst + 1);
      issue_p1(kfirst + 2);
      for (int d = 2; d < GLEAD; ++d)
        if (kfirst + d <= klast) issue_group(kfirst + d);
    }
  }
  auto produce = [&](int k) {   // lane 0 of the producer warp, once per step: planes k+2 (group) and k+3 (P1)
    if (k + GLEAD <= klast) {
      const int q = k + GLEAD - kfirst;
      unsigned char *b = grp + (q % NG) * SZ_GROUP;
      uint64_t *bar = &gbar[q % NG];
      const int z = k + GLEAD + 1;
      mbar_expect_tx(bar, GROUP_BYTES);
      tma_load_3d(b + OFF_P0, &M.p0, bar, xn, yn, z);
      tma_load_3d(b + OFF_BB0, &M.bb0, bar, xn, yn, z);
      tma_load_3d(b + OFF_CY0, &M.cy0, bar, xn, yn, z);
      tma_load_3d(b + OFF_CZ0, &M.cz0, bar, xn, yn, z);
      tma_load_3d(b + OFF_CZ1, &M.cz1, bar, xn, yn, z - 1);
      tma_load_3d(b + OFF_BB1, &M.bb1, bar, xn, yn, z);
      tma_load_3d(b + OFF_CX0, &M.cx0, bar, xw, yn, z);
      tma_load_3d(b + OFF_CX1, &M.cx1, bar, xw, yn, z);
      tma_load_3d(b + OFF_CY1, &M.cy1, bar, xn, ym, z);
    }
    if (k + 3 <= klast + 1) {
      const int q = k + 3 - (kfirst - 1);
      uint64_t *bar = &pbar[q % NP];
      mbar_expect_tx(bar, P1_BYTES);
      tma_load_3d(p1s + (q % NP) * SZ_P1, &M.p1, bar, xw, ym, k + 4);
    }
  };

  // ---------------------------------------------------- compute threads
  // All shared-memory operands are addressed as  slot base (uniform, rotated per step) + thread offset
  // (3 registers) + compile-time displacement, and every predicate that does not depend on k is
  // hoisted: the loop body is loads, the two updates and the stores.
  const int e = tid % TW, ty = tid / TW;             // tile column / row of this thread
  const int ih = h0 + e;
  const int j = j0 + ty;
  const int m = g.m, ihmax = (g.m + 1) >> 1;
  const bool in_dom = j <= g.n + 1 && ih >= -1 && ih <= ihmax;
  const bool own = ty >= 1 && ty <= TR - 2 && j >= 1 && j <= g.n && e >= 1 && e <= TW - 2;
  const double omr = 1. - relux;
  const int dj = (j <= 2) ? g.n * g.HX : ((j >= g.n - 1) ? -g.n * g.HX : 0);
  const int sj = (j + g.koff) & 1;
  const int cbase = g.H0 + ih + g.HX * (j + 1);       // + hplane2*(k+1) = global element index
  // per-parity cell data (s = parity of i in this row at this plane; alternates with k)
  const int iS0 = 2 * ih + 2, iS1 = 2 * ih + 1;
  const bool cellS0 = in_dom && iS0 >= 1 && iS0 <= m, cellS1 = in_dom && iS1 >= 1 && iS1 <= m;
  const uint32_t sb = smem_u32(smem);
  const uint32_t oN = (uint32_t)(ty * TW + e) * 8;                 // narrow box
  const uint32_t oW = (uint32_t)(ty * TWP + e + 2) * 8;            // wide group box / R slot
  const uint32_t oP = (uint32_t)((ty + 1) * TWP + e + 2) * 8;      // P1 box
  const uint32_t gbase = sb, pbase = sb + NG * SZ_GROUP, rbase = pbase + NP * SZ_P1;
  double cza = 0., czb = 0.;                          // cz_red of this element at planes k-2, k-1
  double emax = 0.;
  // rotating slots: group(k), group(k-1); P1(k-1), P1(k), P1(k+1); R(k), R(k-1), R(k-2)
  int gq = 0, pq = 1;                                 // step counter
  int gs = 0, rs = 0;                                 // slot of group(k), slot of R(k)
  uint32_t gphase = 0;                                // bit s = parity of the next completion of group slot s
  uint32_t pphase = 0;                                // bit s = parity of the next completion of P1 slot s
  uint32_t gK = gbase, gKb = gbase;
  uint32_t pA = pbase, pB = pbase + SZ_P1, pC = pbase + 2 * SZ_P1;
  uint32_t rK = rbase, rKb = rbase, rKb1 = rbase;

  if (!is_producer) {
    mbar_wait(&pbar[0], 0);
    mbar_wait(&pbar[1], 0);
  }
  pphase = 3;                                          // slots 0 and 1 have completed phase 0
  int pc = 2;                                          // slot of P1(k+1)
  for (int k = kfirst; k <= klast; ++k) {
    if (is_producer) {
      if (lead) produce(k);
    } else {
    mbar_wait(&gbar[gs], (gphase >> gs) & 1);
    gphase ^= 1u << gs;
    mbar_wait(&pbar[pc], (pphase >> pc) & 1);
    pphase ^= 1u << pc;
    const int s = (sj + k) & 1;       // parity of i: red row at plane k, black row at plane k-1
    const bool cell = s ? cellS1 : cellS0;
    const int i = s ? iS1 : iS0;
    const bool inchunk = k >= kc0 && k <= kc1;
    // ------------------------------------------ red stage, plane k
    const double pold = lds(gK + OFF_P0 + oN);
    const double at = lds(gK + OFF_CZ0 + oN);
    double val = pold;
    if (cell) {
      // west/east neighbours in the black array: s=1 -> {e-1, e}, s=0 -> {e, e+1}
      const uint32_t sh = s ? 8u : 0u;
      val = sor_update(lds(gK + OFF_BB0 + oN), lds(gK + OFF_CX0 + oW), lds(gK + OFF_CX1 + oW - sh),
                       lds(gK + OFF_CY0 + oN), lds(gK + OFF_CY1 + oN) /* row j-1: the box starts at row j0-1 */, at,
                       lds(gK + OFF_CZ1 + oN) /* cz1(k-1) */, lds(pB + oP + 8 - sh), lds(pB + oP - sh),
                       lds(pB + oP + TWP * 8), lds(pB + oP - TWP * 8), lds(pC + oP), lds(pA + oP), pold, relux, omr, i,
                       m);
      if (own && inchunk) {
        const bool lo = k <= 2, hi = k >= g.lz - 1;
        store_with_images(A.pout0, lo ? A.ilo0 : (hi ? A.ihi0 : nullptr), cbase + A.hplane2 * (k + 1), dj,
                          lo ? A.dk_lo : A.dk_hi, val);
      }
    }
    sts(rK + oW, val);
    named_bar(1, NCOMPUTE);
    // ------------------------------------------ black stage, plane k-1
    const int kb = k - 1;
    if (kb >= kc0 && kb <= kc1 && own && cell) {
      const uint32_t sh = s ? 8u : 0u;
      const double bold = lds(pA + oP);                               // black own old value (plane k-1)
      const double v = sor_update(lds(gKb + OFF_BB1 + oN), lds(gKb + OFF_CX1 + oW), lds(gKb + OFF_CX0 + oW - sh),
                                  lds(gKb + OFF_CY1 + oN + TW * 8) /* own row j */, lds(gKb + OFF_CY0 + oN - TW * 8),
                                  lds(gK + OFF_CZ1 + oN) /* cz1(kb) */, cza, lds(rKb + oW + 8 - sh), lds(rKb + oW - sh),
                                  lds(rKb + oW + TWP * 8), lds(rKb + oW - TWP * 8), val, lds(rKb1 + oW), bold, relux,
                                  omr, i, m);
      const bool lo = kb <= 2, hi = kb >= g.lz - 1;
      store_with_images(A.pout1, lo ? A.ilo1 : (hi ? A.ihi1 : nullptr), cbase + A.hplane2 * (kb + 1), dj,
                        lo ? A.dk_lo : A.dk_hi, v);
      emax = fmax(emax, fabs(v - bold));
    }
    cza = czb; czb = at;
    // rotate the slots
    ++gq;
    gs = (gs + 1 == NG) ? 0 : gs + 1;
    gKb = gK; gK = gbase + (uint32_t)gs * SZ_GROUP;
    pA = pB; pB = pC;
    pc = (pc + 1 == NP) ? 0 : pc + 1;
    pC = pbase + (uint32_t)pc * SZ_P1;
    rs = (rs + 1 == NR) ? 0 : rs + 1;
    rKb1 = rKb; rKb = rK; rK = rbase + (uint32_t)rs * SZ_R;
    // order this step's shared-memory reads before the async-proxy writes of the next copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }   // compute role
    named_bar(2, NTHREADS);
  }
  (void)pq;
  if (is_producer) return;
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if ((tid & 31) == 0) wmax[tid >> 5] = emax;
  named_bar(1, NCOMPUTE);
  if (tid < 32) {
    double v = (tid < NCOMPUTE / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

}  // namespace

bool pf_tma_applicable(const Geo &g, const Phys &ph, int nranks) {
  return pf_fused_applicable(g, ph, nranks) && (g.HX * 8) % 16 == 0;
}

// one red-black iteration through the TMA pipeline: reads A.p[in], writes A.p[in^1]
void k_tma_iteration(const Geo &g, const Phys &ph, FusedArrays &A, int in, unsigned long long *err_bits,
                     cudaStream_t st) {
  if (!A.tma_cache) {   // first launch of this solver: opt in to the large shared-memory carve-out on ITS device, and
                        // encode the ten tensor maps of each ping-pong direction once
    PF_CUDA_OK(cudaFuncSetAttribute(sor_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    TmaMaps *c = new TmaMaps[2];
    for (int d = 0; d < 2; ++d) {
      c[d].p0 = make_map(g, A.p[d][0], TW, TR);
      c[d].p1 = make_map(g, A.p[d][1], TWP, TR + 2);
      c[d].cx0 = make_map(g, A.cx[0], TWP, TR);
      c[d].cx1 = make_map(g, A.cx[1], TWP, TR);
      c[d].cy0 = make_map(g, A.cy[0], TW, TR);
      c[d].cy1 = make_map(g, A.cy[1], TW, TR + 1);
      c[d].cz0 = make_map(g, A.cz[0], TW, TR);
      c[d].cz1 = make_map(g, A.cz[1], TW, TR);
      c[d].bb0 = make_map(g, A.bb[0], TW, TR);
      c[d].bb1 = make_map(g, A.bb[1], TW, TR);
    }
    A.tma_cache = c;
  }
  const TmaMaps &M = static_cast<const TmaMaps *>(A.tma_cache)[in];
  TmaArgs a;
  a.NY2 = g.n + 4;
  a.hplane2 = g.HX * (g.n + 4);
  a.cz_planes = A.cz_planes;
  a.pout0 = A.p[in ^ 1][0];
  a.pout1 = A.p[in ^ 1][1];
  a.ilo0 = A.img_lo[in ^ 1][0]; a.ilo1 = A.img_lo[in ^ 1][1];
  a.ihi0 = A.img_hi[in ^ 1][0]; a.ihi1 = A.img_hi[in ^ 1][1];
  a.dk_lo = (int)A.dk_lo;
  a.dk_hi = (int)A.dk_hi;
  const int cols = ((g.m + 1) >> 1) + 2;            // elements -1 .. ihmax
  const int xt = (cols + (TW - 2) - 1) / (TW - 2);
  const int yt = (g.n + (TR - 2) - 1) / (TR - 2);
  const int zt = (g.lz + A.cz_planes - 1) / A.cz_planes;
  sor_tma_kernel<<<dim3(xt, yt, zt), NTHREADS, SMEM_BYTES, st>>>(M, g, a, ph.relux, err_bits);
  pf_count_launch();
}

void pf_tma_release(FusedArrays &A) {
  delete[] static_cast<TmaMaps *>(A.tma_cache);
  A.tma_cache = nullptr;
}

// z-chunk size for the TMA kernel (1 block per SM): pf_chunk_planes() of pf_kernels.cu
int pf_tma_chunk(const Geo &g) {
  if (const char *e = getenv("PF_TMA_CHUNK")) {   // tuning experiments only
    const int v = atoi(e);
    if (v >= 1) return v < g.lz ? v : g.lz;
  }
  const int cols = ((g.m + 1) >> 1) + 2;
  const int xt = (cols + (TW - 2) - 1) / (TW - 2);
  const int yt = (g.n + (TR - 2) - 1) / (TR - 2);
  return pf_chunk_planes(g.lz, (long long)xt * yt, pf_sm_count() * PF_TMA_MINB);
}
