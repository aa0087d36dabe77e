// pf_sor_tma2.cu -- SOR variant 8: the fused red+black pass (pf_sor_fused.cu) with a split operand path.
//
// Variant 6 (pf_sor_tma.cu) stages ALL ten operand boxes of a plane through shared memory: 43 KB per plane, a
// 231 KB pipeline, one 544-thread block per SM, one element per thread, ~175 instructions per cell update; its
// ncu capture (profiles/r01_fused_summary.md) shows 26 % of the warp slots filled and 38 % of the stall samples
// waiting for the next plane's mbarrier.  Here
//   * only the operands a NEIGHBOUR thread reads go through shared memory, by TMA: the other colour's old pressure
//     P1 (widened by the ring) and the face coefficients CX0 CX1 CY0 CY1 -- 22.6 KB per plane;
//   * the operands only the owning thread reads (its own old pressure P0, BB0, BB1, CZ0, CZ1) are loaded straight
//     from global memory into registers one plane ahead (16-byte read-only loads, fully coalesced);
//   * every thread owns TWO adjacent checkerboard elements: all shared-memory traffic is 16-byte LDS/STS (11 + 12
//     loads per pair and iteration instead of 2 x 27), address arithmetic and predicates are shared by the pair,
//     and the two updates are independent dependency chains for the fp64 pipe;
//   * the west/east neighbours of a pair are three consecutive elements of the other colour's row, starting one
//     slot earlier when i is odd: one address offset per step instead of selects;
//   * everything that does not depend on k (predicates, shared-memory offsets) is computed once and pinned in
//     registers; global addresses advance by one plane per step;
//   * the pipeline is 114 KB, so TWO 256-thread blocks share an SM: while one waits at its barrier or for a plane the
//     other one computes.  There is no producer warp (a ninth warp would cost every thread 16 registers of the
//     allocation): thread 0 issues the five box copies of a step right after the block barrier, ~20 instructions;
//   * four R slots instead of three make the second block barrier of a z-step unnecessary (one bar.sync per step).
// Arithmetic, operation order, tile geometry (32x16 elements, ring included), z-chunking, image stores and the
// error reduction are those of variant 6: sor_update() / store_with_images() of pf_tma_common.cuh.
#include <stdlib.h>

#include <algorithm>

#include "pf_tma_common.cuh"

namespace {

using namespace pf_tma;

// Experiment switches (tools/sor_lab.py; never defined in the product build -- results are WRONG with them):
//   PF_TMA2_NOCOMPUTE  all copies, loads, barriers and stores, no SOR arithmetic  -> the memory pipeline alone
//   PF_TMA2_NOLOAD     no TMA copies, no global loads, no mbarrier waits           -> arithmetic + barriers alone
#ifndef PF_TMA2_MINB
#define PF_TMA2_MINB 2          // resident blocks per SM the kernel is compiled for
#endif
constexpr int TW = 32;          // tile columns (elements), ring columns 0 and TW-1
constexpr int TWP = TW + 4;     // widened boxes: columns -2 .. TW+1
constexpr int TR = 16;          // tile rows, ring rows 0 and TR-1
constexpr int NG = 4;           // coefficient-group slots: planes k-1, k in use, k+1 landed / landing, k+2 in flight
constexpr int NP = 5;           // P1 slots: planes k-1, k, k+1 in use, k+2, k+3 in flight
constexpr int NR = 4;           // new-red slots: planes k-2, k-1, k in use, k+1 may already be written
constexpr int NCOMPUTE = (TW / 2) * TR;   // one compute thread per element PAIR (256)
constexpr int NTHREADS = NCOMPUTE;        // thread 0 also issues the TMA copies

constexpr int SZ_CX = TWP * TR * 8;       // wide box   4608
constexpr int SZ_CY = TW * TR * 8;        // narrow box 4096 (CY0: rows j0 .. j0+15, CY1: rows j0-1 .. j0+14)
constexpr int OFF_CX0 = 0, OFF_CX1 = SZ_CX, OFF_CY0 = 2 * SZ_CX, OFF_CY1 = 2 * SZ_CX + SZ_CY;
constexpr int SZ_GROUP = 2 * SZ_CX + 2 * SZ_CY;                 // 17408
constexpr int P1_BYTES = TWP * (TR + 2) * 8;                    // 5184
constexpr int SZ_P1 = (P1_BYTES + 127) / 128 * 128;             // 5248
constexpr int SZ_R = TWP * TR * 8;                              // 4608
constexpr int SMEM_BYTES = NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R + 256 + 128;   // 114,688: two blocks per SM
static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two resident blocks per SM");


struct Maps2 {
  CUtensorMap p1, cx0, cx1, cy0, cy1;
};

struct Args2 {
  int hplane2, cz_planes;
  const double *pin0, *bb0, *bb1, *cz0, *cz1;   // own-element operands, read straight from global memory
  double *pout0, *pout1;
  double *ilo0, *ilo1, *ihi0, *ihi1;            // image destinations of planes 1,2 / lz-1,lz (FusedArrays::img_lo / img_hi)
  int dk_lo, dk_hi;
};

__device__ __forceinline__ double2 lds2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts2(uint32_t addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
// streaming 16-byte load of an operand nobody else on this SM reads.  Unconditional on purpose: a predicated load
// needs a register copy after it, and that copy waits for the data -- one step too early (ncu: 42 % of all stall
// samples on exactly that MOV).  Threads outside the arrays load a valid dummy address instead.
__device__ __forceinline__ double2 ldg2(const double *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// The SOR update of the two cells of a pair (ibm_3d_uniform_omp_cpu.f90:510-515, left to right; ap from the raw
// coefficients :402), written as ONE straight line of code so that the two dependency chains interleave in the fp64
// pipe: no per-cell branch, and the division is quot_fast() (pf_tma_common.cuh) -- the bits of `r / ap` whenever
// its guard holds.  Returns false if a cell that counts (`need` bit 0 / 1) fell outside the guard: the caller then
// redoes the pair with update_pair_slow().  Cells that do not count may produce anything; the caller discards them.
struct Pair {
  double2 bb, ae, aw, an, as, at, ab, pE, pW, pN, pS, pT, pB, pold;
};
__device__ __forceinline__ bool update_pair(const Pair &c, double relux, double omr, int need, double2 &out) {
  const double apx = -c.ae.x - c.aw.x - c.an.x - c.as.x - c.at.x - c.ab.x;
  const double apy = -c.ae.y - c.aw.y - c.an.y - c.as.y - c.at.y - c.ab.y;
  const double rx = c.bb.x - c.ae.x * c.pE.x - c.aw.x * c.pW.x - c.an.x * c.pN.x - c.as.x * c.pS.x - c.at.x * c.pT.x -
                    c.ab.x * c.pB.x;
  const double ry = c.bb.y - c.ae.y * c.pE.y - c.aw.y * c.pW.y - c.an.y * c.pN.y - c.as.y * c.pS.y - c.at.y * c.pT.y -
                    c.ab.y * c.pB.y;
  out.x = quot_fast(rx, apx) * relux + c.pold.x * omr;
  out.y = quot_fast(ry, apy) * relux + c.pold.y * omr;
  return (quot_guard(rx, apx) || !(need & 1)) && (quot_guard(ry, apy) || !(need & 2));
}
// the same with the plain division, and with boundrary_matrix's folds for the cells of the inlet (i == 1) and outlet
// (i == m) columns (:640-641, :651-656): out of the hot path -- x-edge tiles, exact zeros, denormals
__device__ __forceinline__ void update_pair_slow(const Pair &c, double relux, double omr, int need, int ia, int m,
                                              double2 &out) {
  if (need & 1)
    out.x = sor_update(c.bb.x, c.ae.x, c.aw.x, c.an.x, c.as.x, c.at.x, c.ab.x, c.pE.x, c.pW.x, c.pN.x, c.pS.x, c.pT.x,
                       c.pB.x, c.pold.x, relux, omr, ia, m);
  if (need & 2)
    out.y = sor_update(c.bb.y, c.ae.y, c.aw.y, c.an.y, c.as.y, c.at.y, c.ab.y, c.pE.y, c.pW.y, c.pN.y, c.pS.y, c.pT.y,
                       c.pB.y, c.pold.y, relux, omr, ia + 2, m);
}

// the pair's own cells (mask bit 0 / 1 = element a / b): predicated stores, no divergent region
__device__ __forceinline__ void store_pair(double *dst, double2 v, int mask) {
  if (mask == 3) *reinterpret_cast<double2 *>(dst) = v;
  if (mask == 1) dst[0] = v.x;
  if (mask == 2) dst[1] = v.y;
}

struct Own {          // the operands of one z-step that only this thread reads
  double2 p0, bb0, cz0;   // plane k   : old red pressure, red source, at of red(k)
  double2 cz1, bb1;       // plane k-1 : ab of red(k) == at of black(k-1), black source
  double2 czm;            // plane k-2 : cz0 again = ab of black(k-1).  Loaded a second time (an L2 hit) rather than
                          // carried in registers from step k-2: ptxas coalesces such a carried copy with the freshly
                          // loaded value and places the MOV right behind the load -- a full DRAM latency per step
};

__global__ void __launch_bounds__(NTHREADS, PF_TMA2_MINB) sor_tma2_kernel(const __grid_constant__ Maps2 M, Geo g, Args2 A,
                                                                double relux, unsigned long long *err_bits) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  unsigned char *grp = smem;                                  // NG x SZ_GROUP
  unsigned char *p1s = smem + NG * SZ_GROUP;                  // NP x SZ_P1
  uint64_t *gbar = reinterpret_cast<uint64_t *>(smem + NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R);   // NG
  uint64_t *pbar = gbar + NG;                                                                     // NP
  double *wmax = reinterpret_cast<double *>(pbar + NP + 1);

  const int tid = threadIdx.x;
  const int h0 = (int)blockIdx.x * (TW - 2) - 2;     // element index of tile column 0 (stride TW-2, even)
  const int j0 = (int)blockIdx.y * (TR - 2);         // ext row 0 of the tile; owned rows j0+1 .. j0+TR-2
  const int kc0 = (int)blockIdx.z * A.cz_planes + 1;
  const int kc1 = min(kc0 + A.cz_planes - 1, g.lz);
  const int kfirst = kc0 - 1, klast = kc1 + 1;       // red planes
  const int xn = g.H0 + h0, xw = xn - 2;             // narrow / wide box column origin (array coordinates)
  const int yn = j0 + 1, ym = j0;                    // box row origin: rows j0.. / rows j0-1..
  if (tid == 0) {
    for (int q = 0; q < NG; ++q) mbar_init(&gbar[q], 1);
    for (int q = 0; q < NP; ++q) mbar_init(&pbar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // ---------------------------------------------------- the copies (issued by thread 0)
  auto issue_group = [&](int p) {        // coefficient boxes of plane p
    const int q = (p - kfirst) % NG;
    unsigned char *b = grp + q * SZ_GROUP;
    uint64_t *bar = &gbar[q];
    mbar_expect_tx(bar, SZ_GROUP);
    tma_load_3d(b + OFF_CX0, &M.cx0, bar, xw, yn, p + 1);
    tma_load_3d(b + OFF_CX1, &M.cx1, bar, xw, yn, p + 1);
    tma_load_3d(b + OFF_CY0, &M.cy0, bar, xn, yn, p + 1);
    tma_load_3d(b + OFF_CY1, &M.cy1, bar, xn, ym, p + 1);
  };
  auto issue_p1 = [&](int p) {           // old black pressure of plane p, widened by the ring
    const int q = (p - (kfirst - 1)) % NP;
    mbar_expect_tx(&pbar[q], P1_BYTES);
    tma_load_3d(p1s + q * SZ_P1, &M.p1, &pbar[q], xw, ym, p + 1);
  };
#ifdef PF_TMA2_NOLOAD
  const bool issuer = false;
#else
  const bool issuer = tid == 0;
#endif
  if (issuer) {
    // prologue: every slot is free; P1 planes kfirst-1 .. kfirst+NP-2 and groups kfirst .. kfirst+NG-1, roughly in
    // the order of their use
    for (int q = 0; q < NP; ++q) {
      if (kfirst - 1 + q <= klast + 1) issue_p1(kfirst - 1 + q);
      if (q >= 2 && q - 2 < NG && kfirst + q - 2 <= klast) issue_group(kfirst + q - 2);
    }
    for (int q = NP - 2; q < NG; ++q)
      if (kfirst + q <= klast) issue_group(kfirst + q);
  }

  // ---------------------------------------------------- compute
  // warp w holds tile rows w and w+8; lane%16 = pair index along x
  const int lane = tid & 31, wrp = tid >> 5;
  const int e = 2 * (lane & 15), ty = wrp + 8 * (lane >> 4);
  const int ih = h0 + e;                               // element index of the pair's first element (even)
  const int j = j0 + ty;
  const int m = g.m;
  const double omr = 1. - relux;
  // Everything that does not depend on k is computed ONCE, packed and pinned in registers (the empty asm statements
  // keep ptxas from rematerialising a dozen integer instructions per use inside the z-loop).  Flags of this row for
  // the two parities S of i:  bit 0/1 = element a/b is a cell (i in 1..m), bit 2/3 = a/b is stored by this thread
  // (owned and a cell), bit 4 = an inlet/outlet fold (i == 1 or i == m) applies to a or b, bit 5 = S,
  // bit 6/7 = the row has a periodic image n rows up / down (rows 1,2 / n-1,n)
  int fA, fB;
  {
    const int ihmax = (m + 1) >> 1;
    const bool row_dom = j <= g.n + 1;
    const bool dom_a = row_dom && ih >= -1 && ih <= ihmax, dom_b = row_dom && ih + 1 >= -1 && ih + 1 <= ihmax;
    const bool row_own = ty >= 1 && ty <= TR - 2 && j >= 1 && j <= g.n;
    const bool own_a = row_own && e >= 1, own_b = row_own && e + 1 <= TW - 2;
    auto flags = [&](int S) {
      const int ia = 2 * ih + 2 - S, ib = ia + 2;
      const bool ca = dom_a && ia >= 1 && ia <= m, cb = dom_b && ib >= 1 && ib <= m;
      const bool fold = (ca && (ia == 1 || ia == m)) || (cb && (ib == 1 || ib == m));
      return (ca ? 1 : 0) | (cb ? 2 : 0) | (ca && own_a ? 4 : 0) | (cb && own_b ? 8 : 0) | (fold ? 16 : 0) |
             (S ? 32 : 0) | (j <= 2 ? 64 : 0) | (j >= g.n - 1 ? 128 : 0);
    };
    const int SA = (j + g.koff + kfirst) & 1;          // parity of i in this row at the first plane
    fA = flags(SA);
    fB = flags(SA ^ 1);
  }
  const bool ld_ok = j + 1 <= g.n + 3 && g.H0 + ih + 1 <= g.HX - 1;   // the pair lies inside the arrays
  const uint32_t sb = smem_u32(smem);
  uint32_t oN = (uint32_t)(ty * TW + e) * 8;                 // narrow box
  uint32_t oW = (uint32_t)(ty * TWP + e + 2) * 8;            // wide group box / R slot; the P1 box: + one row
  asm volatile("" : "+r"(fA), "+r"(fB), "+r"(oN), "+r"(oW));
  constexpr uint32_t ROW = TWP * 8;
  const uint32_t gbase = sb, pbase = sb + NG * SZ_GROUP, rbase = pbase + NP * SZ_P1;

  // element offset of the pair at the plane of the current step; one plane per step.  Pairs outside the arrays
  // (tile overhang) walk along element 0 of each plane instead: valid memory, values never used, never stored.
  const long long pstride = A.hplane2;
  long long os = (ld_ok ? (long long)g.H0 + ih + (long long)g.HX * (j + 1) : 0) + pstride * (kfirst + 1);

  // operands of the step AFTER the current one (plane os + 1; after the last step that is plane klast + 1 <= lz + 2,
  // which exists in the depth-2-ghost arrays: loaded, never used -- cheaper than branching around the loads)
  auto load_next = [&](Own &o, bool first) {
    const long long ol = os + pstride;
    o.p0 = ldg2(A.pin0 + ol);
    o.bb0 = ldg2(A.bb0 + ol);
    o.cz0 = ldg2(A.cz0 + ol);
    o.cz1 = ldg2(A.cz1 + os);
    o.bb1 = ldg2(A.bb1 + os);
    o.czm = ldg2(A.cz0 + (first ? os : os - pstride));   // (the first two steps have no black stage: any valid address)
  };

  double emax = 0.;
  int gs = 0, gphase = 0;                 // slot of group(k) and the phase bit of its mbarrier
  int pc = 2, pphase = 0;                 // slot of P1(k+1)
  int rs = 0;                             // slot of R(k)
  uint32_t gK = gbase, gKb = gbase;
  uint32_t pA = pbase, pB = pbase + SZ_P1, pC = pbase + 2 * SZ_P1;
  uint32_t rK = rbase, rKb = rbase, rKb1 = rbase;

  Own oa, ob;
  os -= pstride;
  load_next(oa, true);      // the first step's own operands
  os += pstride;
#ifndef PF_TMA2_NOLOAD
  mbar_wait(&pbar[0], 0);
  mbar_wait(&pbar[1], 0);
#endif

  // images of a freshly stored pair: the periodic row image (rows 1,2 <-> n+1,n+2; rows n-1,n <-> -1,0) and the plane
  // image (planes 1,2 / lz-1,lz: the periodic wrap on one rank, the neighbour rank's ghost planes over NVLink on a
  // z-slab; null = exchanged after the launch).  Rare: four rows of n, four planes of lz.
  auto store_images = [&](double *dst, double *imgbase, long long o, int dk, double2 v, int mask, int f) {
    const int dj = (f & 64) ? g.n * g.HX : ((f & 128) ? -g.n * g.HX : 0);
    if (dj) store_pair(dst + dj, v, mask);
    if (imgbase) {
      store_pair(imgbase + o + dk, v, mask);
      if (dj) store_pair(imgbase + o + dk + dj, v, mask);
    }
  };

  // one z-step: red stage of plane k, block barrier, black stage of plane k-1.  f = the row's flags at this step's
  // parity S of i (red at plane k, black at plane k-1 have the same S).
  auto step = [&](int k, const Own &o, Own &next, int f) {
    asm volatile("" : "+r"(f));   // keep the bit tests on f inside the loop (hoisted, they get spilled: 20 live values)
    // west/east neighbours of the pair in the other colour's row = three consecutive elements starting at e-1
    // (S = 1) or e (S = 0):  a -> {x0, x1},  b -> {x1, x2}
    const uint32_t xo = (f & 32) ? (uint32_t)-8 : 0u;
    const int ia = 2 * ih + 2 - ((f >> 5) & 1);
    const bool special = (f & (64 | 128)) != 0;
#ifndef PF_TMA2_NOLOAD
    mbar_wait(&gbar[gs], gphase);
    mbar_wait(&pbar[pc], pphase);
#endif
    // ------------------------------------------ red stage, plane k
    double2 val;
    {
      const uint32_t px = pB + ROW + oW + xo, cx = gK + OFF_CX1 + oW + xo;
      Pair c;
      c.pW.x = lds(px); c.pE.x = lds(px + 8); c.pW.y = c.pE.x; c.pE.y = lds(px + 16);
      c.aw.x = lds(cx); c.aw.y = lds(cx + 8);          // east faces of the west neighbours = the pair's west faces
      c.ae = lds2(gK + OFF_CX0 + oW);                  // own east faces
      c.an = lds2(gK + OFF_CY0 + oN);
      c.as = lds2(gK + OFF_CY1 + oN);                  // row j-1: the box starts at row j0-1
      c.pN = lds2(pB + 2 * ROW + oW); c.pS = lds2(pB + oW);
      c.pT = lds2(pC + ROW + oW); c.pB = lds2(pA + ROW + oW);
      c.bb = o.bb0; c.at = o.cz0; c.ab = o.cz1; c.pold = o.p0;
#ifndef PF_TMA2_NOCOMPUTE
      double2 nv;
      const bool ok = update_pair(c, relux, omr, f & 3, nv);
      if (!ok || (f & 16)) update_pair_slow(c, relux, omr, f & 3, ia, m, nv);
      val.x = (f & 1) ? nv.x : o.p0.x;                 // halo and out-of-domain slots pass their old value on
      val.y = (f & 2) ? nv.y : o.p0.y;
#else
      val.x = o.p0.x + c.pW.x + c.pE.x + c.pE.y + c.aw.x + c.aw.y + c.ae.x + c.an.x + c.as.x + c.pN.x + c.pS.x + c.pT.x + c.pB.x;
      val.y = o.p0.y;
#endif
      if (k >= kc0 && k <= kc1) {
        const int st = (f >> 2) & 3;
        store_pair(A.pout0 + os, val, st);
        const bool lo = k <= 2, hi = k >= g.lz - 1;
        if (special || lo || hi)
          store_images(A.pout0 + os, lo ? A.ilo0 : (hi ? A.ihi0 : nullptr), os, lo ? A.dk_lo : A.dk_hi, val, st, f);
      }
    }
    sts2(rK + oW, val);
    __syncthreads();
    // every warp has finished step k-1: the slots of plane k-2 are free -- refill them.  (No proxy fence: the slots
    // were only READ by this block; the barrier orders those reads before the copies, as a consumer-release
    // mbarrier does in any TMA pipeline.)
    if (issuer) {
      if (k - 2 >= kfirst && k - 2 + NG <= klast) issue_group(k - 2 + NG);
      if (k - 2 >= kfirst - 1 && k - 2 + NP <= klast + 1) issue_p1(k - 2 + NP);
    }
    // The next step's own operands, issued HERE: ptxas tracks every global load of this kernel on one scoreboard, so
    // a first use of this step's operands that came after the issue of the next ones would wait for those too (ncu:
    // 11 % of all stall samples on exactly that instruction).  Behind the barrier every first use of this step's
    // operands -- the red stage -- is past, and the loads have the black stage and the next step's waits to land.
    load_next(next, false);
    // ------------------------------------------ black stage, plane k-1
    const int st = (f >> 2) & 3;
    if (k - 1 >= kc0 && k - 1 <= kc1 && st) {
      const uint32_t rx = rKb + oW + xo, cx = gKb + OFF_CX0 + oW + xo;
      Pair c;
      c.pold = lds2(pA + ROW + oW);                    // black own old value (plane k-1)
      c.pW.x = lds(rx); c.pE.x = lds(rx + 8); c.pW.y = c.pE.x; c.pE.y = lds(rx + 16);
      c.aw.x = lds(cx); c.aw.y = lds(cx + 8);
      c.ae = lds2(gKb + OFF_CX1 + oW);
      c.an = lds2(gKb + OFF_CY1 + oN + TW * 8);        // own row j
      c.as = lds2(gKb + OFF_CY0 + oN - TW * 8);        // row j-1
      c.pN = lds2(rKb + oW + ROW); c.pS = lds2(rKb + oW - ROW);
      c.pB = lds2(rKb1 + oW);
      c.pT = val;
      c.bb = o.bb1; c.at = o.cz1; c.ab = o.czm;
      double2 v = c.pold;
#ifndef PF_TMA2_NOCOMPUTE
      double2 nv;
      const bool ok = update_pair(c, relux, omr, st, nv);
      if (!ok || (f & 16)) update_pair_slow(c, relux, omr, st, ia, m, nv);
      if (st & 1) v.x = nv.x;
      if (st & 2) v.y = nv.y;
#else
      v.x += c.pW.x + c.pE.x + c.pE.y + c.aw.x + c.aw.y + c.ae.x + c.an.x + c.as.x + c.pN.x + c.pS.x + c.pB.x;
#endif
      // unowned elements keep v == pold: they add 0 to the error
      emax = fmax(emax, fmax(fabs(v.x - c.pold.x), fabs(v.y - c.pold.y)));
      store_pair(A.pout1 + os - pstride, v, st);
      const bool lo = k - 1 <= 2, hi = k - 1 >= g.lz - 1;
      if (special || lo || hi)
        store_images(A.pout1 + os - pstride, lo ? A.ilo1 : (hi ? A.ihi1 : nullptr), os - pstride,
                     lo ? A.dk_lo : A.dk_hi, v, st, f);
    }
    // next plane: rotate the slots, advance the destinations
    os += pstride;
    if (++gs == NG) { gs = 0; gphase ^= 1; }
    gKb = gK; gK = gbase + (uint32_t)gs * SZ_GROUP;
    pA = pB; pB = pC;
    if (++pc == NP) { pc = 0; pphase ^= 1; }
    pC = pbase + (uint32_t)pc * SZ_P1;
    if (++rs == NR) rs = 0;
    rKb1 = rKb; rKb = rK; rK = rbase + (uint32_t)rs * SZ_R;
  };

  // Two steps per trip.  The operand sets alternate between two register files (no copies), and since the parity of
  // i alternates with k, each of the two step bodies always sees the same parity: its flags are loop constants.
  for (int k = kfirst; k <= klast; k += 2) {
    step(k, oa, ob, fA);
    if (k + 1 > klast) break;
    step(k + 1, ob, oa, fB);
  }

  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if (lane == 0) wmax[wrp] = emax;
  __syncthreads();
  if (tid < 32) {
    double v = (tid < NCOMPUTE / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

int blocks_per_sm() {
  static int cached = 0;
  if (!cached) {
    PF_CUDA_OK(cudaFuncSetAttribute(sor_tma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int n = 0;
    PF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sor_tma2_kernel, NTHREADS, SMEM_BYTES));
    cached = n > 0 ? n : 1;
  }
  return cached;
}

}  // namespace

bool pf_tma2_applicable(const Geo &g, const Phys &ph, int nranks) { return pf_tma_applicable(g, ph, nranks); }

// z-chunk size: a block takes (cz + 2) z-steps (two redundant red planes); blocks are list-scheduled on
// SMs x resident-blocks slots -- the cost model of pf_tma_chunk()
int pf_tma2_chunk(const Geo &g) {
  if (const char *e = getenv("PF_TMA_CHUNK")) {   // tuning experiments only
    const int v = atoi(e);
    if (v >= 1) return v < g.lz ? v : g.lz;
  }
  const int cols = ((g.m + 1) >> 1) + 2;
  const int xt = (cols + (TW - 2) - 1) / (TW - 2);
  const int yt = (g.n + (TR - 2) - 1) / (TR - 2);
  return pf_chunk_planes(g.lz, (long long)xt * yt, pf_sm_count() * blocks_per_sm());
}

// one red-black iteration: reads A.p[in], writes A.p[in^1]
void k_tma2_iteration(const Geo &g, const Phys &ph, FusedArrays &A, int in, unsigned long long *err_bits,
                      cudaStream_t st) {
  if (!A.tma2_cache) {   // first launch of this solver: the shared-memory opt-in on ITS device and the tensor maps
    PF_CUDA_OK(cudaFuncSetAttribute(sor_tma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    Maps2 *c = new Maps2[2];
    for (int d = 0; d < 2; ++d) {
      c[d].p1 = make_map(g, A.p[d][1], TWP, TR + 2);
      c[d].cx0 = make_map(g, A.cx[0], TWP, TR);
      c[d].cx1 = make_map(g, A.cx[1], TWP, TR);
      c[d].cy0 = make_map(g, A.cy[0], TW, TR);
      c[d].cy1 = make_map(g, A.cy[1], TW, TR);
    }
    A.tma2_cache = c;
  }
  const Maps2 &M = static_cast<const Maps2 *>(A.tma2_cache)[in];
  Args2 a;
  a.hplane2 = g.HX * (g.n + 4);
  a.cz_planes = A.cz_planes;
  a.pin0 = A.p[in][0];
  a.bb0 = A.bb[0]; a.bb1 = A.bb[1];
  a.cz0 = A.cz[0]; a.cz1 = A.cz[1];
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
  sor_tma2_kernel<<<dim3(xt, yt, zt), NTHREADS, SMEM_BYTES, st>>>(M, g, a, ph.relux, err_bits);
  pf_count_launch();
}

// ---- self-check of the branch-free division (pf_tma_common.cuh): random operand pairs with exponents in
// [-exp_range, exp_range] (a few exact zeros, denormals, infinities and NaNs mixed in).  *mismatches = pairs INSIDE the
// guard whose quot_fast() differs from the IEEE quotient (must be 0); *outside = pairs the guard sends to the plain `/`.
namespace {
__global__ void quot_check_kernel(long long n, unsigned long long seed, int exp_range, unsigned long long *out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0, outside = 0;
  auto rnd = [&](unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    return x;
  };
  auto make = [&](unsigned long long x) {
    const unsigned long long mant = x & 0xFFFFFFFFFFFFFull, sign = (x >> 63) << 63;
    const int e = (int)((x >> 52) % (unsigned)(2 * exp_range + 1)) - exp_range;
    double a = __longlong_as_double((long long)(sign | ((unsigned long long)(1023 + e) << 52) | mant));
    const unsigned sel = (unsigned)(x >> 40) & 0xFFFu;
    if (sel == 0) a = 0.0;
    if (sel == 1) a = __longlong_as_double((long long)(sign | (mant >> 7)));          // denormal
    if (sel == 2) a = __longlong_as_double((long long)(sign | 0x7FF0000000000000ull)); // infinity
    if (sel == 3) a = __longlong_as_double((long long)0x7FF8000000000001ull);          // NaN
    return a;
  };
  for (; t < n; t += stride) {
    const unsigned long long x = rnd(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(t + 1));
    const double r = make(x), d = make(rnd(x ^ 0xD6E8FEB86659FD93ull));
    if (!quot_guard(r, d)) { ++outside; continue; }
    if (__double_as_longlong(quot_fast(r, d)) != __double_as_longlong(r / d)) ++bad;
  }
  if (bad) atomicAdd(out, bad);
  if (outside) atomicAdd(out + 1, outside);
}
}  // namespace

extern "C" int pf_debug_quot_mismatches(long long n, unsigned long long seed, int exp_range, long long *mismatches,
                                        long long *outside) {
  if (!mismatches || exp_range < 0 || exp_range > 1000) return 1;
  unsigned long long *dev = nullptr, h[2] = {0, 0};
  if (cudaMalloc(&dev, sizeof(h)) != cudaSuccess) return 1;
  cudaMemset(dev, 0, sizeof(h));
  quot_check_kernel<<<pf_sm_count() * 8, 256>>>(n, seed, exp_range, dev);
  const cudaError_t e = cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  *mismatches = (long long)h[0];
  if (outside) *outside = (long long)h[1];
  return e == cudaSuccess ? 0 : 1;
}

void pf_tma2_release(FusedArrays &A) {
  delete[] static_cast<Maps2 *>(A.tma2_cache);
  A.tma2_cache = nullptr;
}
