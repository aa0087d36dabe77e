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
// streaming 16-byte load of an operand nobody else on this SM reads
__device__ __forceinline__ double2 ldg2(const double *p, bool ok) {
  double2 v = make_double2(0., 0.);
  if (ok) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// the pair's own cells, their periodic row images (dj) and plane images; `mask` bit 0 / 1 = element a / b.
// `dst` / `img` point at element a in the output array / in the array that holds the plane images (this array on
// one rank, the neighbour rank's array over NVLink on a z-slab; null = none), `dk` = offset of the image plane.
__device__ __forceinline__ void store_pair(double *dst, double *img, int dj, int dk, double2 v, int mask) {
  if (mask == 3) {
    *reinterpret_cast<double2 *>(dst) = v;
    if (dj) *reinterpret_cast<double2 *>(dst + dj) = v;
    if (img) {
      *reinterpret_cast<double2 *>(img + dk) = v;
      if (dj) *reinterpret_cast<double2 *>(img + dk + dj) = v;
    }
  } else if (mask) {
    const int o = mask >> 1;   // mask 1 -> element a, mask 2 -> element b
    store_with_images(dst, img, o, dj, dk, o ? v.y : v.x);
  }
}

// the SOR update of one cell (ibm_3d_uniform_omp_cpu.f90:510-515, left to right, ap from the raw coefficients :402)
__device__ __forceinline__ double update(double bb, double ae, double aw, double an, double as, double at, double ab,
                                         double pE, double pW, double pN, double pS, double pT, double pB, double pold,
                                         double relux, double omr) {
  const double ap = -ae - aw - an - as - at - ab;
  const double r = bb - ae * pE - aw * pW - an * pN - as * pS - at * pT - ab * pB;
  return r / ap * relux + pold * omr;
}
// the same for a cell of the inlet (i == 1) or outlet (i == m) column: boundrary_matrix's folds (:640-641, :651-656)
// applied to the coefficients first.  Out of line: only the two x-edge tile columns ever come here, and keeping it
// apart keeps the register copies of the fold out of everybody else's path.
__device__ __forceinline__ double update_folded(double bb, double ae, double aw, double an, double as, double at, double ab,
                                             double pE, double pW, double pN, double pS, double pT, double pB,
                                             double pold, double relux, double omr, int i, int m) {
  return sor_update(bb, ae, aw, an, as, at, ab, pE, pW, pN, pS, pT, pB, pold, relux, omr, i, m);
}

struct Own {          // the operands of one z-step that only this thread reads
  double2 p0, bb0, cz0;   // plane k   : old red pressure, red source, at of red(k)
  double2 cz1, bb1;       // plane k-1 : ab of red(k) == at of black(k-1), black source
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

  // ---------------------------------------------------- compute threads
  // warp w holds tile rows w and w+8; lane%16 = pair index along x
  const int lane = tid & 31, wrp = tid >> 5;
  const int e = 2 * (lane & 15), ty = wrp + 8 * (lane >> 4);
  const int ih = h0 + e;                               // element index of the pair's first element (even)
  const int j = j0 + ty;
  const int m = g.m;
  const double omr = 1. - relux;
  int dj = (j <= 2) ? g.n * g.HX : ((j >= g.n - 1) ? -g.n * g.HX : 0);
  // Everything that does not depend on k is computed ONCE and pinned in registers (the empty asm statements keep
  // ptxas from rematerialising a dozen integer instructions per use inside the z-loop).
  //   f[S], S = parity of i in this row: bit 0/1 = element a/b is a cell (i in 1..m), bit 2/3 = a/b is stored by
  //   this thread (owned and a cell), bit 4 = an inlet/outlet fold (i == 1 or i == m) applies to a or b
  int f0, f1;
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
      return (ca ? 1 : 0) | (cb ? 2 : 0) | (ca && own_a ? 4 : 0) | (cb && own_b ? 8 : 0) | (fold ? 16 : 0);
    };
    f0 = flags(0);
    f1 = flags(1);
  }
  const int sj = (j + g.koff) & 1;                     // + k = parity of i in this row
#ifdef PF_TMA2_NOLOAD
  const bool ld_ok = false;
#else
  const bool ld_ok = j + 1 <= g.n + 3 && g.H0 + ih + 1 <= g.HX - 1;   // the pair lies inside the arrays
#endif
  const uint32_t sb = smem_u32(smem);
  uint32_t oN = (uint32_t)(ty * TW + e) * 8;                 // narrow box
  uint32_t oW = (uint32_t)(ty * TWP + e + 2) * 8;            // wide group box / R slot
  uint32_t oP = (uint32_t)((ty + 1) * TWP + e + 2) * 8;      // P1 box
  asm volatile("" : "+r"(dj), "+r"(oN), "+r"(oW), "+r"(oP));
  const uint32_t gbase = sb, pbase = sb + NG * SZ_GROUP, rbase = pbase + NP * SZ_P1;

  // element offset of the pair at the plane being LOADED / at the plane of the current step; one plane per step
  const long long pstride = A.hplane2;
  long long ol = (long long)g.H0 + ih + (long long)g.HX * (j + 1) + pstride * (kfirst + 1);
  long long os = ol;

  auto load_own = [&](Own &o) {           // operands of plane `ol` (and of the plane below it)
    o.p0 = ldg2(A.pin0 + ol, ld_ok);
    o.bb0 = ldg2(A.bb0 + ol, ld_ok);
    o.cz0 = ldg2(A.cz0 + ol, ld_ok);
    o.cz1 = ldg2(A.cz1 + ol - pstride, ld_ok);
    o.bb1 = ldg2(A.bb1 + ol - pstride, ld_ok);
    ol += pstride;
  };

  double2 cza = make_double2(0., 0.), czb = cza;   // cz0 of this pair at planes k-2, k-1 (ab of black(k-1) = cz0(k-2))
  double emax = 0.;
  int gs = 0, gphase = 0;                 // slot of group(k) and the phase bit of its mbarrier
  int pc = 2, pphase = 0;                 // slot of P1(k+1)
  int rs = 0;                             // slot of R(k)
  uint32_t gK = gbase, gKb = gbase;
  uint32_t pA = pbase, pB = pbase + SZ_P1, pC = pbase + 2 * SZ_P1;
  uint32_t rK = rbase, rKb = rbase, rKb1 = rbase;

  Own oa, ob;
  load_own(oa);
#ifndef PF_TMA2_NOLOAD
  mbar_wait(&pbar[0], 0);
  mbar_wait(&pbar[1], 0);
#endif

  // one z-step: red stage of plane k, block barrier, black stage of plane k-1
  auto step = [&](int k, const Own &o, const int f, const uint32_t xo, const int ia) {
    // f: this row's flags at this step's parity S of i; west/east neighbours of the pair in the other colour's row
    // = three consecutive elements starting at e-1 (S=1, xo = -8) or e (S=0, xo = 0):  a -> {x0, x1},  b -> {x1, x2}
#ifdef PF_TMA2_NOFOLD
    const bool folded = false;
#else
    const bool folded = f & 16;
#endif
#ifndef PF_TMA2_NOLOAD
    mbar_wait(&gbar[gs], gphase);
    mbar_wait(&pbar[pc], pphase);
#endif
    // ------------------------------------------ red stage, plane k
    double2 val = o.p0;
    {
      const uint32_t px = pB + oP + xo, cx = gK + OFF_CX1 + oW + xo;
      const double x0 = lds(px), x1 = lds(px + 8), x2 = lds(px + 16);
      const double w0c = lds(cx), w1c = lds(cx + 8);   // east faces of the west neighbours = the pair's west faces
      const double2 ae = lds2(gK + OFF_CX0 + oW);      // own east faces
      const double2 an = lds2(gK + OFF_CY0 + oN);
      const double2 as = lds2(gK + OFF_CY1 + oN);      // row j-1: the box starts at row j0-1
      const double2 pN = lds2(pB + oP + TWP * 8), pS = lds2(pB + oP - TWP * 8);
      const double2 pT = lds2(pC + oP), pBt = lds2(pA + oP);
#ifndef PF_TMA2_NOCOMPUTE
      if (!folded) {
        if (f & 1)
          val.x = update(o.bb0.x, ae.x, w0c, an.x, as.x, o.cz0.x, o.cz1.x, x1, x0, pN.x, pS.x, pT.x, pBt.x, o.p0.x, relux, omr);
        if (f & 2)
          val.y = update(o.bb0.y, ae.y, w1c, an.y, as.y, o.cz0.y, o.cz1.y, x2, x1, pN.y, pS.y, pT.y, pBt.y, o.p0.y, relux, omr);
      } else {
        if (f & 1)
          val.x = update_folded(o.bb0.x, ae.x, w0c, an.x, as.x, o.cz0.x, o.cz1.x, x1, x0, pN.x, pS.x, pT.x, pBt.x, o.p0.x,
                                relux, omr, ia, m);
        if (f & 2)
          val.y = update_folded(o.bb0.y, ae.y, w1c, an.y, as.y, o.cz0.y, o.cz1.y, x2, x1, pN.y, pS.y, pT.y, pBt.y, o.p0.y,
                                relux, omr, ia + 2, m);
      }
#else
      val.x += x0 + x1 + x2 + w0c + w1c + ae.x + an.x + as.x + pN.x + pS.x + pT.x + pBt.x;   // keep the loads
#endif
      if (k >= kc0 && k <= kc1) {
        const bool lo = k <= 2, hi = k >= g.lz - 1;
        double *img = lo ? A.ilo0 : (hi ? A.ihi0 : nullptr);
        store_pair(A.pout0 + os, img ? img + os : nullptr, dj, lo ? A.dk_lo : A.dk_hi, val, (f >> 2) & 3);
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
    // ------------------------------------------ black stage, plane k-1
    const int st = (f >> 2) & 3;
    if (k - 1 >= kc0 && k - 1 <= kc1 && st) {
      const uint32_t rx = rKb + oW + xo, cx = gKb + OFF_CX0 + oW + xo;
      const double2 bold = lds2(pA + oP);              // black own old value (plane k-1)
      const double x0 = lds(rx), x1 = lds(rx + 8), x2 = lds(rx + 16);
      const double w0c = lds(cx), w1c = lds(cx + 8);
      const double2 ae = lds2(gKb + OFF_CX1 + oW);
      const double2 an = lds2(gKb + OFF_CY1 + oN + TW * 8);   // own row j
      const double2 as = lds2(gKb + OFF_CY0 + oN - TW * 8);   // row j-1
      const double2 pN = lds2(rKb + oW + TWP * 8), pS = lds2(rKb + oW - TWP * 8);
      const double2 pBt = lds2(rKb1 + oW);
      double2 v = bold;
#ifndef PF_TMA2_NOCOMPUTE
      if (!folded) {
        if (st & 1)
          v.x = update(o.bb1.x, ae.x, w0c, an.x, as.x, o.cz1.x, cza.x, x1, x0, pN.x, pS.x, val.x, pBt.x, bold.x, relux, omr);
        if (st & 2)
          v.y = update(o.bb1.y, ae.y, w1c, an.y, as.y, o.cz1.y, cza.y, x2, x1, pN.y, pS.y, val.y, pBt.y, bold.y, relux, omr);
      } else {
        if (st & 1)
          v.x = update_folded(o.bb1.x, ae.x, w0c, an.x, as.x, o.cz1.x, cza.x, x1, x0, pN.x, pS.x, val.x, pBt.x, bold.x,
                              relux, omr, ia, m);
        if (st & 2)
          v.y = update_folded(o.bb1.y, ae.y, w1c, an.y, as.y, o.cz1.y, cza.y, x2, x1, pN.y, pS.y, val.y, pBt.y, bold.y,
                              relux, omr, ia + 2, m);
      }
#else
      v.x += x0 + x1 + x2 + w0c + w1c + ae.x + an.x + as.x + pN.x + pS.x + pBt.x;
#endif
      // unowned elements keep v == bold: they add 0 to the error
      emax = fmax(emax, fmax(fabs(v.x - bold.x), fabs(v.y - bold.y)));
      const bool lo = k - 1 <= 2, hi = k - 1 >= g.lz - 1;
      double *img = lo ? A.ilo1 : (hi ? A.ihi1 : nullptr);
      store_pair(A.pout1 + os - pstride, img ? img + os - pstride : nullptr, dj, lo ? A.dk_lo : A.dk_hi, v, st);
    }
    // next plane: rotate the slots, advance the destinations
    cza = czb; czb = o.cz0;
    os += pstride;
    if (++gs == NG) { gs = 0; gphase ^= 1; }
    gKb = gK; gK = gbase + (uint32_t)gs * SZ_GROUP;
    pA = pB; pB = pC;
    if (++pc == NP) { pc = 0; pphase ^= 1; }
    pC = pbase + (uint32_t)pc * SZ_P1;
    if (++rs == NR) rs = 0;
    rKb1 = rKb; rKb = rK; rK = rbase + (uint32_t)rs * SZ_R;
  };

  // Two steps per trip.  The operand sets alternate between two register files (no copies; the loads of a step's
  // successor are in flight during the whole step), and since the parity of i alternates with k, each of the two
  // step bodies always sees the same parity: its flags, neighbour offset and i are loop constants.
  {
    const int SA = (sj + kfirst) & 1;
    int fA = SA ? f1 : f0, fB = SA ? f0 : f1;
    uint32_t xA = SA ? (uint32_t)-8 : 0u, xB = SA ? 0u : (uint32_t)-8;
    asm volatile("" : "+r"(fA), "+r"(fB), "+r"(xA), "+r"(xB));
    const int iA = 2 * ih + 2 - SA, iB = 2 * ih + 1 + SA;
    for (int k = kfirst; k <= klast; k += 2) {
      if (k < klast) load_own(ob);
      step(k, oa, fA, xA, iA);
      if (k + 1 > klast) break;
      if (k + 1 < klast) load_own(oa);
      step(k + 1, ob, fB, xB, iB);
    }
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

void pf_tma2_release(FusedArrays &A) {
  delete[] static_cast<Maps2 *>(A.tma2_cache);
  A.tma2_cache = nullptr;
}
