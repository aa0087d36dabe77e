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
//   group(p) : P0 BB0 CX0 CX1 CY0 CY1 CZ0 CZ1 BB1          (NG = 5 slots; used by red(p) and, one step later, black(p))
//   P1(p)    : the old black pressure, box widened by the ring   (5 slots; read in-plane by red(p), its own column
//              by red(p-1) and red(p+1) and as black(p)'s old value -- those through registers)
// and a 2-slot ring R of the new red values (227 KB of shared memory in total, one block per SM).  group(p) and
// P1(p+1) -- what step p needs that no earlier step needed -- travel as one bundle on one mbarrier, issued three
// steps ahead.  Arithmetic is the same sor_update() as everywhere else.  Plane images (periodic wrap on one rank, the
// neighbour rank's ghost planes on a z-slab) are stored by the thread that owns the cell.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

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
#define PF_TMA_NR 2
#endif
constexpr int NG = PF_TMA_NG;   // group slots: planes k-1, k in use, the rest landed / in flight (lead = NG-2 steps)
constexpr int NP = 5;         // P1 slots
constexpr int NR = PF_TMA_NR;   // R slots: plane k-1 read by the neighbours, plane k written during step k
#ifndef PF_TMA_MINB
#define PF_TMA_MINB 1           // resident blocks per SM the kernel is compiled for
#endif
// (registers: 17 warps are allocated as 20 -- the granularity is four warps -- so ptxas stops at 96 per thread; 120
// would need the producer folded into a compute warp.  At 96 the steady-state loop does not spill.)
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
constexpr int SMEM_BYTES = NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R + 512 + 128;   // + mbarriers, row flags, block-max scratch, alignment slack

template <int V>
struct ConstInt {   // a step index known at compile time (std::integral_constant's conversion is host-only)
  __host__ __device__ constexpr operator int() const { return V; }
};

struct TmaMaps {
  CUtensorMap p0, p1, cx0, cx1, cy0, cy1, cz0, cz1, bb0, bb1;
};

struct TmaArgs {
  int NY2, hplane2;
  // block -> (tile, z-chunk): the first tA tiles (in x-fastest order) are cut into nzA z-chunks each, the rest into nzB;
  // the blocks of the first group come first (pf_tma_schedule below).  xt = tiles per row of tiles, ntiles = all tiles
  int xt, ntiles, tA, nzA, nzB;
  double *pout0, *pout1;
  double *ilo0, *ilo1, *ihi0, *ihi1;   // image destinations of planes 1,2 / lz-1,lz (FusedArrays::img_lo / img_hi)
  int dk_lo, dk_hi;
  // z-slab ranks with the peer-store transport: the neighbour handshake lives IN this kernel (see slab_sync below).
  // `sync` = this rank's flag words (null: single rank, or the NCCL transport), to_prev / to_next = the words of the
  // neighbours this rank publishes into.
  unsigned long long *sync, *to_prev, *to_next;
};

// ---- neighbour handshake of the z-slab ranks, inside the sweep kernel ---------------------------------------------
// Launch n of a rank (n counts its sweep launches since the solver was created; all ranks launch in step) reads, in
// its bottom z-chunk, the ghost planes -1, 0 that the previous rank's TOP chunk stored during ITS launch n-1, and
// stores the images of its planes 1, 2 into that rank's ghost planes lz+1, lz+2 of the buffer the neighbour was still
// reading during launch n-1; the top chunk does the same with the next rank.  So:
//   * a block of the bottom (top) chunk starts only when the previous (next) rank has published "my top (bottom)
//     chunk has finished launch n-1"; every other block starts at once -- the interior overlaps the handshake;
//   * the block that finishes a boundary chunk LAST publishes "launch n" into the neighbour, after a system-scope
//     fence, so its peer stores are visible there first.  The boundary chunks are scheduled first (z index 0 and 1),
//     so that publication happens early in the launch, long before the interior chunks are done.
// No separate barrier kernel, nothing on the host: the whole iteration loop replays from a CUDA graph.
// Flag words (unsigned long long, in the rank's peer-visible block; [0], [1] belong to the solve-start barrier):
constexpr int SY_FROM_PREV = 8, SY_FROM_NEXT = 9;   // written by the neighbours: their launch number
constexpr int SY_LAUNCH = 10;                       // launches of this rank completed so far
constexpr int SY_CNT_BOT = 11, SY_CNT_TOP = 12, SY_CNT_ALL = 13;   // blocks finished in this launch

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// thread 0 of a boundary-chunk block, before the block touches a ghost plane or a neighbour's memory
__device__ __forceinline__ void slab_wait(const unsigned long long *flag, unsigned long long want) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(flag) < want) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 60000000000ull) __trap();   // 60 s: the neighbour died; fail the launch instead of hanging the GPU
  }
}

// 512 compute threads (one checkerboard element of the 32x16 tile each, 16 warps to hide the fp64
// dependency chains) + one producer warp whose lane 0 issues the TMA copies NG-2 planes ahead.
// Synchronisation inside the z-loop is point to point (no block-wide barrier): TMA "full" mbarriers per slot, one
// "every warp has finished step q" mbarrier per q % 5 for the producer, one progress word per tile row for the rows
// above and below.  Named barrier 1 = the compute threads, once, for the block maximum at the end.
__global__ void __launch_bounds__(NTHREADS, PF_TMA_MINB) sor_tma_kernel(const __grid_constant__ TmaMaps M, Geo g, TmaArgs A,
                                                              double relux, double omr, unsigned long long *err_bits) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  unsigned char *grp = smem;                                  // NG x SZ_GROUP
  unsigned char *p1s = smem + NG * SZ_GROUP;                  // NP x SZ_P1
  double *Rs = reinterpret_cast<double *>(smem + NG * SZ_GROUP + NP * SZ_P1);   // NR x TR x TWP
  uint64_t *gbar = reinterpret_cast<uint64_t *>(smem + NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R);   // NG
  uint64_t *pbar = gbar + NG;                                                                     // NP
  uint64_t *ebar = pbar + NP;                                     // 5: "every compute warp has finished step q" (q % 5)
  double *wmax = reinterpret_cast<double *>(ebar + 5 + 1);        // NCOMPUTE / 32 block-max scratch
  uint32_t *rowdone = reinterpret_cast<uint32_t *>(wmax + NCOMPUTE / 32);   // per compute warp: steps completed

  const int tid = threadIdx.x;
  // tile and z-chunk of this block
  int nzc, tile, zc;
  {
    int b = (int)blockIdx.x;
    const int nA = A.tA * A.nzA;
    if (b < nA) { nzc = A.nzA; tile = b / nzc; zc = b - tile * nzc; }
    else        { b -= nA; nzc = A.nzB; const int t = b / nzc; zc = b - t * nzc; tile = A.tA + t; }
  }
  const int tby = tile / A.xt, tbx = tile - tby * A.xt;
  const int h0 = tbx * (TW - 2) - 2;                 // element index of tile column 0 (stride TW-2, even)
  const int j0 = tby * (TR - 2);                     // ext row 0 of the tile; owned rows j0+1 .. j0+6
  const int czp = (g.lz + nzc - 1) / nzc;            // planes per z-chunk of this tile
  const bool slab_bot = A.sync && zc == 0, slab_top = A.sync && zc == nzc - 1;
  const int kc0 = zc * czp + 1;
  const int kc1 = min(kc0 + czp - 1, g.lz);
  const int kfirst = kc0 - 1, klast = kc1 + 1;       // red planes
  // array coordinates of the boxes
  const int xn = g.H0 + h0, xw = xn - 2;             // narrow / wide box column origin
  const int yn = j0 + 1, ym = j0;                    // box row origin: rows j0.. / rows j0-1..
  if (tid == 0) {
    for (int q = 0; q < NG; ++q) mbar_init(&gbar[q], 1);
    for (int q = 0; q < NP; ++q) mbar_init(&pbar[q], 1);
    for (int q = 0; q < 5; ++q) mbar_init(&ebar[q], NCOMPUTE / 32);
    mbar_init(&ebar[5], NCOMPUTE / 32);   // "every compute warp has read P1(kfirst-1)": its slot is the first to be refilled
    for (int q = 0; q < NCOMPUTE / 32; ++q) rowdone[q] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (slab_bot || slab_top) {   // the neighbour(s) this chunk exchanges planes with have finished the previous launch
      const unsigned long long n = A.sync[SY_LAUNCH] + 1;
      if (slab_bot) slab_wait(A.sync + SY_FROM_PREV, n - 1);
      if (slab_top) slab_wait(A.sync + SY_FROM_NEXT, n - 1);
      asm volatile("fence.proxy.async.global;" ::: "memory");   // the copies below read what the neighbour stored
    }
  }
  __syncthreads();

#ifdef PF_TMA_NOLOAD
  // experiment: benign operands everywhere (no division slow path on garbage)
  for (int q = tid; q < (NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R) / 8; q += NTHREADS) reinterpret_cast<double *>(smem)[q] = 1.25 + 1e-3 * (q & 63);
  __syncthreads();
#endif
  // roles: warps 0..15 compute, warp 16 produces (its lane 0 issues the TMA copies).  Each role has its own z-loop;
  // they meet through mbarriers only (copies landed / operands consumed), the compute rows through progress words.
  const bool is_producer = tid >= NCOMPUTE;
#ifdef PF_TMA_NOLOAD
  const bool lead = false;
#else
  const bool lead = tid == NCOMPUTE;
#endif
  {
    // bundle(p) = group(p) + P1(p+1): what step p needs that no earlier step needed, on ONE mbarrier
    auto issue_bundle = [&](int p) {
      const int q = p - kfirst;
      unsigned char *b = grp + (q % NG) * SZ_GROUP;
      uint64_t *bar = &gbar[q % NG];
      const int z = p + 1;
      mbar_expect_tx(bar, GROUP_BYTES + P1_BYTES);
      tma_load_3d(p1s + ((q + 2) % NP) * SZ_P1, &M.p1, bar, xw, ym, p + 2);
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
    auto issue_p1 = [&](int p) {   // the two planes the first step reads besides its bundle
      const int q = p - (kfirst - 1);
      uint64_t *bar = &pbar[q % NP];
      mbar_expect_tx(bar, P1_BYTES);
      tma_load_3d(p1s + (q % NP) * SZ_P1, &M.p1, bar, xw, ym, p + 1);
    };
    if (lead) {   // prologue: P1 planes kfirst-1, kfirst ; bundles kfirst .. kfirst+GLEAD-1
      issue_p1(kfirst - 1);
      issue_p1(kfirst);
      for (int d = 0; d < GLEAD; ++d)
        if (kfirst + d <= klast) issue_bundle(kfirst + d);
    }
  }
  // lane 0 of the producer warp, once per step: bundle(k+GLEAD) = group(k+GLEAD) into group slot gslot and
  // P1(k+GLEAD+1) into P1 slot pslot
  auto produce = [&](int k, auto gslot, auto pslot) {
    if (k + GLEAD <= klast) {
      unsigned char *b = grp + (int)gslot * SZ_GROUP;
      uint64_t *bar = &gbar[(int)gslot];
      const int z = k + GLEAD + 1;
      mbar_expect_tx(bar, GROUP_BYTES + P1_BYTES);
      tma_load_3d(p1s + (int)pslot * SZ_P1, &M.p1, bar, xw, ym, z + 1);
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
  };

  // ---------------------------------------------------- compute threads
  // Step k (k = kfirst .. klast) updates red on plane k and black on plane k-1 as ONE straight line of two interleaved
  // instruction streams.  black(k-1) reads the red values of plane k-1 from its four in-plane neighbours (slot R(k-1):
  // its own row's since its own step k-1, the rows above and below once their progress words say so), of plane k-2
  // and plane k from this thread's own registers.
  // Instruction diet (round 2, profiles/r02_sor_summary.md: the loop was issue- and latency-bound, ~270 instructions
  // per thread and step): step q = k - kfirst uses group slot q % 5 and P1 slots (q+1) % 5, (q+2) % 5, so the steady
  // state is unrolled by ten -- every shared-memory operand is  per-thread base register + compile-time immediate,
  // the mbarrier parities are one bit per round of five -- the global store pointers advance by one plane per step,
  // the own-column values of P1 and R travel in registers (two loads and one R slot less), the planes near the chunk
  // and domain faces (store predicates, plane images) run a generic flavour of the same step, and the plain `/` of a
  // division outside quot_guard() lives in a function that is not inlined (the compiler's inline division expanded
  // its whole fast path next to quot_fast() just to select between them).
  const int e = tid % TW, ty = tid / TW;             // tile column / row of this thread
  const int ih = h0 + e;
  const int j = j0 + ty;
  const int m = g.m;
  const int sj = (j + g.koff) & 1;                  // (omr = 1 - relux arrives as a kernel parameter: a constant-bank operand)
  const int cbase = g.H0 + ih + g.HX * (j + 1);       // + hplane2*(k+1) = global element index
  // flags per parity s of i (i = 2*ih + 2 - s; s alternates with k):
  //   bit 0 = the element is a cell (i in 1..m), bit 1 = this thread stores it (owned and a cell),
  //   bit 2 = inlet column (i == 1), bit 3 = outlet column (i == m)
  int fl0, fl1;
  {
    const int ihmax = (m + 1) >> 1;
    const bool in_dom = j <= g.n + 1 && ih >= -1 && ih <= ihmax;
    const bool own = ty >= 1 && ty <= TR - 2 && j >= 1 && j <= g.n && e >= 1 && e <= TW - 2;
    auto flags = [&](int sp) {
      const int i = 2 * ih + 2 - sp;
      const bool c = in_dom && i >= 1 && i <= m;
      return (c ? 1 : 0) | (c && own ? 2 : 0) | (c && i == 1 ? 4 : 0) | (c && i == m ? 8 : 0);
    };
    fl0 = flags(0);
    fl1 = flags(1);
  }
  // the row image of a freshly stored cell (rows 1,2 <-> n+1,n+2; rows n-1,n <-> -1,0), in elements; 0 = none
  const int dj = j <= 2 ? g.n * g.HX : (j >= g.n - 1 ? -g.n * g.HX : 0);
  const uint32_t sb = smem_u32(smem);
  constexpr uint32_t ROW = TWP * 8;
  constexpr uint32_t PBASE = NG * SZ_GROUP, RBASE = PBASE + NP * SZ_P1;
  uint32_t bN = sb + (uint32_t)(ty * TW + e) * 8;              // narrow boxes
  uint32_t bW = sb + (uint32_t)(ty * TWP + e + 2) * 8;         // wide group boxes / R slots; the P1 box: + one row
  asm volatile("" : "+r"(bN), "+r"(bW));                       // pinned: not rematerialised inside the z-loop
  // parity s of i (red row at plane k, black row at plane k-1) alternates with k; the west/east neighbours in the
  // other colour's row are s = 1 -> {e-1, e}, s = 0 -> {e, e+1}
  const int s_first = (sj + kfirst) & 1;
  const uint32_t rA = bW + RBASE, rB = bW + RBASE + SZ_R;   // the two R slots (NR == 2)
  double *po0 = A.pout0 + (cbase + (long long)A.hplane2 * (kfirst + 1));   // red cell of plane k
  double *po1 = A.pout1 + (cbase + (long long)A.hplane2 * kfirst);         // black cell of plane k-1
  double cza = 0., czb = 0.;                          // cz_red of this element at planes k-2, k-1
  double vb = 0., vbb = 0.;                           // the red values this thread computed for planes k-1, k-2
  double pb = 0., pbn = 0.;                           // P1 of this column at planes k-1 and k
  double emax = 0.;

#ifndef PF_TMA_NOLOAD
  if (!is_producer) {
    mbar_wait(&pbar[0], 0);
    mbar_wait(&pbar[1], 0);
    pb = lds(bW + PBASE + ROW);                        // P1(kfirst-1), P1(kfirst) of this column
    pbn = lds(bW + PBASE + SZ_P1 + ROW);
    __syncwarp();
    if (e == 0) mbar_arrive(&ebar[5]);                 // slot 0 of P1 may be refilled (the producer's first bundle)
  }
#endif

  // images of a freshly stored cell near a domain face in z (planes 1,2 / lz-1,lz: the periodic wrap on one rank, the
  // neighbour rank's ghost planes over NVLink on a z-slab; null = exchanged after the launch), with their row images
  auto store_plane_images = [&](double *img, long long c, int dk, double v) {
    if (img) {
      img[c + dk] = v;
      if (dj) img[c + dk + dj] = v;
    }
  };

  // MID (block-uniform): chunk-interior, domain-interior step -- both colours are stored by every owning thread and
  //   nothing has a plane image.
  // EDGE (block-uniform): the tile holds the inlet column i == 1 or the outlet column i == m, whose cells fold
  //   boundrary_matrix into their coefficients (:640-641, :651-656) -- with selects, the same operations on the same
  //   values as sor_update().
  // u = (k - kfirst) % 5 as a compile-time constant (steady state) or an int; par = parity of the round of five
  const uint32_t flag_me = smem_u32(rowdone) + (uint32_t)ty * 4u, flag_dn = flag_me - 4u, flag_up = flag_me + 4u;
  auto row_wait = [&](uint32_t flag, uint32_t want) {
    if (ld_acquire_shared(flag) >= want) return;
    for (unsigned spins = 0; ld_acquire_shared(flag) < want; ++spins)
      if (spins > (1u << 24)) __trap();   // never hang the GPU
  };
  // f, bWs_, rK_, rKb_: the quantities that alternate from step to step (flags of this parity of i, west-shifted
  // base, the R slot written / read) -- variables of the generic loop, compile-time renames in the unrolled one
  auto step = [&](auto MID_, auto EDGE_, const int k, auto u, const uint32_t par, const int f, const uint32_t bWs_,
                  const uint32_t rK_, const uint32_t rKb_) {
    constexpr bool MID = decltype(MID_)::value;
    constexpr bool EDGE = decltype(EDGE_)::value;
    const int su = (int)u;
    auto wrap5 = [](int x) { return x >= 5 ? x - 5 : x; };
    const int sG = su, sGb = wrap5(su + 4), sPk = wrap5(su + 1), sPn = wrap5(su + 2);
    {
#ifndef PF_TMA_NOLOAD
      mbar_wait(&gbar[sG], par);                        // group(k) and P1(k+1) have landed
#endif
      const uint32_t G = (uint32_t)sG * SZ_GROUP, Gb = (uint32_t)sGb * SZ_GROUP;
      const uint32_t Pk = PBASE + (uint32_t)sPk * SZ_P1, Pn = PBASE + (uint32_t)sPn * SZ_P1;
      const uint32_t bWs = bWs_, rK = rK_, rKb = rKb_;
      const uint32_t rKbs = rKb + (bWs - bW);
      // ------------------------------------------ operands: red stage (plane k), black stage (plane k-1)
      const double pold = lds(bN + G + OFF_P0), bb0 = lds(bN + G + OFF_BB0);
      double ae0 = lds(bW + G + OFF_CX0), aw0 = lds(bWs + G + OFF_CX1);
      double an0 = lds(bN + G + OFF_CY0), as0 = lds(bN + G + OFF_CY1);      // row j-1: the box starts at row j0-1
      double at0 = lds(bN + G + OFF_CZ0), ab0 = lds(bN + G + OFF_CZ1);      // cz1 of plane k-1
      const double czk = at0;
      const double pE0 = lds(bWs + Pk + ROW + 8), pW0 = lds(bWs + Pk + ROW);
      const double pN0 = lds(bW + Pk + 2 * ROW), pS0 = lds(bW + Pk);
      const double pT0 = lds(bW + Pn + ROW), pB0 = pb;
      const double bold = pb;                                               // black own old value (plane k-1)
      const double bb1 = lds(bN + Gb + OFF_BB1);
      double ae1 = lds(bW + Gb + OFF_CX1), aw1 = lds(bWs + Gb + OFF_CX0);
      double an1 = lds(bN + Gb + OFF_CY1 + TW * 8), as1 = lds(bN + Gb + OFF_CY0 - TW * 8);   // own row j / row j-1
      double at1 = ab0, ab1 = cza;                                          // cz1(k-1), cz0(k-2)
      // the red values of plane k-1 in the rows above and below are the neighbour warps': they have finished step
      // k-1 (published R(k-1), and read the last of R(k-2), whose slot R(k) overwrites below)
#ifndef PF_TMA_BLOCKBAR
      // every TMA-filled operand of this step is in a register (the arrive below is a release: it is ordered after
      // the loads above): tell the producer, which refills the slots of plane k-1 once all sixteen warps have arrived
      // -- most of a step earlier than "step finished" would, and every step of lead hides DRAM latency.  No proxy
      // fence: the slots were only READ by this block (the consumer release of any TMA pipeline).
      __syncwarp();
      if (e == 0) mbar_arrive(&ebar[su]);
      {
        const uint32_t q = (uint32_t)(k - kfirst);
        if (ty > 0) row_wait(flag_dn, q);
        if (ty < TR - 1) row_wait(flag_up, q);
      }
#endif
      const double pE1 = lds(rKbs + 8), pW1 = lds(rKbs);
      const double pN1 = lds(rKb + ROW), pS1 = lds(rKb - ROW);
      const double pB1 = vbb;
#ifndef PF_TMA_NOCOMPUTE
      // ------------------------------------------ red update (ibm_3d_uniform_omp_cpu.f90:510-515; ap :402, raw)
      const double ap0 = -ae0 - aw0 - an0 - as0 - at0 - ab0;
      const double ap1 = -ae1 - aw1 - an1 - as1 - at1 - ab1;
      if (EDGE) {
        if (f & 4) { ae0 = ae0 + aw0; aw0 = 0.; ae1 = ae1 + aw1; aw1 = 0.; }
        if (f & 8) { ae0 = aw0 = an0 = as0 = at0 = ab0 = 0.; ae1 = aw1 = an1 = as1 = at1 = ab1 = 0.; }
      }
      const double r0 = bb0 - ae0 * pE0 - aw0 * pW0 - an0 * pN0 - as0 * pS0 - at0 * pT0 - ab0 * pB0;
      double q0 = quot_fast(r0, ap0);
      if ((f & 1) && !quot_guard(r0, ap0)) q0 = quot_plain(r0, ap0);      // exact zeros, denormals: rare
      const double val = (f & 1) ? q0 * relux + pold * omr : pold;         // halo / out-of-domain slots pass through
      // ------------------------------------------ black update, plane k-1 (its top neighbour is `val`)
      const double r1 = bb1 - ae1 * pE1 - aw1 * pW1 - an1 * pN1 - as1 * pS1 - at1 * val - ab1 * pB1;
      double q1 = quot_fast(r1, ap1);
      const bool black = MID ? (f & 2) != 0 : (f & 2) && k - 1 >= kc0 && k - 1 <= kc1;
      if (black && !quot_guard(r1, ap1)) q1 = quot_plain(r1, ap1);
      const double v = q1 * relux + bold * omr;
#else
      const double val = pold + bb0 + ae0 + aw0 + an0 + as0 + at0 + ab0 + pE0 + pW0 + pN0 + pS0 + pT0 + pB0;
      const double v = bold + bb1 + ae1 + aw1 + an1 + as1 + at1 + ab1 + pE1 + pW1 + pN1 + pS1 + pB1;
      const bool black = MID ? (f & 2) != 0 : (f & 2) && k - 1 >= kc0 && k - 1 <= kc1;
#endif
      // ------------------------------------------ stores
      sts(rK, val);
      if (MID ? (f & 2) != 0 : (f & 2) && k >= kc0 && k <= kc1) {
        *po0 = val;
        if (dj) po0[dj] = val;
        if (!MID) {
          const bool lo = k <= 2, hi = k >= g.lz - 1;
          if (lo || hi)
            store_plane_images(lo ? A.ilo0 : A.ihi0, cbase + (long long)A.hplane2 * (k + 1), lo ? A.dk_lo : A.dk_hi, val);
        }
      }
      if (black) {
        *po1 = v;
        if (dj) po1[dj] = v;
        if (!MID) {
          const bool lo = k - 1 <= 2, hi = k - 1 >= g.lz - 1;
          if (lo || hi)
            store_plane_images(lo ? A.ilo1 : A.ihi1, cbase + (long long)A.hplane2 * k, lo ? A.dk_lo : A.dk_hi, v);
        }
        emax = fmax(emax, fabs(v - bold));
      }
      // histories and per-step toggles
      cza = czb; czb = czk;
      vbb = vb; vb = val;
      pb = pbn; pbn = pT0;
      po0 += A.hplane2; po1 += A.hplane2;
      // this warp has finished step k: its row of R(k) is written and it has read the last of R(k-1).  Tell the rows
      // above and below.  No block-wide barrier: rows may drift apart by a step, the block by the depth of the pipeline.
#ifndef PF_TMA_BLOCKBAR
      __syncwarp();
      if (e == 0) st_release_shared(flag_me, (uint32_t)(k - kfirst) + 1u);
#else
      named_bar(2, NTHREADS);   // experiment: the block-wide barrier per step this scheme replaced
#endif
    }
  };

  auto sweep = [&](auto EDGE_) {
    int k = kfirst, u = 0, sp = s_first;
    uint32_t par = 0, rK = rA, rKb = rB;              // R(k) is written, R(k-1) is read by the neighbours
    // steady state: kc0 + 1 <= k <= kc1 (both colours stored), red plane k in 3 .. lz-2 and black plane k-1 in
    // 3 .. lz-2 (no plane images)
    const int mid_lo = max(kc0 + 1, 4), mid_hi = min(kc1, g.lz - 2);
    auto generic = [&]() {
      step(std::false_type{}, EDGE_, k, u, par, sp ? fl1 : fl0, bW - (sp ? 8u : 0u), rK, rKb);
      ++k;
      sp ^= 1;
      if (++u == 5) { u = 0; par ^= 1u; }
      { const uint32_t t = rK; rK = rKb; rKb = t; }
    };
    while (k <= klast && (k < mid_lo || u != 0)) generic();
    if (k + 9 <= mid_hi) {
      const int fE = sp ? fl1 : fl0, fO = sp ? fl0 : fl1;          // flags of the even / odd steps of a round of ten
      const uint32_t wE = bW - (sp ? 8u : 0u), wO = bW - (sp ? 0u : 8u);
      const uint32_t rE = rK, rO = rKb;
      do {
        step(std::true_type{}, EDGE_, k, ConstInt<0>{}, par, fE, wE, rE, rO);
        step(std::true_type{}, EDGE_, k + 1, ConstInt<1>{}, par, fO, wO, rO, rE);
        step(std::true_type{}, EDGE_, k + 2, ConstInt<2>{}, par, fE, wE, rE, rO);
        step(std::true_type{}, EDGE_, k + 3, ConstInt<3>{}, par, fO, wO, rO, rE);
        step(std::true_type{}, EDGE_, k + 4, ConstInt<4>{}, par, fE, wE, rE, rO);
        step(std::true_type{}, EDGE_, k + 5, ConstInt<0>{}, par ^ 1u, fO, wO, rO, rE);
        step(std::true_type{}, EDGE_, k + 6, ConstInt<1>{}, par ^ 1u, fE, wE, rE, rO);
        step(std::true_type{}, EDGE_, k + 7, ConstInt<2>{}, par ^ 1u, fO, wO, rO, rE);
        step(std::true_type{}, EDGE_, k + 8, ConstInt<3>{}, par ^ 1u, fE, wE, rE, rO);
        step(std::true_type{}, EDGE_, k + 9, ConstInt<4>{}, par ^ 1u, fO, wO, rO, rE);
        k += 10;
      } while (k + 9 <= mid_hi);                      // ten steps later: the same u, par, parity and R slots
    }
    while (k <= klast) generic();
  };
  if (is_producer) {
    // the producer walks the same steps with its own (tiny) loop behind the slowest compute warp: the copies issued
    // at step k overwrite the slots of plane k-2, free once every compute warp has loaded its operands of step k-1
#ifdef PF_TMA_BLOCKBAR
    {
      int u = 0;
      if (lead) mbar_wait(&ebar[5], 0);
      for (int k = kfirst; k <= klast; ++k) {
        if (lead) produce(k, u + 3 >= 5 ? u - 2 : u + 3, u);   // P1(k+4) takes the slot of P1(k-1)
        u = u == 4 ? 0 : u + 1;
        named_bar(2, NTHREADS);
      }
      return;
    }
#endif
    if (lead) {
      int u = 0;
      uint32_t epar = 0;
      for (int k = kfirst; k <= klast; ++k) {
        if (k > kfirst) {
          const int up = u == 0 ? 4 : u - 1;
          mbar_wait(&ebar[up], up == 4 ? epar ^ 1u : epar);
        } else {
          mbar_wait(&ebar[5], 0);
        }
        produce(k, u + 3 >= 5 ? u - 2 : u + 3, u);   // P1(k+4) takes the slot of P1(k-1)
        if (++u == 5) { u = 0; epar ^= 1u; }
      }
    }
    return;
  }
  // elements h0 .. h0+TW-1 of this tile: ih = 0 holds i = 1, i = m sits at ih = (m-1)/2 or m/2 - 1
  if (h0 <= 0 || 2 * (h0 + TW - 1) + 2 >= m) sweep(std::true_type{});
  else                                       sweep(std::false_type{});
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if ((tid & 31) == 0) wmax[tid >> 5] = emax;
  named_bar(1, NCOMPUTE);
  if (tid < 32) {
    double v = (tid < NCOMPUTE / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
  if (A.sync && tid == 0) {
    // every compute thread of this block is past its last store (the barrier above); make them visible everywhere,
    // then count the block.  The last block of a boundary chunk publishes the launch number into the neighbour; the
    // last block of the launch resets the counters and advances the launch number for the next launch.
    const unsigned long long n = A.sync[SY_LAUNCH] + 1;
    const unsigned long long per_chunk = (unsigned long long)A.ntiles;   // every tile has one bottom and one top chunk
    __threadfence_system();
    if (slab_bot && atomicAdd(A.sync + SY_CNT_BOT, 1ull) + 1 == per_chunk) {
      __threadfence_system();
      st_release_sys(A.to_prev, n);
    }
    if (slab_top && atomicAdd(A.sync + SY_CNT_TOP, 1ull) + 1 == per_chunk) {
      __threadfence_system();
      st_release_sys(A.to_next, n);
    }
    if (atomicAdd(A.sync + SY_CNT_ALL, 1ull) + 1 == (unsigned long long)gridDim.x) {
      __threadfence();
      A.sync[SY_CNT_BOT] = 0;
      A.sync[SY_CNT_TOP] = 0;
      A.sync[SY_CNT_ALL] = 0;
      A.sync[SY_LAUNCH] = n;
    }
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
  a.pout0 = A.p[in ^ 1][0];
  a.pout1 = A.p[in ^ 1][1];
  a.ilo0 = A.img_lo[in ^ 1][0]; a.ilo1 = A.img_lo[in ^ 1][1];
  a.ihi0 = A.img_hi[in ^ 1][0]; a.ihi1 = A.img_hi[in ^ 1][1];
  a.dk_lo = (int)A.dk_lo;
  a.dk_hi = (int)A.dk_hi;
  a.sync = A.sync;
  a.to_prev = A.sync_to_prev;
  a.to_next = A.sync_to_next;
  const int cols = ((g.m + 1) >> 1) + 2;            // elements -1 .. ihmax
  a.xt = (cols + (TW - 2) - 1) / (TW - 2);
  a.ntiles = a.xt * ((g.n + (TR - 2) - 1) / (TR - 2));
  a.tA = A.tma_tA; a.nzA = A.tma_nzA; a.nzB = A.tma_nzB;
  const int blocks = a.tA * a.nzA + (a.ntiles - a.tA) * a.nzB;
  sor_tma_kernel<<<blocks, NTHREADS, SMEM_BYTES, st>>>(M, g, a, ph.relux, 1. - ph.relux, err_bits);
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

void pf_tma_release(FusedArrays &A) {
  delete[] static_cast<TmaMaps *>(A.tma_cache);
  A.tma_cache = nullptr;
}

// z-chunks of the TMA kernel (one block per SM).  A block costs (planes of its chunk + 2) z-steps plus about one step
// of start-up, and blocks are handed to SMs in launch order as SMs become free.  Cutting every tile's column the same
// way leaves a ragged last wave -- 666 tiles of 64 planes on 148 SMs: whole columns are 4.5 waves (5 x 66 steps), halves
// 9 waves of 34 (306) -- so the first tA tiles (the full waves) may be cut differently from the rest: 592 whole
// columns and 74 halved ones are 4 x 66 + 34 = 298 steps.  The schedule with the shortest simulated makespan wins.
// (Measured gains are larger than the step count predicts -- 1024x512x512: 2.77 vs 2.89 ms per iteration where the
// model sees 0.1 % -- a block's real start-up, pipeline fill plus the generic head and tail steps, is worth more than
// the three steps charged here; the choices do not change with a larger constant.)
void pf_tma_schedule(const Geo &g, FusedArrays &A) {
  const int cols = ((g.m + 1) >> 1) + 2;
  const int xt = (cols + (TW - 2) - 1) / (TW - 2);
  const int T = xt * ((g.n + (TR - 2) - 1) / (TR - 2));
  const int S = pf_sm_count() * PF_TMA_MINB;
  auto norm = [&](int nz) {   // chunks of ceil(lz/nz) planes: the number of non-empty ones
    const int cz = (g.lz + nz - 1) / nz;
    return (g.lz + cz - 1) / cz;
  };
  if (const char *e = getenv("PF_TMA_CHUNK")) {   // tuning experiments only: every tile in chunks of this many planes
    const int v = atoi(e);
    if (v >= 1) {
      A.tma_tA = T;
      A.tma_nzA = A.tma_nzB = norm((g.lz + v - 1) / v);
      return;
    }
  }
  auto makespan = [&](int tA, int nzA, int nzB) {
    std::vector<double> heap(S, 0.0);   // min-heap of the times the SMs become free
    auto cmp = [](double x, double y) { return x > y; };
    auto run = [&](long long nblocks, double cost) {
      for (long long b = 0; b < nblocks; ++b) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        heap.back() += cost;
        std::push_heap(heap.begin(), heap.end(), cmp);
      }
    };
    run((long long)tA * nzA, (g.lz + nzA - 1) / nzA + 3.0);
    run((long long)(T - tA) * nzB, (g.lz + nzB - 1) / nzB + 3.0);
    return *std::max_element(heap.begin(), heap.end());
  };
  double best = 1e300;
  const int nzmax = std::max(1, std::min(16, g.lz / 8));
  for (int nzA = 1; nzA <= nzmax; ++nzA) {
    if (norm(nzA) != nzA) continue;
    for (int nzB = nzA; nzB <= nzmax; ++nzB) {
      if (norm(nzB) != nzB) continue;
      // candidates for tA: everything in group A, or as many tiles as fill whole waves of group-A blocks
      int cands[2] = {T, (int)((long long)T * nzA / S * S / nzA)};
      for (int tA : cands) {
        if (tA < 0 || tA > T || (nzB == nzA && tA != T)) continue;
        if (tA == T && nzB != nzA) continue;
        const double c = makespan(tA, nzA, nzB);
        if (c < best - 1e-9) { best = c; A.tma_tA = tA; A.tma_nzA = nzA; A.tma_nzB = nzB; }
      }
    }
  }
}
