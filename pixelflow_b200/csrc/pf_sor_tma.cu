// pf_sor_tma.cu -- SOR variant 6: the fused red+black pass of pf_sor_fused.cu with its operands staged
// by TMA (cp.async.bulk.tensor, 3-D tensor maps over the depth-2-ghost checkerboard arrays) into a
// multi-stage shared-memory pipeline guarded by mbarriers.
//
// Why: the register-prefetch version (variant 3) moves the right bytes (48 B/cell/sweep) but is bound by
// latency and instruction issue -- every z-step is a short dependent chain of scalar 8-byte loads.  Here
// one elected thread issues ten box copies per plane two steps ahead of their use; the 256 compute
// threads only touch shared memory (16-byte accesses, two cells of each colour per thread and step).
//
// Tile: 64 columns x 8 rows of checkerboard elements (ring included: owned columns 1..62, owned rows
// 1..6), streamed along a z-chunk.  Per plane p the pipeline holds
//   group(p) : P0 BB0 CX0 CX1 CY0 CY1 CZ0 CZ1 BB1          (4 slots; used by red(p) and black(p))
//   P1(p)    : the old black pressure, box widened by the ring   (5 slots; used by red(p-1), red(p), red(p+1),
//              black(p))
// and a 4-slot ring R of the new red values.  Arithmetic is the same sor_update() as everywhere else.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "pf_internal.cuh"

namespace {

constexpr int TW = 64;        // tile columns (elements), ring columns 0 and 63
constexpr int TWP = 68;       // widened boxes: columns -2 .. 65
constexpr int TR = 8;         // tile rows, ring rows 0 and 7
constexpr int NG = 4;         // group slots
constexpr int NP = 5;         // P1 slots
constexpr int NR = 4;         // R slots
constexpr int NTHREADS = 256;

// byte sizes of the staged boxes (all multiples of 128)
constexpr int SZ_N = TW * TR * 8;             // narrow box            4096
constexpr int SZ_W = TWP * TR * 8;            // wide box              4352
constexpr int SZ_CY1 = TW * (TR + 1) * 8;     // rows -1 .. TR-1       4608
constexpr int SZ_P1 = 5504;                   // 68 x 10 x 8 = 5440, padded to a multiple of 128
constexpr int P1_BYTES = TWP * (TR + 2) * 8;  // bytes actually copied 5440
// group layout: P0, BB0, CY0, CZ0, CZ1, BB1 (narrow) ; CX0, CX1 (wide) ; CY1
constexpr int OFF_P0 = 0, OFF_BB0 = SZ_N, OFF_CY0 = 2 * SZ_N, OFF_CZ0 = 3 * SZ_N, OFF_CZ1 = 4 * SZ_N,
              OFF_BB1 = 5 * SZ_N, OFF_CX0 = 6 * SZ_N, OFF_CX1 = 6 * SZ_N + SZ_W, OFF_CY1 = 6 * SZ_N + 2 * SZ_W;
constexpr int SZ_GROUP = 6 * SZ_N + 2 * SZ_W + SZ_CY1;   // 37888
constexpr int GROUP_BYTES = SZ_GROUP;
constexpr int SZ_R = TWP * TR * 8;                        // 4352 per slot
constexpr int SMEM_BYTES = NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R + 256 + 128;   // + mbarriers, block-max scratch, alignment slack

struct TmaMaps {
  CUtensorMap p0, p1, cx0, cx1, cy0, cy1, cz0, cz1, bb0, bb1;
};

struct TmaArgs {
  int NY2, hplane2, cz_planes;
  double *pout0, *pout1;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spins > (1u << 24)) __trap();   // never hang the GPU: a lost copy aborts the kernel instead
  }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ double sor_update(double bb, double ae, double aw, double an, double as, double at,
                                             double ab, double pE, double pW, double pN, double pS, double pT,
                                             double pB, double pold, double relux, double omr, int i, int m) {
  const double ap = -ae - aw - an - as - at - ab;   // ibm_3d_uniform_omp_cpu.f90:402, raw coefficients
  if (i == 1 || i == m) {
    if (i == 1) { ae = ae + aw; aw = 0.; }            // :640-641
    if (i == m) { ae = aw = an = as = at = ab = 0.; } // :651-656
  }
  const double r = bb - ae * pE - aw * pW - an * pN - as * pS - at * pT - ab * pB;   // :510-515
  return r / ap * relux + pold * omr;
}

__device__ __forceinline__ void store_with_images(double *dst, int c, int dj, int dk, double v) {
  dst[c] = v;
  if (dj) dst[c + dj] = v;
  if (dk) {
    dst[c + dk] = v;
    if (dj) dst[c + dk + dj] = v;
  }
}

__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }

__global__ void __launch_bounds__(NTHREADS, 1) sor_tma_kernel(const __grid_constant__ TmaMaps M, Geo g, TmaArgs A,
                                                              double relux, unsigned long long *err_bits) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  unsigned char *grp = smem;                                  // NG x SZ_GROUP
  unsigned char *p1s = smem + NG * SZ_GROUP;                  // NP x SZ_P1
  double *Rs = reinterpret_cast<double *>(smem + NG * SZ_GROUP + NP * SZ_P1);   // NR x TR x TWP
  uint64_t *gbar = reinterpret_cast<uint64_t *>(smem + NG * SZ_GROUP + NP * SZ_P1 + NR * SZ_R);   // NG
  uint64_t *pbar = gbar + NG;                                                                     // NP

  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int h0 = (int)blockIdx.x * (TW - 2) - 2;     // element index of tile column 0 (stride 62, even)
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

  auto issue_group = [&](int p) {   // plane p (local index), called by thread 0
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
  auto wait_group = [&](int p) { const int q = p - kfirst; mbar_wait(&gbar[q % NG], (q / NG) & 1); };
  auto wait_p1 = [&](int p) { const int q = p - (kfirst - 1); mbar_wait(&pbar[q % NP], (q / NP) & 1); };
  auto G = [&](int p, int off) { return reinterpret_cast<const double *>(grp + ((p - kfirst) % NG) * SZ_GROUP + off); };
  auto P1 = [&](int p) { return reinterpret_cast<const double *>(p1s + ((p - (kfirst - 1)) % NP) * SZ_P1); };
  auto R = [&](int p) { return Rs + ((p - kfirst) % NR) * (TR * TWP); };

  if (tid == 0) {   // prologue: group planes kfirst, kfirst+1 ; P1 planes kfirst-1 .. kfirst+2
    issue_group(kfirst);
    if (kfirst + 1 <= klast) issue_group(kfirst + 1);
    for (int p = kfirst - 1; p <= kfirst + 2; ++p) issue_p1(p);
  }
  wait_p1(kfirst - 1);
  wait_p1(kfirst);

  const int e0 = 2 * tx;                              // tile columns of this thread: e0, e0+1
  const int ih0 = h0 + e0;
  const int j = j0 + ty;
  const int m = g.m, ihmax = (g.m + 1) >> 1;
  const bool row_in = j <= g.n + 1;
  const bool own_row = ty >= 1 && ty <= TR - 2 && j >= 1 && j <= g.n;
  const double omr = 1. - relux;
  const int dj = (j <= 2) ? g.n * g.HX : ((j >= g.n - 1) ? -g.n * g.HX : 0);
  const int sj = (j + g.koff) & 1;
  const int cbase = g.H0 + ih0 + g.HX * (j + 1);      // + hplane2*(k+1) = global element index
  double cza0 = 0., cza1 = 0., czb0 = 0., czb1 = 0.;  // cz_red of this thread's two columns at planes k-2 (a), k-1 (b)
  double emax = 0.;

  for (int k = kfirst; k <= klast; ++k) {
    if (tid == 0) {
      if (k + 2 <= klast) issue_group(k + 2);
      if (k + 3 <= klast + 1) issue_p1(k + 3);
    }
    wait_group(k);
    wait_p1(k + 1);
    const int s = (sj + k) & 1;       // parity of i: red row at plane k, black row at plane k-1
    const int i0 = 2 * ih0 + 2 - s, i1 = i0 + 2;
    const bool cell0 = row_in && ih0 >= -1 && ih0 <= ihmax && i0 >= 1 && i0 <= m;
    const bool cell1 = row_in && ih0 + 1 <= ihmax && i1 >= 1 && i1 <= m;
    // ------------------------------------------ red stage, plane k
    const double *p1a = P1(k - 1), *p1b = P1(k), *p1c = P1(k + 1);
    const int wr = (ty + 1) * TWP + e0 + 2;           // wide-box index of (row ty, column e0)
    const int nr = ty * TW + e0;                      // narrow-box index
    const double2 pold = lds2(G(k, OFF_P0) + nr);
    const double2 at = lds2(G(k, OFF_CZ0) + nr);
    double2 val = pold;
    {
      const double2 bb = lds2(G(k, OFF_BB0) + nr), an = lds2(G(k, OFF_CY0) + nr);
      const double2 as = lds2(G(k, OFF_CY1) + ty * TW + e0);          // row j-1 (box starts at row j0-1)
      const double2 ab = lds2(G(k, OFF_CZ1) + nr);                    // cz1(k-1)
      const double2 ae = lds2(G(k, OFF_CX0) + ty * TWP + e0 + 2);
      const double *cxw = G(k, OFF_CX1) + ty * TWP + e0 + 2;
      const double2 pc = lds2(p1b + wr);                              // black at (e0, e0+1) of this row
      const double pl = p1b[wr - 1], pr = p1b[wr + 2];
      const double2 pN = lds2(p1b + wr + TWP), pS = lds2(p1b + wr - TWP);
      const double2 pT = lds2(p1c + wr), pB = lds2(p1a + wr);
      // west/east neighbours: s=1 -> {e-1, e}, s=0 -> {e, e+1}
      const double pW0 = s ? pl : pc.x, pE0 = s ? pc.x : pc.y;
      const double pW1 = s ? pc.x : pc.y, pE1 = s ? pc.y : pr;
      const double aw0 = cxw[-s], aw1 = cxw[1 - s];
      if (cell0) val.x = sor_update(bb.x, ae.x, aw0, an.x, as.x, at.x, ab.x, pE0, pW0, pN.x, pS.x, pT.x, pB.x, pold.x,
                                    relux, omr, i0, m);
      if (cell1) val.y = sor_update(bb.y, ae.y, aw1, an.y, as.y, at.y, ab.y, pE1, pW1, pN.y, pS.y, pT.y, pB.y, pold.y,
                                    relux, omr, i1, m);
    }
    *reinterpret_cast<double2 *>(R(k) + ty * TWP + e0 + 2) = val;
    if (own_row && k >= kc0 && k <= kc1) {
      const int dk = (k <= 2) ? g.lz * A.hplane2 : ((k >= g.lz - 1) ? -g.lz * A.hplane2 : 0);
      const int c = cbase + A.hplane2 * (k + 1);
      if (cell0 && e0 >= 1) store_with_images(A.pout0, c, dj, dk, val.x);
      if (cell1 && e0 + 1 <= TW - 2) store_with_images(A.pout0, c + 1, dj, dk, val.y);
    }
    __syncthreads();
    // ------------------------------------------ black stage, plane k-1
    const int kb = k - 1;
    if (kb >= kc0 && kb <= kc1 && own_row) {
      const double2 pold = lds2(p1a + wr);                            // black own old value (plane k-1)
      const double2 bb = lds2(G(kb, OFF_BB1) + nr);
      const double2 ae = lds2(G(kb, OFF_CX1) + ty * TWP + e0 + 2);
      const double *cxw = G(kb, OFF_CX0) + ty * TWP + e0 + 2;
      const double2 an = lds2(G(kb, OFF_CY1) + (ty + 1) * TW + e0);   // own row j (box starts at row j0-1)
      const double2 as = lds2(G(kb, OFF_CY0) + (ty - 1) * TW + e0);
      const double2 atb = lds2(G(k, OFF_CZ1) + nr);                   // cz1(k-1) == cz1(kb)
      const double *rb = R(kb) + ty * TWP + e0 + 2;
      const double2 rc = lds2(rb);
      const double rl = rb[-1], rr = rb[2];
      const double2 rN = lds2(rb + TWP), rS = lds2(rb - TWP);
      const double2 rB = (kb - 1 >= kfirst) ? lds2(R(kb - 1) + ty * TWP + e0 + 2) : make_double2(0., 0.);
      const double pW0 = s ? rl : rc.x, pE0 = s ? rc.x : rc.y;
      const double pW1 = s ? rc.x : rc.y, pE1 = s ? rc.y : rr;
      const double aw0 = cxw[-s], aw1 = cxw[1 - s];
      const int dk = (kb <= 2) ? g.lz * A.hplane2 : ((kb >= g.lz - 1) ? -g.lz * A.hplane2 : 0);
      const int c = cbase + A.hplane2 * (kb + 1);
      if (cell0 && e0 >= 1) {
        const double v = sor_update(bb.x, ae.x, aw0, an.x, as.x, atb.x, cza0, pE0, pW0, rN.x, rS.x, val.x, rB.x, pold.x,
                                    relux, omr, i0, m);
        store_with_images(A.pout1, c, dj, dk, v);
        emax = fmax(emax, fabs(v - pold.x));
      }
      if (cell1 && e0 + 1 <= TW - 2) {
        const double v = sor_update(bb.y, ae.y, aw1, an.y, as.y, atb.y, cza1, pE1, pW1, rN.y, rS.y, val.y, rB.y, pold.y,
                                    relux, omr, i1, m);
        store_with_images(A.pout1, c + 1, dj, dk, v);
        emax = fmax(emax, fabs(v - pold.y));
      }
    }
    cza0 = czb0; cza1 = czb1; czb0 = at.x; czb1 = at.y;
    // order this step's shared-memory reads before the async-proxy writes of the next step's copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  double *wmax = reinterpret_cast<double *>(pbar + NP + 1);
  if ((tid & 31) == 0) wmax[tid >> 5] = emax;
  __syncthreads();
  if (tid < 32) {
    double v = (tid < NTHREADS / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PF_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::string("cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

CUtensorMap make_map(const Geo &g, const double *base, int box_cols, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)g.HX, (cuuint64_t)(g.n + 4), (cuuint64_t)(g.lz + 5)};
  const cuuint64_t strides[2] = {(cuuint64_t)g.HX * 8, (cuuint64_t)g.HX * (g.n + 4) * 8};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::string("cuTensorMapEncodeTiled failed (") + std::to_string((int)r) + ")";
  return m;
}

}  // namespace

bool pf_tma_applicable(const Geo &g, const Phys &ph, int nranks) {
  return pf_fused_applicable(g, ph, nranks) && (g.HX * 8) % 16 == 0;
}

// one red-black iteration through the TMA pipeline: reads A.p[in], writes A.p[in^1]
void k_tma_iteration(const Geo &g, const Phys &ph, FusedArrays &A, int in, unsigned long long *err_bits,
                     cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    PF_CUDA_OK(cudaFuncSetAttribute(sor_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  if (!A.tma_cache) {   // the ten tensor maps of each ping-pong direction, encoded once
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
  const int cols = ((g.m + 1) >> 1) + 2;            // elements -1 .. ihmax
  const int xt = (cols + (TW - 2) - 1) / (TW - 2);
  const int yt = (g.n + (TR - 2) - 1) / (TR - 2);
  const int zt = (g.lz + A.cz_planes - 1) / A.cz_planes;
  sor_tma_kernel<<<dim3(xt, yt, zt), NTHREADS, SMEM_BYTES, st>>>(M, g, a, ph.relux, err_bits);
  pf_count_launch();
}

// z-chunk size for the TMA kernel (1 block per SM): whole waves, chunks >= 16 planes
int pf_tma_chunk(const Geo &g) {
  const int cols = ((g.m + 1) >> 1) + 2;
  const int xt = (cols + (TW - 2) - 1) / (TW - 2);
  const int yt = (g.n + (TR - 2) - 1) / (TR - 2);
  int best = g.lz;
  double best_cost = 1e30;
  for (int cz = g.lz; cz >= 16; --cz) {
    const long long blocks = (long long)xt * yt * ((g.lz + cz - 1) / cz);
    const long long waves = (blocks + 147) / 148;
    const double cost = (double)waves * (cz + 2 + 3);
    if (cost < best_cost) { best_cost = cost; best = cz; }
  }
  return best;
}
