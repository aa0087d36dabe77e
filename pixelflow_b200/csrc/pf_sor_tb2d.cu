// pf_sor_tb2d.cu -- SOR variant 8: TEMPORALLY BLOCKED red-black iterations for the 2D cases.
//
// Reference: solve_matrix_vec_omp, src/omp_parallel/ibm_2d_uniform_omp_cpu.f90:293-406 (the backstep and drag programs
// carry the same routine).  Per iteration: {halo refresh; p_old = p; sweep of the cells with (i+j) odd; halo refresh;
// p_old = p; sweep of the cells with (i+j) even}, the running maximum of |p - p_old| taken in BOTH sweeps (:351, :385).
//
// Why: the 2D grids of the reference's decks (cylinder 1024x512, backstep 2251x411) live in L2, and a solve is 200
// dependent half-sweeps -- 200 launches of 3-4 us each (variant 1), or 200 grid-wide barriers of about the same price
// (variant 7): latency, not bandwidth.  A cell of iteration t+T depends only on the cells within 2T of it, so a block
// that loads its tile with a ring 2T cells deep can run T whole iterations in shared memory without talking to anybody:
// one launch per T iterations instead of 2T, the ring recomputed redundantly by the neighbouring blocks (same
// operations on the same values, hence the same bits).
//
//   * operands: exactly the arrays variant 1 reads -- the folded coefficients ae aw an as, ap, bb and p in checkerboard
//     layout -- copied into shared memory in natural layout (7 doubles per cell);
//   * the periodic y direction is a wrap of the row index when the tile is loaded (the halo rows of the reference
//     hold, at every half-sweep, the current value of the row they image: :323-330); p(0,j) and p(m+1,j) are never
//     written inside the solver (:433-614 touches i = 1..m only) and ride along as constants;
//   * a half-sweep is two phases -- every thread computes its new values from the tile, barrier, then stores them --
//     which is the reference's p_old copy in miniature, and is what makes an odd n safe: across the periodic seam two
//     cells of the SAME colour are neighbours, and each must read the other's value from before the half-sweep;
//   * cells on the outermost ring of the tile are never updated; after h half-sweeps the cells closer than h to that
//     ring are stale, the owned cells (2T away) never are;
//   * update (:339-352), evaluated left to right, no FMA:
//       p = (bb - ae*pE - aw*pW - an*pN - as*pS) / ap * relux + p*(1 - relux)
//   * launches ping-pong between the solver's checkerboard p arrays and a second pair; an odd number of launches ends
//     with one device copy back.
#include <algorithm>

#include "pf_internal.cuh"

namespace {

constexpr int TB_THREADS = 512;
constexpr int TB_MAXR = 8;          // updates per thread and half-sweep held in registers between the two phases
constexpr int TB_ARRAYS = 7;

struct TbArgs {
  SplitSet in[2];                   // per colour: coefficients, bb and the pressure this launch reads
  double *pout[2];                  // per colour: the pressure this launch writes (owned cells + x-halo columns)
  int ow, oh;                       // owned cells per tile
  int T;                            // iterations of this launch
  int pitch;                        // tile row pitch in doubles (ow + 4T rounded up to odd: no bank conflicts by row)
  double relux;
};

__device__ __forceinline__ int ih_of(int i) { return ((i + 1) >> 1) - 1; }

__global__ void __launch_bounds__(TB_THREADS) sor_tb2d_kernel(Geo g, TbArgs A, unsigned long long *err_bits) {
  extern __shared__ double sm[];
  const int D = 2 * A.T;
  const int i0 = (int)blockIdx.x * A.ow + 1, j0 = (int)blockIdx.y * A.oh + 1;   // first owned cell
  const int i1 = min(i0 + A.ow - 1, g.m), j1 = min(j0 + A.oh - 1, g.n);         // last owned cell
  const int xlo = max(i0 - D, 0), xhi = min(i1 + D, g.m + 1);                   // tile columns: real i
  const int ew = xhi - xlo + 1, eh = (j1 - j0 + 1) + 2 * D;                     // tile size
  const int pitch = A.pitch;
  const int cells = eh * pitch;
  double *P = sm, *AE = sm + cells, *AW = sm + 2 * cells, *AN = sm + 3 * cells, *AS = sm + 4 * cells,
         *AP = sm + 5 * cells, *BB = sm + 6 * cells;
  const int tid = threadIdx.x;
  const int n = g.n;
  // tile row r holds the virtual row j0 - D + r, i.e. the real row wrapped into 1..n
  auto real_row = [&](int r) {
    int jv = (j0 - D + r - 1) % n;
    if (jv < 0) jv += n;
    return jv + 1;
  };
  // ---- load
  for (int t = tid; t < eh * ew; t += TB_THREADS) {
    const int r = t / ew, x = t - r * ew;
    const int i = xlo + x, jr = real_row(r);
    const int c = (i + jr) & 1;
    const long long h = split_row(g, jr, 0) + ih_of(i);
    const int o = r * pitch + x;
    const SplitSet &S = A.in[c];
    P[o] = S.p[h];
    if (i >= 1 && i <= g.m) {
      AE[o] = S.ae[h]; AW[o] = S.aw[h]; AN[o] = S.an[h]; AS[o] = S.as[h]; AP[o] = S.ap[h]; BB[o] = S.bb[h];
    }
  }
  __syncthreads();
  // ---- 2T half-sweeps in shared memory
  const double relux = A.relux, omr = 1. - relux;
  double emax = 0.0;
  const int hw = (ew + 1) >> 1;                       // candidate columns of one colour in a tile row
  const int per_sweep = eh * hw;
  // the candidate cells of this thread: tile row, first column of the pair, row parity, row inside the ring / owned
  int q_o[TB_MAXR], q_fl[TB_MAXR];
#pragma unroll
  for (int q = 0; q < TB_MAXR; ++q) {
    const int t = tid + q * TB_THREADS;
    q_o[q] = 0;
    q_fl[q] = 0;
    if (t < per_sweep) {
      const int r = t / hw, xq = t - r * hw;
      const int jv = j0 - D + r;
      q_o[q] = r * pitch + 2 * xq;
      q_fl[q] = ((xlo + real_row(r)) & 1) | ((r > 0 && r < eh - 1) ? 2 : 0) | ((jv >= j0 && jv <= j1) ? 4 : 0) | (2 * xq << 3);
    }
  }
  for (int hs = 0; hs < 2 * A.T; ++hs) {
    const int colour = (hs & 1) ^ 1;                  // (i+j) odd first (:339-344)
    double nv[TB_MAXR];
    int no[TB_MAXR];
#pragma unroll
    for (int q = 0; q < TB_MAXR; ++q) {
      no[q] = -1;
      const int fl = q_fl[q];
      const int off = (colour + fl) & 1;              // (xlo + x + jr) & 1 == colour
      const int x = (fl >> 3) + off;
      const int i = xlo + x;
      if ((fl & 2) && x > 0 && x < ew - 1 && i >= 1 && i <= g.m) {
        const int o = q_o[q] + off;
        const double pc = P[o];
        const double ra = BB[o] - AE[o] * P[o + 1] - AW[o] * P[o - 1] - AN[o] * P[o + pitch] - AS[o] * P[o - pitch];
        const double out = ra / AP[o] * relux + pc * omr;
        nv[q] = out;
        no[q] = o;
        if ((fl & 4) && i >= i0 && i <= i1) emax = fmax(emax, fabs(out - pc));
      }
    }
    __syncthreads();                                  // every read of this half-sweep is done
#pragma unroll
    for (int q = 0; q < TB_MAXR; ++q)
      if (no[q] >= 0) P[no[q]] = nv[q];
    __syncthreads();
  }
  // ---- store the owned cells (and the x-halo columns next to them, so the output arrays are complete)
  const int own_w = i1 - i0 + 1, own_h = j1 - j0 + 1;
  const int sx0 = (i0 == 1) ? 0 : i0, sx1 = (i1 == g.m) ? g.m + 1 : i1;
  const int sw = sx1 - sx0 + 1;
  (void)own_w;
  for (int t = tid; t < own_h * sw; t += TB_THREADS) {
    const int rr = t / sw, i = sx0 + (t - rr * sw);
    const int j = j0 + rr;
    const int o = (D + rr) * pitch + (i - xlo);
    A.pout[(i + j) & 1][split_row(g, j, 0) + ih_of(i)] = P[o];
  }
  // ---- block maximum of |p - p_old| over the owned cells (non-negative doubles order like their bit patterns)
  for (int s = 16; s > 0; s >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, s));
  __shared__ double wmax[TB_THREADS / 32];
  if ((tid & 31) == 0) wmax[tid >> 5] = emax;
  __syncthreads();
  if (tid < 32) {
    double v = (tid < TB_THREADS / 32) ? wmax[tid] : 0.0;
    for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e && atoi(e) > 0 ? atoi(e) : dflt;
}

}  // namespace

// one GPU, 2D, and rows long enough to wrap a tile around (the wrap of real_row() is general, the check is for sanity)
bool pf_tb2d_applicable(const Geo &g, int nranks) { return nranks == 1 && g.dim == 2 && g.n >= 2 && g.m >= 2; }

// iterations per launch and owned-tile shape.  Shared memory holds 7 doubles per tile cell, a thread keeps at most
// TB_MAXR updates of a half-sweep in registers.  Defaults: T = 4 (ring of 8), owned 64 x 32 -> tile 80 x 48 cells,
// 215 KB; PF_TB_T / PF_TB_OW / PF_TB_OH override them for tuning runs.
void pf_tb2d_shape(const Geo &g, int iters_left, int &T, int &ow, int &oh) {
  T = std::min(env_int("PF_TB_T", 4), std::max(iters_left, 1));
  ow = std::min(env_int("PF_TB_OW", 64), g.m);
  oh = std::min(env_int("PF_TB_OH", 32), g.n);
  const long long budget = 227 * 1024 - 1024;
  auto fits = [&](int t, int w, int h) {
    const long long ew = w + 4 * t, eh = h + 4 * t, pitch = ew | 1;
    return eh * pitch * 8 * TB_ARRAYS <= budget && eh * ((ew + 1) / 2) <= (long long)TB_MAXR * TB_THREADS;
  };
  while (!fits(T, ow, oh)) {
    if (oh > 8) oh -= 4;
    else if (ow > 16) ow -= 8;
    else if (T > 1) --T;
    else break;
  }
}

// `iters` red-black iterations; reads S[c].p, leaves the result in S[c].p; alt[c] = a second pair of checkerboard arrays
void k_sor_tb2d(const Geo &g, const Phys &ph, const SplitSet S[2], double *const alt[2], int iters,
                unsigned long long *err_bits, cudaStream_t st) {
  if (iters <= 0) return;
  static bool attr_set[64] = {};
  int dev = 0;
  PF_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PF_CUDA_OK(cudaFuncSetAttribute(sor_tb2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    attr_set[dev] = true;
  }
  double *cur[2] = {S[0].p, S[1].p}, *oth[2] = {alt[0], alt[1]};
  int left = iters;
  while (left > 0) {
    TbArgs A;
    A.in[0] = S[0]; A.in[1] = S[1];
    A.in[0].p = cur[0]; A.in[1].p = cur[1];
    A.pout[0] = oth[0]; A.pout[1] = oth[1];
    pf_tb2d_shape(g, left, A.T, A.ow, A.oh);
    A.pitch = (A.ow + 4 * A.T) | 1;
    A.relux = ph.relux;
    const int eh = A.oh + 4 * A.T;
    const size_t smem = (size_t)eh * A.pitch * 8 * TB_ARRAYS;
    const dim3 grid((g.m + A.ow - 1) / A.ow, (g.n + A.oh - 1) / A.oh);
    sor_tb2d_kernel<<<grid, TB_THREADS, smem, st>>>(g, A, err_bits);
    pf_count_launch();
    PF_CUDA_OK(cudaGetLastError());
    std::swap(cur[0], oth[0]);
    std::swap(cur[1], oth[1]);
    left -= A.T;
  }
  if (cur[0] != S[0].p) {   // an odd number of launches: the result sits in the second pair
    PF_CUDA_OK(cudaMemcpyAsync(S[0].p, cur[0], (size_t)g.split_elems * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PF_CUDA_OK(cudaMemcpyAsync(S[1].p, cur[1], (size_t)g.split_elems * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
}
