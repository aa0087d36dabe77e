// pf_sor_persistent.cu -- SOR variant 7: the colour half-sweeps of pf_sor.cu with the iteration loop ON THE DEVICE --
// one launch per solve (cooperative, so that every block is resident), one block of 1024 threads per SM, and a
// hand-rolled grid barrier between half-sweeps (one atomic per block on an L2 word + an acquire spin: about a third
// of cooperative_groups' grid.sync(), which round 2 measured at ~4.6 us on 512 blocks -- slower than the launch
// gaps it was meant to remove).
//
// Why: the reference's own decks (cylinder 1024x512, backstep 2251x411, room 64^3) are L2-resident; their solve is
// 200 launches of sor_sweep_kernel of 4-7 us each, about twice the L2-bandwidth time of a half-sweep
// (profiles/r01_v6_decks_bench.jsonl) -- launch-bound even when replayed from a CUDA graph.  Where a solve is
// exactly `iters x 2` half-sweeps with nothing in between (all 2D cases: the sweep keeps its own periodic y-halo
// rows, pf_sor.cu YIMG; 3D air-condition on one GPU: no halo refresh inside the solve,
// ibm_3d_air_condition_omp_cpu.f90:509-527) the loop can live in one kernel.
//
// Same arithmetic, same order as sor_sweep_kernel (reference update :510-515, left to right, no FMA).  Differences:
//   * tiles of the half-sweep's grid are distributed over the resident blocks with a grid-stride loop;
//   * the other colour's pressure is written by other blocks one half-sweep earlier, so every pressure load goes to
//     L2 (ld.global.cg: L1 is not coherent between SMs) and the grid barrier orders the half-sweeps; coefficients and
//     bb are read-only for the whole solve and keep the streaming loads;
//   * the error (running max over all iterations: :575-583 second colour only in 3D, both colours in 2D :351,:385)
//     is accumulated per thread for the whole solve and reduced once at the end.
#include "pf_internal.cuh"

namespace {

constexpr int PBX = 64, PBY = 16;   // 64 pair-threads x 16 rows: one block of 1024 threads per SM

__device__ __forceinline__ double2 ld2c(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ double ld1c(const double *p) { return __ldcg(p); }

// every block has arrived `target` times in total: thread 0 publishes the block's stores (fence), counts it in and
// spins on the L2 word; the block waits for it at the closing __syncthreads()
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
__device__ __forceinline__ double2 ld2s(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

struct PersistArgs {
  SplitSet S[2];
  int iters;
  int first_colour;   // 3D: 0 ((i+j+k) even first, :492-517); 2D: 1 ((i+j) odd first, ibm_2d_uniform_omp_cpu.f90:339-352)
  int gx, gy, gz;     // tiles of one half-sweep
  double relux;
  unsigned int *bar;  // grid-barrier word, zero at launch
};

template <int DIM, int YIMG>
__global__ void __launch_bounds__(PBX *PBY) sor_persistent_kernel(Geo g, PersistArgs A, unsigned long long *err_bits) {
  unsigned int arrivals = 0;
  const long long ntiles = (long long)A.gx * A.gy * A.gz;
  const double relux = A.relux, omr = 1. - relux;
  double emax = 0.0;
  for (int it = 0; it < A.iters; ++it) {
    for (int half = 0; half < 2; ++half) {
      const int colour = half ? (A.first_colour ^ 1) : A.first_colour;
      const bool with_err = (DIM == 2) || half == 1;
      // pointer sets selected with ternaries (a run-time index into the parameter struct would spill it to local memory)
      SplitSet S;
      S.ap = colour ? A.S[1].ap : A.S[0].ap; S.bb = colour ? A.S[1].bb : A.S[0].bb;
      S.ae = colour ? A.S[1].ae : A.S[0].ae; S.aw = colour ? A.S[1].aw : A.S[0].aw;
      S.an = colour ? A.S[1].an : A.S[0].an; S.as = colour ? A.S[1].as : A.S[0].as;
      S.at = colour ? A.S[1].at : A.S[0].at; S.ab = colour ? A.S[1].ab : A.S[0].ab;
      S.p = colour ? A.S[1].p : A.S[0].p;
      double *po = colour ? A.S[0].p : A.S[1].p;   // written by the previous half-sweep: coherent loads only
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int bx = (int)(t % A.gx);
        const int by = (int)((t / A.gx) % A.gy);
        const int bz = (int)(t / ((long long)A.gx * A.gy));
        const int q = bx * PBX + threadIdx.x;
        const int j = by * PBY + threadIdx.y + 1;
        const int k = (DIM == 3) ? bz + 1 : 0;
        if (j > g.n) continue;
        const int s = (colour + j + k + g.koff) & 1;        // parity of i in this row of this colour
        const int cnt = s ? (g.m + 1) >> 1 : g.m >> 1;      // interior cells of that parity
        const int ih = 2 * q;
        if (ih < cnt) {
          const long long r = split_row(g, j, k) + ih;
          const double2 ap = ld2s(S.ap + r), bb = ld2s(S.bb + r);
          const double2 ae = ld2s(S.ae + r), aw = ld2s(S.aw + r);
          const double2 an = ld2s(S.an + r), as = ld2s(S.as + r);
          const double2 pc = ld2c(S.p + r);
          const double2 px = ld2c(po + r);
          const double xtra = s ? ld1c(po + r - 1) : ld1c(po + r + 2);
          const double2 pn = ld2c(po + r + g.HX), ps = ld2c(po + r - g.HX);
          // west/east neighbours: s=1 -> {ih-1, ih}, s=0 -> {ih, ih+1}
          const double wa = s ? xtra : px.x, ea = s ? px.x : px.y;
          const double wb = s ? px.x : px.y, eb = s ? px.y : xtra;
          double ra = bb.x - ae.x * ea - aw.x * wa - an.x * pn.x - as.x * ps.x;
          double rb = bb.y - ae.y * eb - aw.y * wb - an.y * pn.y - as.y * ps.y;
          if (DIM == 3) {
            const double2 at = ld2s(S.at + r), ab = ld2s(S.ab + r);
            const double2 pt = ld2c(po + r + g.hplane), pb = ld2c(po + r - g.hplane);
            ra = ra - at.x * pt.x - ab.x * pb.x;
            rb = rb - at.y * pt.y - ab.y * pb.y;
          }
          double2 out;
          out.x = ra / ap.x * relux + pc.x * omr;
          out.y = (ih + 1 < cnt) ? rb / ap.y * relux + pc.y * omr : pc.y;
          *reinterpret_cast<double2 *>(S.p + r) = out;
          if (YIMG == 1 && (j == 1 || j == g.n)) {           // even n: the image has the colour of the cell
            double *img = S.p + r + (j == 1 ? (long long)g.n * g.HX : -(long long)g.n * g.HX);
            img[0] = out.x;
            if (ih + 1 < cnt) img[1] = out.y;
          }
          if (with_err) emax = fmax(emax, fmax(fabs(out.x - pc.x), fabs(out.y - pc.y)));
        }
        if (YIMG == 2 && (j == 1 || j == g.n)) {             // odd n: this colour's halo rows image the other colour
          const int jh = (j == 1) ? g.n + 1 : 0;             // row n+1 <- row 1, row 0 <- row n
          const int sh = (colour + jh + k + g.koff) & 1;
          const int cnth = sh ? (g.m + 1) >> 1 : g.m >> 1;
          const long long dst = split_row(g, jh, k), src = split_row(g, j, k);
          if (ih < cnth) S.p[dst + ih] = ld1c(po + src + ih);
          if (ih + 1 < cnth) S.p[dst + ih + 1] = ld1c(po + src + ih + 1);
        }
      }
      // every cell of this colour is written and visible before the other colour reads it
      arrivals += gridDim.x;
      grid_barrier(A.bar, arrivals);
    }
  }
  // one reduction for the whole solve: non-negative doubles order like their bit patterns
  for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  __shared__ double wmax[PBX * PBY / 32];
  const int tid = threadIdx.y * PBX + threadIdx.x;
  if ((tid & 31) == 0) wmax[tid >> 5] = emax;
  __syncthreads();
  if (tid < 32) {
    double v = (tid < PBX * PBY / 32) ? wmax[tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0 && v > 0.0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(v));
  }
}

template <int DIM, int YIMG>
void launch(const Geo &g, PersistArgs &A, unsigned long long *err_bits, cudaStream_t st) {
  int dev = 0, sms = 0, per_sm = 0, coop = 0;
  PF_CUDA_OK(cudaGetDevice(&dev));
  PF_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) throw std::string("sor_variant 7 needs cooperative launch support");
  PF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sor_persistent_kernel<DIM, YIMG>, PBX * PBY, 0));
  if (per_sm < 1) throw std::string("sor_persistent_kernel does not fit on an SM");
  const long long ntiles = (long long)A.gx * A.gy * A.gz;
  const long long resident = (long long)sms * per_sm;
  const int blocks = (int)(ntiles < resident ? ntiles : resident);
  Geo gg = g;
  // the barrier word: stream-ordered allocation, zeroed, released after the launch (one solve = one launch)
  PF_CUDA_OK(cudaMallocAsync(reinterpret_cast<void **>(&A.bar), sizeof(unsigned int), st));
  PF_CUDA_OK(cudaMemsetAsync(A.bar, 0, sizeof(unsigned int), st));
  void *args[] = {&gg, &A, &err_bits};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void *)sor_persistent_kernel<DIM, YIMG>, dim3(blocks),
                                                    dim3(PBX, PBY, 1), args, 0, st);
  cudaFreeAsync(A.bar, st);
  PF_CUDA_OK(e);
  pf_count_launch();
}

}  // namespace

// one rank, and the solve is nothing but half-sweeps: every 2D case (self-kept y-halo rows), 3D air-condition
bool pf_persistent_applicable(const Geo &g, bool air, int nranks) {
  if (nranks != 1) return false;
  if (g.dim == 2) return pf_sor_stores_y_images(g);
  return air;
}

void k_sor_persistent(const Geo &g, const Phys &ph, const SplitSet S[2], int iters, unsigned long long *err_bits,
                      cudaStream_t st) {
  if (iters <= 0) return;
  PersistArgs A;
  A.S[0] = S[0];
  A.S[1] = S[1];
  A.iters = iters;
  A.first_colour = g.dim == 3 ? 0 : 1;
  const int pairs = ((g.m + 1) / 2 + 1) / 2;   // as k_sor_sweep
  A.gx = (pairs + PBX - 1) / PBX;
  A.gy = (g.n + PBY - 1) / PBY;
  A.gz = g.dim == 3 ? g.lz : 1;
  A.relux = ph.relux;
  if (g.dim == 3) launch<3, 0>(g, A, err_bits, st);
  else if (g.n % 2 == 0) launch<2, 1>(g, A, err_bits, st);
  else launch<2, 2>(g, A, err_bits, st);
}
