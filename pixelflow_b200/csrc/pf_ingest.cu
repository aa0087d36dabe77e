// pf_ingest.cu -- porosity CSV records parsed on the GPU.
//
// The reference reads its input with one list-directed `read(52,*) x, y, z, poro_val` per cell and stores
// porosity(x,y,z) = max(poro_val, threshold) (lib/grid.f90:281-294, 2D :38-47): at the 1024x512x512 size that
// is 268 M text records, ~10 GB -- minutes of formatted I/O before the first time step.  Records carry their
// own indices, so their order is irrelevant: here the text is copied to the device once and every thread
// looks for the line starts inside its own 32-byte window, parses those lines (three integers, one real in
// any of the F / E / D forms list-directed input accepts, separated by commas and/or blanks) and scatters
// the value into the array.  Decimal -> binary is exact where one IEEE operation suffices (<= 15 significant
// digits after stripping, |decimal exponent| <= 22: mantissa * 10^e or mantissa / 10^-e with both operands
// exact -- every value stl2poro (.6E) and voxel2poro (.10f) write); anything else is reported back by byte
// offset and converted by the caller with strtod, so the result always equals a correctly rounded read.
//
// What the reference's loop does beyond "one record per cell" is kept (round 2): it executes exactly m*n*l READs, so
// records after the first m*n*l non-blank lines are never looked at (a prefix sum of the line starts gives every record
// its ordinal); the indices address porosity(0:md,0:nd,0:ld), so 0 and m+1 are legal (halo cells, overwritten by the
// halo rules later); and of several records for one cell the LAST one read wins (the largest ordinal, found with an
// atomic maximum per cell; a second pass runs only if a duplicate was seen).
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_scan.cuh>

#include "pf_internal.cuh"

namespace {

constexpr int WIN = 32;          // bytes of text per thread

struct IngestOut {
  double *eps;                   // (m+2) x (n+2) x (l+2 | 1), Fortran order
  unsigned long long *owner;     // per cell: 1 + ordinal of the last record read for it (0 = none)
  unsigned long long *count;     // records stored
  unsigned long long *nflag;     // records left to the host
  unsigned long long *flagged;   // their (byte offset, ordinal) pairs (first `flag_cap`)
  unsigned long long *nbad;      // malformed / out-of-range records among those the reference would read
  unsigned long long *ndup;      // records for a cell that already had one
  unsigned long long flag_cap;
};

__device__ __forceinline__ bool is_sep(char c) { return c == ' ' || c == ',' || c == '\t' || c == '\r'; }

// parses a decimal integer at p (after separators); returns false if none
__device__ bool parse_int(const char *t, size_t n, size_t &p, long long &v) {
  while (p < n && is_sep(t[p])) ++p;
  bool neg = false;
  if (p < n && (t[p] == '+' || t[p] == '-')) { neg = t[p] == '-'; ++p; }
  if (p >= n || t[p] < '0' || t[p] > '9') return false;
  long long a = 0;
  while (p < n && t[p] >= '0' && t[p] <= '9') { if (a < (1ll << 40)) a = a * 10 + (t[p] - '0'); ++p; }
  if (p < n && t[p] == '.') {    // "12." or "12.000": list-directed input of an integer item would reject it;
    ++p;                         // accept a zero fraction, which is what index columns written as reals look like
    while (p < n && t[p] == '0') ++p;
    if (p < n && t[p] >= '1' && t[p] <= '9') return false;
  }
  v = neg ? -a : a;
  return true;
}

// status: 0 = ok (exact), 1 = needs the host (too many digits / large exponent), 2 = malformed
__device__ int parse_real(const char *t, size_t n, size_t &p, double &v) {
  while (p < n && is_sep(t[p])) ++p;
  bool neg = false;
  if (p < n && (t[p] == '+' || t[p] == '-')) { neg = t[p] == '-'; ++p; }
  unsigned long long mant = 0;
  int ndig = 0, dropped = 0, e10 = 0;
  bool any = false, inexact = false;
  auto digit = [&](int d, bool frac) {
    any = true;
    if (mant == 0 && d == 0) { if (frac) --e10; return; }          // leading zeros carry no digits
    if (ndig < 18) { mant = mant * 10 + d; ++ndig; if (frac) --e10; }
    else { if (d) inexact = true; ++dropped; if (!frac) ++e10; }
  };
  while (p < n && t[p] >= '0' && t[p] <= '9') digit(t[p++] - '0', false);
  if (p < n && t[p] == '.') {
    ++p;
    while (p < n && t[p] >= '0' && t[p] <= '9') digit(t[p++] - '0', true);
  }
  if (!any) return 2;
  if (p < n && (t[p] == 'e' || t[p] == 'E' || t[p] == 'd' || t[p] == 'D' || t[p] == '+' || t[p] == '-')) {
    if (t[p] != '+' && t[p] != '-') ++p;
    bool eneg = false;
    if (p < n && (t[p] == '+' || t[p] == '-')) { eneg = t[p] == '-'; ++p; }
    if (p >= n || t[p] < '0' || t[p] > '9') return 2;
    int ex = 0;
    while (p < n && t[p] >= '0' && t[p] <= '9') { if (ex < 10000) ex = ex * 10 + (t[p] - '0'); ++p; }
    e10 += eneg ? -ex : ex;
  }
  (void)dropped;
  if (mant == 0) { v = neg ? -0.0 : 0.0; return 0; }
  while (mant % 10 == 0) { mant /= 10; ++e10; }                       // strip trailing zeros: 1.000000E-06 -> 1 E-6
  if (inexact || mant >= (1ull << 53) || e10 > 22 || e10 < -22) return 1;
  double p10 = 1.0;
  for (int q = 0; q < (e10 < 0 ? -e10 : e10); ++q) p10 *= 10.0;       // exact up to 10^22
  double r;
  if (e10 >= 0) {
    r = (double)mant * p10;                                           // one rounding: correctly rounded
  } else {
    r = (double)mant / p10;                                           // IEEE division of two exact operands
  }
  v = neg ? -r : r;
  return 0;
}

// a record starts at s: the first byte of a line that is not blank
__device__ __forceinline__ bool record_start(const char *t, size_t n, size_t s, size_t &p) {
  if (s != 0 && t[s - 1] != '\n') return false;
  p = s;
  while (p < n && is_sep(t[p])) ++p;
  return p < n && t[p] != '\n';
}

__global__ void count_records_kernel(const char *__restrict__ t, size_t n, size_t nthreads, unsigned long long *cnt) {
  const size_t th = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (th >= nthreads) return;
  const size_t w0 = th * WIN;
  unsigned long long c = 0;
  size_t p;
  for (size_t s = w0; s < w0 + WIN && s < n; ++s) c += record_start(t, n, s, p) ? 1 : 0;
  cnt[th] = c;
}

// PASS 0: claim cells (atomic maximum of the ordinal) and store; PASS 1 (only after duplicates were seen): store again,
// the final owner of each cell only -- the order of two plain stores to one cell is not defined
template <int PASS>
__global__ void ingest_kernel(const char *__restrict__ t, size_t n, size_t nthreads, int m, int nn, int l, int d3,
                              double threshold, unsigned long long limit, const unsigned long long *__restrict__ first,
                              IngestOut o) {
  const size_t th = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (th >= nthreads) return;
  const size_t w0 = th * WIN;
  unsigned long long stored = 0, ord = first[th];
  for (size_t s = w0; s < w0 + WIN && s < n; ++s) {
    size_t p;
    if (!record_start(t, n, s, p)) continue;
    const unsigned long long my = ord++;
    if (my >= limit) break;                                           // the reference has stopped reading (m*n*l READs)
    long long x = 0, y = 0, z = 0;
    double v = 0.0;
    int st = 2;
    if (parse_int(t, n, p, x) && parse_int(t, n, p, y) && parse_int(t, n, p, z)) st = parse_real(t, n, p, v);
    // anything after the fourth item is ignored, as list-directed input does (the next READ starts a new record)
    if (st == 2 || x < 0 || x > m + 1 || y < 0 || y > nn + 1 || (d3 && (z < 0 || z > l + 1))) {
      if (PASS == 0) atomicAdd(o.nbad, 1ull);
      continue;
    }
    const size_t idx = (size_t)x + (size_t)(m + 2) * ((size_t)y + (size_t)(nn + 2) * (size_t)(d3 ? z : 0));
    if (PASS == 0) {
      const unsigned long long prev = atomicMax(o.owner + idx, my + 1);
      if (prev != 0) atomicAdd(o.ndup, 1ull);
      if (st == 1) {
        const unsigned long long slot = atomicAdd(o.nflag, 1ull);
        if (slot < o.flag_cap) { o.flagged[2 * slot] = (unsigned long long)s; o.flagged[2 * slot + 1] = my; }
        continue;
      }
      if (prev < my + 1) o.eps[idx] = fmax(v, threshold);             // lib/grid.f90:289 / :44
      ++stored;
    } else if (st == 0 && o.owner[idx] == my + 1) {
      o.eps[idx] = fmax(v, threshold);
    }
  }
  if (PASS == 0 && stored) atomicAdd(o.count, stored);
}

}  // namespace

extern "C" int pf_parse_porosity_csv(const char *text, size_t nbytes, int m, int n, int l, double threshold,
                                     double *porosity, long long *nrecords, int device) {
  char *d_text = nullptr;
  double *d_eps = nullptr;
  unsigned long long *d_ctr = nullptr, *d_owner = nullptr, *d_first = nullptr;
  void *d_tmp = nullptr;
  int rc = 0;
  try {
    if (!text || !porosity) throw std::string("pf_parse_porosity_csv: null argument");
    if (m < 1 || n < 1 || l < 0) throw std::string("pf_parse_porosity_csv: bad dimensions");
    if (device >= 0) PF_CUDA_OK(cudaSetDevice(device));
    const int d3 = l > 0;
    const size_t elems = (size_t)(m + 2) * (n + 2) * (d3 ? (size_t)l + 2 : 1);
    const unsigned long long flag_cap = 1 << 20;
    const unsigned long long limit = (unsigned long long)m * n * (d3 ? l : 1);    // the reference's m*n*l READs
    const size_t nthreads = (nbytes + WIN - 1) / WIN;
    PF_CUDA_OK(cudaMalloc(&d_text, nbytes + 1));
    PF_CUDA_OK(cudaMalloc(&d_eps, elems * sizeof(double)));
    PF_CUDA_OK(cudaMalloc(&d_owner, elems * sizeof(unsigned long long)));
    PF_CUDA_OK(cudaMalloc(&d_ctr, (4 + 2 * flag_cap) * sizeof(unsigned long long)));
    PF_CUDA_OK(cudaMalloc(&d_first, (nthreads + 1) * sizeof(unsigned long long)));
    PF_CUDA_OK(cudaMemcpy(d_text, text, nbytes, cudaMemcpyHostToDevice));
    PF_CUDA_OK(cudaMemcpy(d_eps, porosity, elems * sizeof(double), cudaMemcpyHostToDevice));   // keep what is there
    PF_CUDA_OK(cudaMemset(d_owner, 0, elems * sizeof(unsigned long long)));
    PF_CUDA_OK(cudaMemset(d_ctr, 0, 4 * sizeof(unsigned long long)));
    IngestOut o;
    o.eps = d_eps; o.owner = d_owner; o.count = d_ctr; o.nflag = d_ctr + 1; o.nbad = d_ctr + 2; o.ndup = d_ctr + 3;
    o.flagged = d_ctr + 4; o.flag_cap = flag_cap;
    unsigned long long ctr[4] = {0, 0, 0, 0};
    if (nthreads) {
      const unsigned grid = (unsigned)((nthreads + 255) / 256);
      count_records_kernel<<<grid, 256>>>(d_text, nbytes, nthreads, d_first);
      pf_count_launch();
      size_t tmp_bytes = 0;   // ordinal of the first record of every window: exclusive prefix sum of the counts, in place
      PF_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_first, d_first, (int)nthreads));
      PF_CUDA_OK(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
      PF_CUDA_OK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_first, d_first, (int)nthreads));
      ingest_kernel<0><<<grid, 256>>>(d_text, nbytes, nthreads, m, n, l, d3, threshold, limit, d_first, o);
      pf_count_launch();
      PF_CUDA_OK(cudaGetLastError());
      PF_CUDA_OK(cudaMemcpy(ctr, d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost));
      if (ctr[3]) {   // some cell has more than one record: the last one read wins
        ingest_kernel<1><<<grid, 256>>>(d_text, nbytes, nthreads, m, n, l, d3, threshold, limit, d_first, o);
        pf_count_launch();
        PF_CUDA_OK(cudaGetLastError());
      }
    }
    if (ctr[2]) throw std::string("pf_parse_porosity_csv: ") + std::to_string(ctr[2]) + " malformed or out-of-range record(s)";
    if (ctr[1] > flag_cap) throw std::string("pf_parse_porosity_csv: too many records need extended-precision conversion");
    PF_CUDA_OK(cudaMemcpy(porosity, d_eps, elems * sizeof(double), cudaMemcpyDeviceToHost));
    if (ctr[1]) {   // the few records outside the exact one-operation range: strtod, same rule
      std::vector<unsigned long long> off(2 * ctr[1]);
      PF_CUDA_OK(cudaMemcpy(off.data(), d_ctr + 4, 2 * ctr[1] * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      for (unsigned long long q = 0; q < ctr[1]; ++q) {
        const unsigned long long s = off[2 * q], ordinal = off[2 * q + 1];
        const void *nl = memchr(text + s, '\n', nbytes - s);
        std::string line(text + s, nl ? (size_t)(static_cast<const char *>(nl) - (text + s)) : nbytes - s);
        for (char &c : line) if (c == ',' || c == 'd' || c == 'D') c = (c == ',') ? ' ' : 'e';
        long long x = 0, y = 0, z = 0;
        double v = 0;
        char *end = nullptr;
        const char *qq = line.c_str();
        x = strtoll(qq, &end, 10); if (*end == '.') { ++end; while (*end == '0') ++end; } qq = end;
        y = strtoll(qq, &end, 10); if (*end == '.') { ++end; while (*end == '0') ++end; } qq = end;
        z = strtoll(qq, &end, 10); if (*end == '.') { ++end; while (*end == '0') ++end; } qq = end;
        v = strtod(qq, &end);
        if (end == qq) throw std::string("pf_parse_porosity_csv: malformed record at byte ") + std::to_string(s);
        const size_t idx = (size_t)x + (size_t)(m + 2) * ((size_t)y + (size_t)(n + 2) * (size_t)(d3 ? z : 0));
        unsigned long long own = 0;   // the last record read for this cell wins
        PF_CUDA_OK(cudaMemcpy(&own, d_owner + idx, sizeof(own), cudaMemcpyDeviceToHost));
        if (own == ordinal + 1) porosity[idx] = v > threshold ? v : threshold;
      }
    }
    if (nrecords) *nrecords = (long long)(ctr[0] + ctr[1]);
  } catch (const std::string &e) {
    pf_set_global_error(e);
    rc = 1;
  }
  cudaFree(d_text);
  cudaFree(d_eps);
  cudaFree(d_ctr);
  cudaFree(d_owner);
  cudaFree(d_first);
  cudaFree(d_tmp);
  return rc;
}
