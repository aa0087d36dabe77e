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
#include <stdlib.h>
#include <string.h>

#include "pf_internal.cuh"

namespace {

constexpr int WIN = 32;          // bytes of text per thread

struct IngestOut {
  double *eps;                   // (m+2) x (n+2) x (l+2 | 1), Fortran order
  unsigned long long *count;     // records stored
  unsigned long long *nflag;     // records left to the host
  unsigned long long *flagged;   // their byte offsets (first `flag_cap`)
  unsigned long long *nbad;      // malformed / out-of-range records
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

__global__ void ingest_kernel(const char *__restrict__ t, size_t n, int m, int nn, int l, int d3, double threshold,
                              IngestOut o) {
  const size_t w0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * WIN;
  unsigned long long stored = 0;
  for (size_t s = w0; s < w0 + WIN && s < n; ++s) {
    if (s != 0 && t[s - 1] != '\n') continue;                         // not a line start
    size_t p = s;
    while (p < n && is_sep(t[p])) ++p;
    if (p >= n || t[p] == '\n') continue;                             // blank line
    long long x = 0, y = 0, z = 0;
    double v = 0.0;
    int st = 2;
    if (parse_int(t, n, p, x) && parse_int(t, n, p, y) && parse_int(t, n, p, z)) st = parse_real(t, n, p, v);
    // anything after the fourth item is ignored, as list-directed input does (the next READ starts a new record)
    if (st == 2 || x < 1 || x > m || y < 1 || y > nn || (d3 && (z < 1 || z > l))) {
      atomicAdd(o.nbad, 1ull);
      continue;
    }
    if (st == 1) {
      const unsigned long long slot = atomicAdd(o.nflag, 1ull);
      if (slot < o.flag_cap) o.flagged[slot] = (unsigned long long)s;
      continue;
    }
    const size_t idx = (size_t)x + (size_t)(m + 2) * ((size_t)y + (size_t)(nn + 2) * (size_t)(d3 ? z : 0));
    o.eps[idx] = fmax(v, threshold);                                   // lib/grid.f90:289 / :44
    ++stored;
  }
  if (stored) atomicAdd(o.count, stored);
}

}  // namespace

extern "C" int pf_parse_porosity_csv(const char *text, size_t nbytes, int m, int n, int l, double threshold,
                                     double *porosity, long long *nrecords, int device) {
  char *d_text = nullptr;
  double *d_eps = nullptr;
  unsigned long long *d_ctr = nullptr;
  int rc = 0;
  try {
    if (!text || !porosity) throw std::string("pf_parse_porosity_csv: null argument");
    if (m < 1 || n < 1 || l < 0) throw std::string("pf_parse_porosity_csv: bad dimensions");
    if (device >= 0) PF_CUDA_OK(cudaSetDevice(device));
    const int d3 = l > 0;
    const size_t elems = (size_t)(m + 2) * (n + 2) * (d3 ? (size_t)l + 2 : 1);
    const unsigned long long flag_cap = 1 << 20;
    PF_CUDA_OK(cudaMalloc(&d_text, nbytes + 1));
    PF_CUDA_OK(cudaMalloc(&d_eps, elems * sizeof(double)));
    PF_CUDA_OK(cudaMalloc(&d_ctr, (3 + flag_cap) * sizeof(unsigned long long)));
    PF_CUDA_OK(cudaMemcpy(d_text, text, nbytes, cudaMemcpyHostToDevice));
    PF_CUDA_OK(cudaMemcpy(d_eps, porosity, elems * sizeof(double), cudaMemcpyHostToDevice));   // keep what is there
    PF_CUDA_OK(cudaMemset(d_ctr, 0, 3 * sizeof(unsigned long long)));
    IngestOut o;
    o.eps = d_eps; o.count = d_ctr; o.nflag = d_ctr + 1; o.nbad = d_ctr + 2; o.flagged = d_ctr + 3; o.flag_cap = flag_cap;
    const size_t nthreads = (nbytes + WIN - 1) / WIN;
    if (nthreads) {
      ingest_kernel<<<(unsigned)((nthreads + 255) / 256), 256>>>(d_text, nbytes, m, n, l, d3, threshold, o);
      pf_count_launch();
      PF_CUDA_OK(cudaGetLastError());
    }
    unsigned long long ctr[3];
    PF_CUDA_OK(cudaMemcpy(ctr, d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost));
    if (ctr[2]) throw std::string("pf_parse_porosity_csv: ") + std::to_string(ctr[2]) + " malformed or out-of-range record(s)";
    if (ctr[1] > flag_cap) throw std::string("pf_parse_porosity_csv: too many records need extended-precision conversion");
    PF_CUDA_OK(cudaMemcpy(porosity, d_eps, elems * sizeof(double), cudaMemcpyDeviceToHost));
    if (ctr[1]) {   // the few records outside the exact one-operation range: strtod, same rule
      std::vector<unsigned long long> off(ctr[1]);
      PF_CUDA_OK(cudaMemcpy(off.data(), d_ctr + 3, ctr[1] * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      for (unsigned long long s : off) {
        const void *nl = memchr(text + s, '\n', nbytes - s);
        std::string line(text + s, nl ? (size_t)(static_cast<const char *>(nl) - (text + s)) : nbytes - s);
        for (char &c : line) if (c == ',' || c == 'd' || c == 'D') c = (c == ',') ? ' ' : 'e';
        long long x = 0, y = 0, z = 0;
        double v = 0;
        char *end = nullptr;
        const char *q = line.c_str();
        x = strtoll(q, &end, 10); if (*end == '.') { ++end; while (*end == '0') ++end; } q = end;
        y = strtoll(q, &end, 10); if (*end == '.') { ++end; while (*end == '0') ++end; } q = end;
        z = strtoll(q, &end, 10); if (*end == '.') { ++end; while (*end == '0') ++end; } q = end;
        v = strtod(q, &end);
        if (end == q) throw std::string("pf_parse_porosity_csv: malformed record at byte ") + std::to_string(s);
        porosity[(size_t)x + (size_t)(m + 2) * ((size_t)y + (size_t)(n + 2) * (size_t)(d3 ? z : 0))] = v > threshold ? v : threshold;
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
  return rc;
}
