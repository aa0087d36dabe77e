// pf_output.cu -- the bodies of the reference's ASCII VTK snapshots, formatted on the GPU.
//
// output_paraview_temp_3d / _2d (lib/output.f90:968-1088 / :421-537) write every point of every field with
// the format "(3(f16.4,1x))": 210 bytes per cell and snapshot in 3D -- 3.5 GB at 256^3, and with the shipped
// decks' snapshot cadence the formatted write costs several times the time steps it reports on.  Here one
// thread formats one record (fixed length: 3 x f16.4 separated by blanks = 50 characters + newline, or one
// f16.4 + newline; gfortran drops the trailing 1x of a record), the derived fields (velocityInFluid,
// dimless_v, VelocityDivergent, abs_dimless_v) are computed in the same thread in the reference's expression
// order, and a block's records leave through shared memory as 16-byte stores.  The host only copies the
// finished text to the file.
//
// f16.4 is exact: the double is decomposed into mantissa * 2^e, multiplied by 10^4 in 128-bit integers and
// rounded half-to-even on the exact remainder -- what glibc's printf("%16.4f") and libgfortran (which formats
// through snprintf) produce; values too wide for 16 columns become asterisks, NaN / Infinity are spelled the
// gfortran way.
#include "pf_internal.cuh"

namespace {

constexpr int OB = 256;                 // records per block
constexpr int VEC_LEN = 51, SCA_LEN = 17;

__device__ void fmt_f16_4(double x, char *d) {
#pragma unroll
  for (int q = 0; q < 16; ++q) d[q] = ' ';
  const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
  const bool neg = bits >> 63;
  const int be = (int)((bits >> 52) & 0x7ff);
  const unsigned long long frac = bits & ((1ull << 52) - 1);
  if (be == 0x7ff) {
    if (frac) { d[13] = 'N'; d[14] = 'a'; d[15] = 'N'; return; }
    const char *t = "Infinity";
    for (int q = 0; q < 8; ++q) d[8 + q] = t[q];
    if (neg) d[7] = '-';
    return;
  }
  if (be - 1023 >= 37) {                // |x| >= 2^37 > 1e11: cannot fit 16 columns with 4 decimals
    for (int q = 0; q < 16; ++q) d[q] = '*';
    return;
  }
  const unsigned long long mant = be ? (frac | (1ull << 52)) : frac;
  const int s = 1075 - (be ? be : 1);   // x = mant * 2^-s, s in [16, 1074]
  unsigned long long n;                 // round_half_even(|x| * 10^4)
  if (s >= 120) {
    n = 0;
  } else {
    const unsigned __int128 M = (unsigned __int128)mant * 10000u;
    unsigned __int128 q = M >> s;
    const unsigned __int128 rem = M - (q << s), half = (unsigned __int128)1 << (s - 1);
    if (rem > half || (rem == half && (q & 1))) q += 1;
    n = (unsigned long long)q;
  }
  unsigned long long ip = n / 10000u;
  unsigned int fp = (unsigned int)(n % 10000u);
  int ndig = 1;
  for (unsigned long long t = ip; t >= 10; t /= 10) ++ndig;
  if ((neg ? 1 : 0) + ndig + 5 > 16) {
    for (int q = 0; q < 16; ++q) d[q] = '*';
    return;
  }
  for (int q = 15; q >= 12; --q) { d[q] = (char)('0' + fp % 10); fp /= 10; }
  d[11] = '.';
  int pos = 10;
  do { d[pos--] = (char)('0' + (int)(ip % 10)); ip /= 10; } while (ip);
  if (neg) d[pos] = '-';
}

struct OutArgs {
  int section, k0, nplanes;
  const double *xp, *yp, *zp;           // device copies; zp indexed by the global k (null in 2D)
  double uin;
};

__global__ void __launch_bounds__(OB) vtk_section_kernel(Geo g, Fields f, OutArgs a, char *out, long long nrec) {
  __shared__ __align__(16) char sh[OB * VEC_LEN];
  const bool vec = a.section <= PF_VTK_DIMLESS_V;
  const int len = vec ? VEC_LEN : SCA_LEN;
  const long long r0 = (long long)blockIdx.x * OB;
  const long long r = r0 + threadIdx.x;
  if (r < nrec) {
    const int i = (int)(r % g.m) + 1;
    const long long t = r / g.m;
    const int j = (int)(t % g.n) + 1;
    const int kl = g.dim == 3 ? (int)(t / g.n) + a.k0 : 0;
    const long long c = nat_idx(g, i, j, kl);
    const bool d3 = g.dim == 3;
    double v0 = 0., v1 = 0., v2 = 0.;
    switch (a.section) {
      case PF_VTK_POINTS: v0 = a.xp[i]; v1 = a.yp[j]; v2 = d3 ? a.zp[kl + g.koff] : 0.; break;
      case PF_VTK_VELOCITY: v0 = f.u[c]; v1 = f.v[c]; v2 = d3 ? f.w[c] : 0.; break;
      case PF_VTK_VELOCITY_IN_FLUID: {
        const double e = f.eps[c];
        v0 = f.u[c] * e; v1 = f.v[c] * e; v2 = d3 ? f.w[c] * e : 0.;
        break;
      }
      case PF_VTK_DIMLESS_V: {          // 2D only (:468-474)
        const double e = f.eps[c];
        v0 = f.u[c] * e / a.uin; v1 = f.v[c] * e / a.uin; v2 = 0.;
        break;
      }
      case PF_VTK_POROSITY: v0 = f.eps[c]; break;
      case PF_VTK_PRESSURE: v0 = f.p[c]; break;
      case PF_VTK_DIVERGENT: {          // :1060-1064 / :500
        v0 = (f.u[c + 1] - f.u[c - 1]) / (a.xp[i + 1] - a.xp[i - 1]) +
             (f.v[c + g.NX] - f.v[c - g.NX]) / (a.yp[j + 1] - a.yp[j - 1]);
        if (d3) {
          const int kg = kl + g.koff;
          v0 = v0 + (f.w[c + g.plane] - f.w[c - g.plane]) / (a.zp[kg + 1] - a.zp[kg - 1]);
        }
        break;
      }
      case PF_VTK_ABS_DIMLESS_V: {      // 2D only (:523)
        const double e = f.eps[c];
        const double a0 = f.u[c] * e / a.uin, a1 = f.v[c] * e / a.uin;
        v0 = sqrt(a0 * a0 + a1 * a1);
        break;
      }
    }
    char *d = sh + threadIdx.x * len;
    fmt_f16_4(v0, d);
    if (vec) {
      d[16] = ' ';
      fmt_f16_4(v1, d + 17);
      d[33] = ' ';
      fmt_f16_4(v2, d + 34);
      d[50] = '\n';
    } else {
      d[16] = '\n';
    }
  }
  __syncthreads();
  const long long first = r0 * len;
  const long long nbytes = (long long)min((long long)OB, nrec - r0) * len;
  if (nbytes == (long long)OB * len) {  // whole block: 16-byte stores (block starts are multiples of 16 bytes)
    const int4 *src = reinterpret_cast<const int4 *>(sh);
    int4 *dst = reinterpret_cast<int4 *>(out + first);
    for (int q = threadIdx.x; q < OB * len / 16; q += OB) dst[q] = src[q];
  } else {
    for (long long q = threadIdx.x; q < nbytes; q += OB) out[first + q] = sh[q];
  }
}

}  // namespace

bool pf_vtk_section_valid(const Geo &g, int section) {
  if (section < PF_VTK_POINTS || section > PF_VTK_ABS_DIMLESS_V) return false;
  if (g.dim == 3 && (section == PF_VTK_DIMLESS_V || section == PF_VTK_ABS_DIMLESS_V)) return false;
  return true;
}

size_t pf_vtk_record_bytes(int section) { return section <= PF_VTK_DIMLESS_V ? VEC_LEN : SCA_LEN; }

// formats `nplanes` local planes starting at k0 (3D) or the whole field (2D) into the device buffer `out`
void k_vtk_section(const Geo &g, const Fields &f, int section, int k0, int nplanes, const double *xp_dev,
                   const double *yp_dev, const double *zp_dev, double inlet_velocity, char *out, cudaStream_t st) {
  OutArgs a;
  a.section = section; a.k0 = k0; a.nplanes = nplanes;
  a.xp = xp_dev; a.yp = yp_dev; a.zp = zp_dev;
  a.uin = inlet_velocity;
  const long long nrec = (long long)g.m * g.n * (g.dim == 3 ? nplanes : 1);
  if (nrec <= 0) return;
  static_assert((OB * VEC_LEN) % 16 == 0 && (OB * SCA_LEN) % 16 == 0, "block text is a whole number of int4");
  vtk_section_kernel<<<(unsigned)((nrec + OB - 1) / OB), OB, 0, st>>>(g, f, a, out, nrec);
  pf_count_launch();
}
