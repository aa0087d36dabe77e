// pf_voxel.cu -- voxel -> porosity: the 3-D "tanh filter" of the reference's tools/voxel2poro/voxel2poro.py.
//
// The reference smooths a binary voxel model into a porosity field with
//     porosity = scipy.ndimage.convolve(array_3d, kernel, mode='nearest')          (voxel2poro.py:33)
// where kernel = 1 - tanh(r / thickness) on a (2*int(14*thickness)+1)^3 cube, normalised (voxel2poro.py:189-197)
// -- 43^3 = 79,507 taps per voxel for the shipped thickness 1.5, ~45 s in scipy for the 32^3 sample and
// hours for the 256^3 grid of BASELINE configs[3].  Here: one thread per output voxel, the input row and
// the weight row of each (a0, a1) pair staged in shared memory, taps accumulated SEQUENTIALLY in double
// in scipy's order (C order of the input offsets, mul and add rounded separately: -fmad=false), result
// rounded to float32 -- the same bits as scipy's NI_Correlate (scipy 1.18: ni_filters.c; convolve =
// correlate with the weights reversed, no origin shift for odd sizes; weights with fabs(w) <= DBL_EPSILON are
// outside its footprint -- they are zeroed here, which leaves a sum of finite terms unchanged).
#include "pf_internal.cuh"

namespace {

constexpr int VTX = 128;   // outputs per block along the fastest axis
constexpr int VTY = 4;     // rows per block

__global__ void __launch_bounds__(VTX *VTY) convolve3d_nearest_kernel(const float *__restrict__ in, int n0, int n1, int n2,
                                                                      const double *__restrict__ w, int k0, int k1, int k2,
                                                                      float *__restrict__ out) {
  extern __shared__ double smem_d[];
  double *wrow = smem_d;                                       // k2 weights of the current (a0, a1)
  float *rows = reinterpret_cast<float *>(smem_d + k2);        // VTY x (VTX + k2 - 1) inputs
  const int span = VTX + k2 - 1;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * VTX + tx;
  const int i2 = blockIdx.x * VTX + tx, i1 = blockIdx.y * VTY + ty, i0 = blockIdx.z;
  const int h0 = k0 / 2, h1 = k1 / 2, h2 = k2 / 2;
  const int x0 = blockIdx.x * VTX - h2;                        // input coordinate of rows[.][0]
  double acc = 0.0;
  for (int a0 = 0; a0 < k0; ++a0) {
    const int z = min(max(i0 + a0 - h0, 0), n0 - 1);           // mode='nearest': clamp to the edge voxel
    for (int a1 = 0; a1 < k1; ++a1) {
      __syncthreads();
      // reversed weights: convolve(in, w)[i] = sum_a w[K-1-a] * in[i + a - K/2]
      if (tid < k2) {
        const double wv = w[((size_t)(k0 - 1 - a0) * k1 + (k1 - 1 - a1)) * k2 + (k2 - 1 - tid)];
        wrow[tid] = fabs(wv) > 2.220446049250313e-16 ? wv : 0.0;   // NI_Correlate's footprint test (DBL_EPSILON)
      }
      {
        const int y = min(max(i1 + a1 - h1, 0), n1 - 1);
        const float *src = in + ((size_t)z * n1 + y) * n2;
        float *dst = rows + ty * span;
        for (int t = tx; t < span; t += VTX) dst[t] = src[min(max(x0 + t, 0), n2 - 1)];
      }
      __syncthreads();
      const float *r = rows + ty * span + tx;
#pragma unroll 4
      for (int a2 = 0; a2 < k2; ++a2) acc = acc + (double)r[a2] * wrow[a2];
    }
  }
  if (i1 < n1 && i2 < n2) out[((size_t)i0 * n1 + i1) * n2 + i2] = (float)acc;
}

}  // namespace

extern "C" int pf_convolve3d_nearest(const float *in, int n0, int n1, int n2, const double *weights, int k0, int k1,
                                     int k2, float *out, int device) {
  float *d_in = nullptr, *d_out = nullptr;
  double *d_w = nullptr;
  int rc = 0;
  try {
    if (!in || !weights || !out) throw std::string("pf_convolve3d_nearest: null array");
    if (n0 < 1 || n1 < 1 || n2 < 1) throw std::string("pf_convolve3d_nearest: empty input");
    if (k0 < 1 || k1 < 1 || k2 < 1 || !(k0 & 1) || !(k1 & 1) || !(k2 & 1))
      throw std::string("pf_convolve3d_nearest: kernel sizes must be odd (the reference's are 2*int(14*thickness)+1)");
    if (k2 > VTX * VTY) throw std::string("pf_convolve3d_nearest: kernel wider than 512 taps per row");
    if (n0 > 65535 || (n1 + VTY - 1) / VTY > 65535) throw std::string("pf_convolve3d_nearest: grid too large");
    if (device >= 0) PF_CUDA_OK(cudaSetDevice(device));
    const size_t nvox = (size_t)n0 * n1 * n2, nw = (size_t)k0 * k1 * k2;
    PF_CUDA_OK(cudaMalloc(&d_in, nvox * sizeof(float)));
    PF_CUDA_OK(cudaMalloc(&d_out, nvox * sizeof(float)));
    PF_CUDA_OK(cudaMalloc(&d_w, nw * sizeof(double)));
    PF_CUDA_OK(cudaMemcpy(d_in, in, nvox * sizeof(float), cudaMemcpyHostToDevice));
    PF_CUDA_OK(cudaMemcpy(d_w, weights, nw * sizeof(double), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)k2 * sizeof(double) + (size_t)VTY * (VTX + k2 - 1) * sizeof(float);
    const dim3 grid((n2 + VTX - 1) / VTX, (n1 + VTY - 1) / VTY, n0);
    convolve3d_nearest_kernel<<<grid, dim3(VTX, VTY, 1), smem>>>(d_in, n0, n1, n2, d_w, k0, k1, k2, d_out);
    pf_count_launch();
    PF_CUDA_OK(cudaGetLastError());
    PF_CUDA_OK(cudaMemcpy(out, d_out, nvox * sizeof(float), cudaMemcpyDeviceToHost));
  } catch (const std::string &e) {
    pf_set_global_error(e);
    rc = 1;
  }
  cudaFree(d_in);
  cudaFree(d_out);
  cudaFree(d_w);
  return rc;
}
