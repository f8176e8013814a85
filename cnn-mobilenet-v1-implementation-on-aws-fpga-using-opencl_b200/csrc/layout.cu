// layout.cu — boundary conversions and the on-device synthetic image stream.
//
// The reference's host contract is planar [N][C][H][W] (kernel.cl:14,56,73,90,103,107);
// the device keeps NHWC so that channels are the vectorised inner dimension and a feature
// map is directly the K-major A operand of the pointwise GEMM.  These kernels run only at
// the boundary (mnv1_upload_planar / mnv1_download_planar), never between layers.
#include "common.cuh"

namespace mnv1 {

// 32x32 smem transpose between the (c) and (h*w) axes of one image.
template <typename T, bool TO_NHWC>
__global__ void __launch_bounds__(256) transpose_kernel(T* __restrict__ nhwc, float* __restrict__ nchw,
                                                        int c, int hw) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long base = (long)img * c * hw;
  if (TO_NHWC) {
    for (int j = ty; j < 32; j += 8) {  // read planar: pixel fastest
      int cc = c0 + j, p = p0 + tx;
      tile[j][tx] = (cc < c && p < hw) ? nchw[base + (long)cc * hw + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {  // write nhwc: channel fastest
      int p = p0 + j, cc = c0 + tx;
      if (cc < c && p < hw) nhwc[base + (long)p * c + cc] = from_f32<T>(tile[tx][j]);
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      int p = p0 + j, cc = c0 + tx;
      tile[j][tx] = (cc < c && p < hw) ? to_f32<T>(nhwc[base + (long)p * c + cc]) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      int cc = c0 + j, p = p0 + tx;
      if (cc < c && p < hw) nchw[base + (long)cc * hw + p] = tile[tx][j];
    }
  }
}

template <bool TO_NHWC>
static cudaError_t launch_transpose(mnv1_dtype dt, void* nhwc, float* nchw, int n, int c, int h, int w,
                                    cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(256);
  if (dt == MNV1_BF16)
    transpose_kernel<bf16, TO_NHWC><<<grid, block, 0, st>>>((bf16*)nhwc, nchw, c, hw);
  else
    transpose_kernel<float, TO_NHWC><<<grid, block, 0, st>>>((float*)nhwc, nchw, c, hw);
  return cudaGetLastError();
}

cudaError_t launch_nchw_to_nhwc(mnv1_dtype dt, void* out_nhwc, const float* in_nchw, int n, int c, int h,
                                int w, cudaStream_t st) {
  return launch_transpose<true>(dt, out_nhwc, const_cast<float*>(in_nchw), n, c, h, w, st);
}
cudaError_t launch_nhwc_to_nchw(mnv1_dtype dt, float* out_nchw, const void* in_nhwc, int n, int c, int h,
                                int w, cudaStream_t st) {
  return launch_transpose<false>(dt, const_cast<void*>(in_nhwc), out_nchw, n, c, h, w, st);
}

// Synthetic image stream of SURVEY §8(d): byte j of the stream is byte (j % 8) of
// splitmix64(seed, j / 8).  Must stay bit-identical to synth.py:images().
__device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t counter) {
  uint64_t z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) synth_images_kernel(uint64_t* __restrict__ out, uint64_t first_word,
                                                           long nwords, uint64_t seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long)gridDim.x * blockDim.x)
    out[i] = splitmix64(seed, first_word + (uint64_t)i);
}
cudaError_t launch_synth_images(uint8_t* out, long first_byte, long nbytes, uint64_t seed, cudaStream_t st) {
  if ((first_byte | nbytes) & 7) return cudaErrorInvalidValue;
  if (nbytes == 0) return cudaSuccess;
  long nwords = nbytes / 8;
  int grid = (int)((nwords + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  synth_images_kernel<<<grid, 256, 0, st>>>((uint64_t*)out, (uint64_t)(first_byte / 8), nwords, seed);
  return cudaGetLastError();
}

}  // namespace mnv1
