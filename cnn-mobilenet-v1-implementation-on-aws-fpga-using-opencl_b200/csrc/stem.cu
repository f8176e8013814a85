// stem.cu — layer 1: 3x3 conv over the three u8 colour planes -> 32 feature maps.
//
// Replaces `convolute` (kernel.cl:2-60; launched at MobileNet.c:208-315).  Tap order is the
// reference's: for each filter R taps, G taps, B taps, row offset outer / column offset inner
// (kernel.cl:15-51).  Intended semantics (SURVEY App. C D-01/02/05 not reproduced): the
// accumulator restarts per filter, `stride` is a stride, borders are zero padded.
//
// Reads raw u8 (planar planes or the interleaved PPM payload), applies x*in_scale+in_bias,
// accumulates in fp32 and writes NHWC with the folded-BN scale/shift + ReLU/ReLU6 epilogue.
// Each CTA stages its input halo tile in shared memory once (coalesced byte loads) and keeps
// the 27x32 filter bank there; each thread produces 2 output pixels x 32 channels.
#include "common.cuh"

namespace mnv1 {

constexpr int STEM_C = 32;    // op_size of layer 1 (MobileNet.c:123, FILTER_SIZE_L1)
constexpr int STEM_TX = 32;   // threads along x; each does pixels tx and tx+32
constexpr int STEM_TY = 4;    // output rows per CTA
constexpr int STEM_OW = 64;   // output columns per CTA

template <typename T, int S>
__global__ void __launch_bounds__(STEM_TX * STEM_TY) stem_kernel(T* __restrict__ out, StemArgs a,
                                                                 const float* __restrict__ w27xC,
                                                                 Epilogue ep) {
  constexpr int IN_ROWS = (STEM_TY - 1) * S + 3;
  constexpr int IN_COLS = (STEM_OW - 1) * S + 3;
  constexpr int IN_PITCH = IN_COLS + 1;
  __shared__ float s_in[3][IN_ROWS][IN_PITCH];
  __shared__ __align__(16) float s_w[27 * STEM_C];

  const int orows = a.rows / S, ocols = a.cols / S;
  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * STEM_TY, ox0 = blockIdx.x * STEM_OW;
  const int iy0 = oy0 * S - a.pad_lo, ix0 = ox0 * S - a.pad_lo;
  const int tid = threadIdx.y * STEM_TX + threadIdx.x;

  for (int i = tid; i < 27 * STEM_C; i += STEM_TX * STEM_TY) s_w[i] = w27xC[i];
  const uint8_t* planes[3] = {a.r + (long)img * a.img_stride, a.g + (long)img * a.img_stride,
                              a.b + (long)img * a.img_stride};
  for (int i = tid; i < 3 * IN_ROWS * IN_COLS; i += STEM_TX * STEM_TY) {
    int p = i / (IN_ROWS * IN_COLS), rem = i % (IN_ROWS * IN_COLS);
    int ry = rem / IN_COLS, rx = rem % IN_COLS;
    int y = iy0 + ry, x = ix0 + rx;
    float v = 0.f;  // zero padding is applied AFTER the input transform (pads are true zeros)
    if (y >= 0 && y < a.rows && x >= 0 && x < a.cols)
      v = fmaf((float)planes[p][((long)y * a.cols + x) * a.pix_stride], a.in_scale, a.in_bias);
    s_in[p][ry][rx] = v;
  }
  __syncthreads();

  float acc0[STEM_C], acc1[STEM_C];
#pragma unroll
  for (int c = 0; c < STEM_C; ++c) acc0[c] = acc1[c] = 0.f;
  const int ly = threadIdx.y * S, lx0 = threadIdx.x * S, lx1 = (threadIdx.x + STEM_TX) * S;
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float v0 = s_in[p][ly + i][lx0 + j], v1 = s_in[p][ly + i][lx1 + j];
        const float4* wv = reinterpret_cast<const float4*>(&s_w[(p * 9 + i * 3 + j) * STEM_C]);
#pragma unroll
        for (int q = 0; q < STEM_C / 4; ++q) {
          const float4 wq = wv[q];
          acc0[4 * q + 0] = fmaf(v0, wq.x, acc0[4 * q + 0]);
          acc0[4 * q + 1] = fmaf(v0, wq.y, acc0[4 * q + 1]);
          acc0[4 * q + 2] = fmaf(v0, wq.z, acc0[4 * q + 2]);
          acc0[4 * q + 3] = fmaf(v0, wq.w, acc0[4 * q + 3]);
          acc1[4 * q + 0] = fmaf(v1, wq.x, acc1[4 * q + 0]);
          acc1[4 * q + 1] = fmaf(v1, wq.y, acc1[4 * q + 1]);
          acc1[4 * q + 2] = fmaf(v1, wq.z, acc1[4 * q + 2]);
          acc1[4 * q + 3] = fmaf(v1, wq.w, acc1[4 * q + 3]);
        }
      }

  const int oy = oy0 + threadIdx.y;
  if (oy >= orows) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int ox = ox0 + threadIdx.x + half * STEM_TX;
    if (ox >= ocols) continue;
    float* acc = half ? acc1 : acc0;
#pragma unroll
    for (int c = 0; c < STEM_C; ++c) {
      const float s = ep.scale ? __ldg(ep.scale + c) : 1.f, t = ep.shift ? __ldg(ep.shift + c) : 0.f;
      acc[c] = apply_epilogue(acc[c], s, t, ep.act);
    }
    T* o = out + (((long)img * orows + oy) * ocols + ox) * STEM_C;
    if constexpr (sizeof(T) == 2) {
      uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
      for (int q = 0; q < STEM_C / 8; ++q)
        o4[q] = make_uint4(pack_bf16x2(acc[8 * q], acc[8 * q + 1]), pack_bf16x2(acc[8 * q + 2], acc[8 * q + 3]),
                           pack_bf16x2(acc[8 * q + 4], acc[8 * q + 5]), pack_bf16x2(acc[8 * q + 6], acc[8 * q + 7]));
    } else {
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int q = 0; q < STEM_C / 4; ++q) o4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
  }
}

cudaError_t launch_stem(mnv1_dtype dt, void* out, const StemArgs& a, const float* w27xC, Epilogue ep,
                        cudaStream_t st) {
  if (a.cout != STEM_C || (a.stride != 1 && a.stride != 2)) return cudaErrorInvalidValue;
  if (a.n <= 0) return cudaSuccess;
  const int orows = a.rows / a.stride, ocols = a.cols / a.stride;
  dim3 grid((ocols + STEM_OW - 1) / STEM_OW, (orows + STEM_TY - 1) / STEM_TY, a.n), block(STEM_TX, STEM_TY);
  if (dt == MNV1_BF16) {
    if (a.stride == 2) stem_kernel<bf16, 2><<<grid, block, 0, st>>>((bf16*)out, a, w27xC, ep);
    else stem_kernel<bf16, 1><<<grid, block, 0, st>>>((bf16*)out, a, w27xC, ep);
  } else {
    if (a.stride == 2) stem_kernel<float, 2><<<grid, block, 0, st>>>((float*)out, a, w27xC, ep);
    else stem_kernel<float, 1><<<grid, block, 0, st>>>((float*)out, a, w27xC, ep);
  }
  return cudaGetLastError();
}

}  // namespace mnv1
