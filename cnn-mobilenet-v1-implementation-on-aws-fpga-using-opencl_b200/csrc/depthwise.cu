// depthwise.cu — per-channel 3x3 stencil, stride 1 or 2, fused scale/shift + ReLU/ReLU6.
//
// Replaces `depthwise` (kernel.cl:62-92; 13 launches).  Tap order i (row
// offset) outer, j (column offset) inner (kernel.cl:75-77).  Intended semantics: every
// output channel reads ITS input channel (App. C D-03), `stride` is a stride (D-02), zero
// padding on all borders (D-08), accumulator restarts per channel (D-01).
//
// HBM-bound (AI 1.8-4.5 FLOP/B): the layout is NHWC so a thread owns one 128-bit channel
// vector (8 bf16 / 4 fp32) of one output column and walks DW_R output rows with a rolling
// 3-row register window; lanes of a warp cover consecutive channel vectors and then
// consecutive pixels, i.e. contiguous memory.  The 9 x VEC filter taps live in registers.
#include "common.cuh"

namespace mnv1 {

constexpr int DW_R = 7;  // output rows per thread: divides 112, 56, 28, 14 and 7

template <typename T> struct Vec;
template <> struct Vec<bf16> { static constexpr int N = 8; };
template <> struct Vec<float> { static constexpr int N = 4; };

template <typename T>
__device__ __forceinline__ void unpack(const uint4& raw, float (&f)[Vec<T>::N]) {
  if constexpr (sizeof(T) == 2) {
    f[0] = bf16lo_to_f32(raw.x); f[1] = bf16hi_to_f32(raw.x);
    f[2] = bf16lo_to_f32(raw.y); f[3] = bf16hi_to_f32(raw.y);
    f[4] = bf16lo_to_f32(raw.z); f[5] = bf16hi_to_f32(raw.z);
    f[6] = bf16lo_to_f32(raw.w); f[7] = bf16hi_to_f32(raw.w);
  } else {
    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
    f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack(const float (&f)[Vec<T>::N]) {
  if constexpr (sizeof(T) == 2)
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                      pack_bf16x2(f[6], f[7]));
  else
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
}

template <typename T, int S>
__global__ void __launch_bounds__(128) depthwise_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                        const float* __restrict__ w9xC, Epilogue ep, int n,
                                                        int H, int W, int C, int Ho, int Wo, int pad_lo,
                                                        int strips, long total) {
  constexpr int V = Vec<T>::N;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cvn = C / V;
  const int cv = (int)(idx % cvn);
  long t = idx / cvn;
  const int ox = (int)(t % Wo); t /= Wo;
  const int strip = (int)(t % strips);
  const int img = (int)(t / strips);
  const int c0 = cv * V;

  float wt[9][V];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if constexpr (V == 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(w9xC + (long)k * C + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(w9xC + (long)k * C + c0 + 4));
      wt[k][0] = a.x; wt[k][1] = a.y; wt[k][2] = a.z; wt[k][3] = a.w;
      wt[k][4] = b.x; wt[k][5] = b.y; wt[k][6] = b.z; wt[k][7] = b.w;
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(w9xC + (long)k * C + c0));
      wt[k][0] = a.x; wt[k][1] = a.y; wt[k][2] = a.z; wt[k][3] = a.w;
    }
  }
  float sc[V], sh[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    sc[v] = ep.scale ? __ldg(ep.scale + c0 + v) : 1.f;
    sh[v] = ep.shift ? __ldg(ep.shift + c0 + v) : 0.f;
  }

  const T* in_img = in + (long)img * H * W * C + c0;
  const int ix0 = ox * S - pad_lo;
  const int oy0 = strip * DW_R;
  const int iy0 = oy0 * S - pad_lo;
  const bool xok[3] = {ix0 >= 0 && ix0 < W, ix0 + 1 >= 0 && ix0 + 1 < W, ix0 + 2 >= 0 && ix0 + 2 < W};

  auto load_row = [&](int iy, uint4 (&row)[3]) {
    const bool yok = iy >= 0 && iy < H;
    const T* p = in_img + ((long)iy * W + ix0) * C;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      row[j] = (yok && xok[j]) ? __ldg(reinterpret_cast<const uint4*>(p + (long)j * C)) : make_uint4(0, 0, 0, 0);
  };

  // rolling window: rows[q % 3] holds input row iy0 + q
  uint4 rows[3][3];
  if (S == 1) { load_row(iy0, rows[0]); load_row(iy0 + 1, rows[1]); }
  else        { load_row(iy0, rows[0]); }

#pragma unroll
  for (int r = 0; r < DW_R; ++r) {
    const int oy = oy0 + r;
    if (oy >= Ho) break;
    // input rows q = r*S + {0,1,2}
    if (S == 1) { load_row(iy0 + r + 2, rows[(r + 2) % 3]); }
    else        { load_row(iy0 + 2 * r + 1, rows[(2 * r + 1) % 3]); load_row(iy0 + 2 * r + 2, rows[(2 * r + 2) % 3]); }
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float x[V];
        unpack<T>(rows[(r * S + i) % 3][j], x);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = fmaf(x[v], wt[i * 3 + j][v], acc[v]);
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = apply_epilogue(acc[v], sc[v], sh[v], ep.act);
    *reinterpret_cast<uint4*>(out + (((long)img * Ho + oy) * Wo + ox) * C + c0) = pack<T>(acc);
  }
}

cudaError_t launch_depthwise(mnv1_dtype dt, void* out, const void* in, const float* w9xC, int n, int rows,
                             int cols, int stride, int c, int pad_lo, Epilogue ep, cudaStream_t st) {
  const int V = dt == MNV1_BF16 ? 8 : 4;
  if ((stride != 1 && stride != 2) || c % V) return cudaErrorInvalidValue;
  if (n <= 0) return cudaSuccess;
  const int Ho = rows / stride, Wo = cols / stride;
  const int strips = (Ho + DW_R - 1) / DW_R;
  const long total = (long)n * strips * Wo * (c / V);
  const unsigned grid = (unsigned)((total + 127) / 128);
#define DW_LAUNCH(T, S) \
  depthwise_kernel<T, S><<<grid, 128, 0, st>>>((T*)out, (const T*)in, w9xC, ep, n, rows, cols, c, Ho, Wo, pad_lo, strips, total)
  if (dt == MNV1_BF16) { if (stride == 1) DW_LAUNCH(bf16, 1); else DW_LAUNCH(bf16, 2); }
  else                 { if (stride == 1) DW_LAUNCH(float, 1); else DW_LAUNCH(float, 2); }
#undef DW_LAUNCH
  return cudaGetLastError();
}

}  // namespace mnv1
