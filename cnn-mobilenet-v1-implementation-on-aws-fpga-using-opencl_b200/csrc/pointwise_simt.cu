// pointwise_simt.cu — 1x1 convolution / FC as a CUDA-core GEMM with fp32 accumulation.
//
// Replaces `pointwise` (kernel.cl:94-114) for fp32 contexts (BASELINE configs 1-3, where the
// 1e-4 per-layer bar rules out bf16/tf32 tensor-core inputs) and the FC layer
// (MobileNet.c:2681-2763, pointwise at rows=cols=1).  bf16 contexts use pointwise_tc.cu.
//   out[m][co] = act( scale[co] * sum_k in[m][k] * w[co][k] + shift[co] )
// in = NHWC feature map viewed as [M = N*H*W][K = Cin]; w in the reference's [Cout][Cin]
// `findex` order (kernel.cl:106).  64x64 tile, 64-deep k-slices (16 for narrow contractions), 4x4 outputs per thread.
#include <type_traits>

#include "common.cuh"

namespace mnv1 {

constexpr int PS_BM = 64, PS_BN = 64;

// BK: depth of a k-slice.  16 for narrow contractions; 64 otherwise — at batch 1 (BASELINE configs 1-3) a layer is a few
// dozen CTAs walking down K, every slice costs one exposed global-load round trip (~1 us) whatever its depth, and 16-deep
// slices made a 512 -> 512 layer 32 round trips long (35 us; 8 us with 64-deep slices, each thread keeping eight 16-byte
// loads in flight).
template <typename TA, typename TW, typename TO, int BK>
__global__ void __launch_bounds__(256) pointwise_simt_kernel(TO* __restrict__ out, const TA* __restrict__ in,
                                                             const TW* __restrict__ w, long M, int K, int Cout,
                                                             Epilogue ep) {
  __shared__ __align__(16) float sA[BK][PS_BM + 4];
  __shared__ __align__(16) float sB[BK][PS_BN + 4];
  const long m0 = (long)blockIdx.x * PS_BM;
  const int n0 = blockIdx.y * PS_BN;
  const int tid = threadIdx.x;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  // loader mapping: 256 threads fetch a 64 x BK slice of each operand, k fastest (contiguous in memory): thread = row
  // tid / 4, k = 4 (tid % 4) + 16 j + q  (j < BK / 16, q < 4: one 16-byte load per j)
  constexpr int NJ = BK / 16;
  const int lr = tid / 4, lk = (tid % 4) * 4;
  float ra[NJ][4], rb[NJ][4];
  const bool vec_ok = (K & 3) == 0;
  auto fetch4 = [&](const auto* base, long row, long rows, int k, float (&dst)[4]) {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(base)>>;
    if (row < rows && k + 3 < K && vec_ok) {
      if constexpr (sizeof(T) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + row * K + k));
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(base + row * K + k));
        dst[0] = __uint_as_float(v.x << 16); dst[1] = __uint_as_float(v.x & 0xffff0000u);
        dst[2] = __uint_as_float(v.y << 16); dst[3] = __uint_as_float(v.y & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = (row < rows && k + q < K) ? to_f32<T>(base[row * K + k + q]) : 0.f;
    }
  };
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      fetch4(in, m0 + lr, M, k0 + lk + 16 * j, ra[j]);
      fetch4(w, (long)(n0 + lr), (long)Cout, k0 + lk + 16 * j, rb[j]);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) { sA[lk + 16 * j + q][lr] = ra[j][q]; sB[lk + 16 * j + q][lr] = rb[j][q]; }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[kk][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[kk][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tn + j;
      if (co >= Cout) continue;
      const float s = ep.scale ? __ldg(ep.scale + co) : 1.f, t = ep.shift ? __ldg(ep.shift + co) : 0.f;
      out[m * Cout + co] = from_f32<TO>(apply_epilogue(acc[i][j], s, t, ep.act));
    }
  }
}

template <typename TA, typename TW, typename TO>
static void launch_ps(TO* out, const TA* in, const TW* w, long m, int k, int cout, Epilogue ep, cudaStream_t st) {
  dim3 grid((unsigned)((m + PS_BM - 1) / PS_BM), (cout + PS_BN - 1) / PS_BN), block(256);
  if (k >= 64) pointwise_simt_kernel<TA, TW, TO, 64><<<grid, block, 0, st>>>(out, in, w, m, k, cout, ep);
  else pointwise_simt_kernel<TA, TW, TO, 16><<<grid, block, 0, st>>>(out, in, w, m, k, cout, ep);
}

cudaError_t launch_pointwise_simt(mnv1_dtype dt, void* out, const void* in, const float* w_f32,
                                  const bf16* w_bf16, long m, int k, int cout, Epilogue ep, bool out_f32,
                                  cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  if (dt == MNV1_F32) launch_ps((float*)out, (const float*)in, w_f32, m, k, cout, ep, st);
  else if (out_f32) launch_ps((float*)out, (const bf16*)in, w_bf16, m, k, cout, ep, st);
  else launch_ps((bf16*)out, (const bf16*)in, w_bf16, m, k, cout, ep, st);
  return cudaGetLastError();
}

cudaError_t launch_fc_f32in(float* out, const float* in, const float* w_f32, const bf16* w_bf16, long m, int k,
                            int cout, Epilogue ep, cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  if (w_bf16) launch_ps(out, in, w_bf16, m, k, cout, ep, st);
  else launch_ps(out, in, w_f32, m, k, cout, ep, st);
  return cudaGetLastError();
}

}  // namespace mnv1
