// pointwise_simt.cu — 1x1 convolution / FC as a CUDA-core GEMM with fp32 accumulation.
//
// Replaces `pointwise` (kernel.cl:94-114) for fp32 contexts (BASELINE configs 1-3, where the
// 1e-4 per-layer bar rules out bf16/tf32 tensor-core inputs) and the FC layer
// (MobileNet.c:2681-2763, pointwise at rows=cols=1).  bf16 contexts use pointwise_tc.cu.
//   out[m][co] = act( scale[co] * sum_k in[m][k] * w[co][k] + shift[co] )
// in = NHWC feature map viewed as [M = N*H*W][K = Cin]; w in the reference's [Cout][Cin]
// `findex` order (kernel.cl:106).  64x64 tile, 16-deep k-slices, 4x4 outputs per thread.
#include "common.cuh"

namespace mnv1 {

constexpr int PS_BM = 64, PS_BN = 64, PS_BK = 16;

template <typename TA, typename TW, typename TO>
__global__ void __launch_bounds__(256) pointwise_simt_kernel(TO* __restrict__ out, const TA* __restrict__ in,
                                                             const TW* __restrict__ w, long M, int K, int Cout,
                                                             Epilogue ep) {
  __shared__ float sA[PS_BK][PS_BM + 4];
  __shared__ float sB[PS_BK][PS_BN + 4];
  const long m0 = (long)blockIdx.x * PS_BM;
  const int n0 = blockIdx.y * PS_BN;
  const int tid = threadIdx.x;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  // loader mapping: 256 threads fetch a 64 x 16 slice, k fastest (contiguous in memory)
  const int lr = tid / 4, lk = (tid % 4) * 4;
  // register-prefetched k loop: the global loads of slice k0+BK are in flight while slice k0 is
  // multiplied out of shared memory
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + lk + q;
      const long m = m0 + lr;
      ra[q] = (m < M && k < K) ? to_f32<TA>(in[m * K + k]) : 0.f;
      const int co = n0 + lr;
      rb[q] = (co < Cout && k < K) ? to_f32<TW>(w[(long)co * K + k]) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += PS_BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { sA[lk + q][lr] = ra[q]; sB[lk + q][lr] = rb[q]; }
    __syncthreads();
    if (k0 + PS_BK < K) fetch(k0 + PS_BK);
#pragma unroll
    for (int kk = 0; kk < PS_BK; ++kk) {
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

cudaError_t launch_pointwise_simt(mnv1_dtype dt, void* out, const void* in, const float* w_f32,
                                  const bf16* w_bf16, long m, int k, int cout, Epilogue ep, bool out_f32,
                                  cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  dim3 grid((unsigned)((m + PS_BM - 1) / PS_BM), (cout + PS_BN - 1) / PS_BN), block(256);
  if (dt == MNV1_F32) {
    pointwise_simt_kernel<float, float, float><<<grid, block, 0, st>>>((float*)out, (const float*)in, w_f32, m, k, cout, ep);
  } else if (out_f32) {
    pointwise_simt_kernel<bf16, bf16, float><<<grid, block, 0, st>>>((float*)out, (const bf16*)in, w_bf16, m, k, cout, ep);
  } else {
    pointwise_simt_kernel<bf16, bf16, bf16><<<grid, block, 0, st>>>((bf16*)out, (const bf16*)in, w_bf16, m, k, cout, ep);
  }
  return cudaGetLastError();
}

// explicit instantiations used by head.cu (fp32 pooled activations x {fp32, bf16} FC weights)
template __global__ void pointwise_simt_kernel<float, bf16, float>(float*, const float*, const bf16*, long, int, int, Epilogue);

cudaError_t launch_fc_f32in(float* out, const float* in, const float* w_f32, const bf16* w_bf16, long m, int k,
                            int cout, Epilogue ep, cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  dim3 grid((unsigned)((m + PS_BM - 1) / PS_BM), (cout + PS_BN - 1) / PS_BN), block(256);
  if (w_bf16) pointwise_simt_kernel<float, bf16, float><<<grid, block, 0, st>>>(out, in, w_bf16, m, k, cout, ep);
  else pointwise_simt_kernel<float, float, float><<<grid, block, 0, st>>>(out, in, w_f32, m, k, cout, ep);
  return cudaGetLastError();
}

}  // namespace mnv1
