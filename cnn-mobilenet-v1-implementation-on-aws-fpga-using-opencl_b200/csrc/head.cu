// head.cu — layers 28-30: global average pool, FC, softmax + argmax.
//
// Replaces `pool` (kernel.cl:116-131; MobileNet.c:2601-2679), the FC launch of `pointwise`
// at rows=cols=1 (MobileNet.c:2681-2763) and the host softmax/argmax loop
// (MobileNet.c:2769-2792).  Intended semantics: one mean per channel (App. C D-12), bias and
// no activation on the logits (D-13), softmax with max subtraction, first maximum wins
// (strict `>` at MobileNet.c:2786), 0-based index (the host prints index+1).
#include "common.cuh"

namespace mnv1 {

cudaError_t launch_fc_f32in(float* out, const float* in, const float* w_f32, const bf16* w_bf16, long m, int k,
                            int cout, Epilogue ep, cudaStream_t st);

// in: NHWC [n][hw][c].  A CTA reduces 256 channels of one image: 64 channel quads x 4 pixel
// phases (pixels p = phase, phase+4, ...) so the 49 row reads of a quad are spread over 4 threads
// with independent loads in flight, then a shared-memory fold of the 4 partial sums.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) pool_kernel(TO* __restrict__ out, const T* __restrict__ in, int n, int hw,
                                                   int c) {
  __shared__ float part[4][64][4];
  const int img = blockIdx.x, q = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int c0 = blockIdx.y * 256 + q * 4;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c0 < c) {
    const T* p = in + (long)img * hw * c + c0;
#pragma unroll 4
    for (int i = ph; i < hw; i += 4) {
      if constexpr (sizeof(T) == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p + (long)i * c));
        s0 += bf16lo_to_f32(v.x); s1 += bf16hi_to_f32(v.x); s2 += bf16lo_to_f32(v.y); s3 += bf16hi_to_f32(v.y);
      } else {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + (long)i * c));
        s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
      }
    }
  }
  part[ph][q][0] = s0; part[ph][q][1] = s1; part[ph][q][2] = s2; part[ph][q][3] = s3;
  __syncthreads();
  if (ph == 0 && c0 < c) {
    const float inv = 1.0f / (float)hw;
    TO* o = out + (long)img * c + c0;
#pragma unroll
    for (int v = 0; v < 4; ++v)
      o[v] = from_f32<TO>((part[0][q][v] + part[1][q][v] + part[2][q][v] + part[3][q][v]) * inv);
  }
}

cudaError_t launch_pool(mnv1_dtype dt, void* out, const void* in, int n, int hw, int c, bool out_f32,
                        cudaStream_t st) {
  if (c % 4) return cudaErrorInvalidValue;
  if (n <= 0) return cudaSuccess;
  dim3 grid(n, (c + 255) / 256);
  if (dt == MNV1_F32) pool_kernel<float, float><<<grid, 256, 0, st>>>((float*)out, (const float*)in, n, hw, c);
  else if (out_f32)   pool_kernel<bf16, float><<<grid, 256, 0, st>>>((float*)out, (const bf16*)in, n, hw, c);
  else                pool_kernel<bf16, bf16><<<grid, 256, 0, st>>>((bf16*)out, (const bf16*)in, n, hw, c);
  return cudaGetLastError();
}

// FC (MobileNet.c:2681-2763): logits[img][cls] = bias[cls] + sum_k pooled[img][k] * w[cls][k].
// A CTA holds the pooled vectors of 16 images in shared memory (fp32) and owns 64 classes; each
// warp walks 8 filter rows: lanes stride the contraction (coalesced 16-byte filter loads, each
// filter value reused for the 16 images), then a warp-shuffle tree folds the 32 partial sums.
constexpr int FC_IMGS = 16, FC_CLS = 64;
template <typename TW>
__global__ void __launch_bounds__(256) fc_kernel(float* __restrict__ out, const float* __restrict__ pooled,
                                                 const TW* __restrict__ w, const float* __restrict__ bias, int n, int k,
                                                 int classes) {
  extern __shared__ float s_a[];  // [FC_IMGS][k]
  const int img0 = blockIdx.y * FC_IMGS, cls0 = blockIdx.x * FC_CLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 4; i < FC_IMGS * k; i += 256 * 4) {
    const int im = i / k;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (img0 + im < n) v = __ldg(reinterpret_cast<const float4*>(pooled + (long)(img0 + im) * k + (i - im * k)));
    *reinterpret_cast<float4*>(&s_a[i]) = v;
  }
  __syncthreads();
  for (int c = 0; c < FC_CLS / 8; ++c) {
    const int cls = cls0 + warp * (FC_CLS / 8) + c;
    if (cls >= classes) break;
    float acc[FC_IMGS];
#pragma unroll
    for (int im = 0; im < FC_IMGS; ++im) acc[im] = 0.f;
    for (int kk = lane * 8; kk < k; kk += 256) {
      float wv[8];
      if constexpr (sizeof(TW) == 2) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(w + (long)cls * k + kk));
        wv[0] = bf16lo_to_f32(raw.x); wv[1] = bf16hi_to_f32(raw.x); wv[2] = bf16lo_to_f32(raw.y); wv[3] = bf16hi_to_f32(raw.y);
        wv[4] = bf16lo_to_f32(raw.z); wv[5] = bf16hi_to_f32(raw.z); wv[6] = bf16lo_to_f32(raw.w); wv[7] = bf16hi_to_f32(raw.w);
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w + (long)cls * k + kk));
        const float4 b = __ldg(reinterpret_cast<const float4*>(w + (long)cls * k + kk + 4));
        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
      }
#pragma unroll
      for (int im = 0; im < FC_IMGS; ++im) {
        const float4 a0 = *reinterpret_cast<const float4*>(&s_a[im * k + kk]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s_a[im * k + kk + 4]);
        acc[im] = fmaf(a0.x, wv[0], fmaf(a0.y, wv[1], fmaf(a0.z, wv[2], fmaf(a0.w, wv[3], acc[im]))));
        acc[im] = fmaf(a1.x, wv[4], fmaf(a1.y, wv[5], fmaf(a1.z, wv[6], fmaf(a1.w, wv[7], acc[im]))));
      }
    }
#pragma unroll
    for (int im = 0; im < FC_IMGS; ++im) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[im] += __shfl_xor_sync(0xffffffffu, acc[im], off);
    }
    if (lane < FC_IMGS && img0 + lane < n) {
      float v = acc[0];
#pragma unroll
      for (int im = 1; im < FC_IMGS; ++im) v = lane == im ? acc[im] : v;
      out[(long)(img0 + lane) * classes + cls] = v + (bias ? __ldg(bias + cls) : 0.f);
    }
  }
}

cudaError_t launch_fc(float* out, const float* pooled, const float* w_f32, const bf16* w_bf16, const float* bias, int n,
                      int k, int classes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (k % 256) return cudaErrorNotSupported;
  dim3 grid((classes + FC_CLS - 1) / FC_CLS, (n + FC_IMGS - 1) / FC_IMGS);
  const size_t smem = (size_t)FC_IMGS * k * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fc_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (smem > 96 * 1024) return cudaErrorNotSupported;
  if (w_bf16) fc_kernel<bf16><<<grid, 256, smem, st>>>(out, pooled, w_bf16, bias, n, k, classes);
  else fc_kernel<float><<<grid, 256, smem, st>>>(out, pooled, w_f32, bias, n, k, classes);
  return cudaGetLastError();
}

// one warp per image: max / argmax, sum of exp, optional probabilities — warp shuffles only.
__global__ void __launch_bounds__(128) softmax_kernel(const float* __restrict__ logits, int n, int classes,
                                                      float* __restrict__ prob, int* __restrict__ top1,
                                                      float* __restrict__ top1_prob) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* z = logits + (long)warp * classes;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int k = lane; k < classes; k += 32) {
    const float v = z[k];
    if (v > mx) { mx = v; arg = k; }  // per lane indices increase, so strict '>' keeps the first
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, off);
    const int oarg = __shfl_xor_sync(0xffffffffu, arg, off);
    if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
  }
  float sum = 0.f;
  for (int k = lane; k < classes; k += 32) sum += __expf(z[k] - mx);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float inv = 1.0f / sum;
  if (prob)
    for (int k = lane; k < classes; k += 32) prob[(long)warp * classes + k] = __expf(z[k] - mx) * inv;
  if (lane == 0) {
    if (top1) top1[warp] = arg;
    if (top1_prob) top1_prob[warp] = inv;
  }
}

cudaError_t launch_softmax(const float* logits, int n, int classes, float* prob, int* top1, float* top1_prob,
                           cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n + 3) / 4);
  softmax_kernel<<<grid, 128, 0, st>>>(logits, n, classes, prob, top1, top1_prob);
  return cudaGetLastError();
}

cudaError_t launch_head(mnv1_dtype dt, const void* in, int n, int hw, int c, const mnv1_filter* fc,
                        float* pooled_scratch, float* logits, int* top1, float* top1_prob, int classes,
                        cudaStream_t st, int* launches) {
  cudaError_t e = launch_pool(dt, pooled_scratch, in, n, hw, c, /*out_f32=*/true, st);
  if (e != cudaSuccess) return e;
  e = launch_fc(logits, pooled_scratch, fc->w_f32, dt == MNV1_BF16 ? fc->w_bf16 : nullptr, fc->shift, n, c, classes, st);
  if (e == cudaErrorNotSupported) {
    Epilogue ep{nullptr, fc->shift, MNV1_ACT_NONE};
    e = launch_fc_f32in(logits, pooled_scratch, fc->w_f32, dt == MNV1_BF16 ? fc->w_bf16 : nullptr, n, c, classes, ep, st);
  }
  if (e != cudaSuccess) return e;
  if (launches) *launches = 2;
  if (top1 || top1_prob) {
    e = launch_softmax(logits, n, classes, nullptr, top1, top1_prob, st);
    if (launches) *launches = 3;
  }
  return e;
}

}  // namespace mnv1
