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
// A CTA holds the pooled vectors of 8 images in shared memory (fp32) and owns 64 classes; a warp
// owns 8 filter rows at once.  Per 256-wide k step a lane loads 8 filter values per row (kept in
// registers, 64 in all) and streams the 8 images past them: 2 conflict-free LDS.128 feed 64 FFMA.
// A warp-shuffle tree folds the 32 per-lane partial sums of the 8 x 8 outputs at the end.
constexpr int FC_IMGS = 8, FC_CLS = 64, FC_WCLS = 8;
template <typename TW>
__global__ void __launch_bounds__(256, 1) fc_kernel(float* __restrict__ out, const float* __restrict__ pooled,
                                                    const TW* __restrict__ w, const float* __restrict__ bias, int n,
                                                    int k, int classes) {
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
  const int clsw = cls0 + warp * FC_WCLS;
  float acc[FC_WCLS][FC_IMGS];
#pragma unroll
  for (int c = 0; c < FC_WCLS; ++c)
#pragma unroll
    for (int im = 0; im < FC_IMGS; ++im) acc[c][im] = 0.f;
  for (int k0 = 0; k0 < k; k0 += 256) {
    // this lane's 8 contraction indices of the step: k0 + lane*4 + {0..3} and + 128
    const int ka = k0 + lane * 4, kb = ka + 128;
    float wv[FC_WCLS][8];
#pragma unroll
    for (int c = 0; c < FC_WCLS; ++c) {
      const int cls = min(clsw + c, classes - 1);  // clamp: rows past the end are computed but never stored
      if constexpr (sizeof(TW) == 2) {
        const uint2 r0 = __ldg(reinterpret_cast<const uint2*>(w + (long)cls * k + ka));
        const uint2 r1 = __ldg(reinterpret_cast<const uint2*>(w + (long)cls * k + kb));
        wv[c][0] = bf16lo_to_f32(r0.x); wv[c][1] = bf16hi_to_f32(r0.x); wv[c][2] = bf16lo_to_f32(r0.y); wv[c][3] = bf16hi_to_f32(r0.y);
        wv[c][4] = bf16lo_to_f32(r1.x); wv[c][5] = bf16hi_to_f32(r1.x); wv[c][6] = bf16lo_to_f32(r1.y); wv[c][7] = bf16hi_to_f32(r1.y);
      } else {
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(w + (long)cls * k + ka));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(w + (long)cls * k + kb));
        wv[c][0] = r0.x; wv[c][1] = r0.y; wv[c][2] = r0.z; wv[c][3] = r0.w;
        wv[c][4] = r1.x; wv[c][5] = r1.y; wv[c][6] = r1.z; wv[c][7] = r1.w;
      }
    }
#pragma unroll
    for (int im = 0; im < FC_IMGS; ++im) {
      const float4 a0 = *reinterpret_cast<const float4*>(&s_a[im * k + ka]);
      const float4 a1 = *reinterpret_cast<const float4*>(&s_a[im * k + kb]);
#pragma unroll
      for (int c = 0; c < FC_WCLS; ++c) {
        float t = acc[c][im];
        t = fmaf(a0.x, wv[c][0], t); t = fmaf(a0.y, wv[c][1], t); t = fmaf(a0.z, wv[c][2], t); t = fmaf(a0.w, wv[c][3], t);
        t = fmaf(a1.x, wv[c][4], t); t = fmaf(a1.y, wv[c][5], t); t = fmaf(a1.z, wv[c][6], t); t = fmaf(a1.w, wv[c][7], t);
        acc[c][im] = t;
      }
    }
  }
  // fold the 32 lanes; lane (c*8 + im) ends up owning output (class c, image im)
  float mine = 0.f, mine2 = 0.f;
#pragma unroll
  for (int c = 0; c < FC_WCLS; ++c)
#pragma unroll
    for (int im = 0; im < FC_IMGS; ++im) {
      float v = acc[c][im];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == ((c * FC_IMGS + im) & 31)) {
        if (c * FC_IMGS + im < 32) mine = v; else mine2 = v;
      }
    }
  // lane L now holds outputs L (mine) and 32 + L (mine2)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int idx = h * 32 + lane, c = idx / FC_IMGS, im = idx % FC_IMGS;
    const int cls = clsw + c;
    const float v = h == 0 ? mine : mine2;
    if (cls < classes && img0 + im < n) out[(long)(img0 + im) * classes + cls] = v + (bias ? __ldg(bias + cls) : 0.f);
  }
}

cudaError_t launch_fc(float* out, const float* pooled, const float* w_f32, const bf16* w_bf16, const float* bias, int n,
                      int k, int classes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (k % 256) return cudaErrorNotSupported;
  dim3 grid((classes + FC_CLS - 1) / FC_CLS, (n + FC_IMGS - 1) / FC_IMGS);
  const size_t smem = (size_t)FC_IMGS * k * sizeof(float);
  if (k % 256 || smem > 96 * 1024) return cudaErrorNotSupported;
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
