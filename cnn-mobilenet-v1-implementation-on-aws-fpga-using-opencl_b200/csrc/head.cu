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
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.x, q = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int c0 = blockIdx.y * 256 + q * 4;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c0 < c) {
    const T* p = in + (long)img * hw * c + c0;
#pragma unroll 13   // hw = 49: 12 or 13 rows per phase, all of a thread's loads in flight at once (same summation order)
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
  else if (out_f32)   return launch_pdl(pool_kernel<bf16, float>, grid, dim3(256), 0, st, (float*)out, (const bf16*)in, n, hw, c);
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
  // The filter rows stream straight from L2; the 16 loads of step k0+256 are issued before the 512
  // FFMAs of step k0 (register double buffer), so their latency hides under the arithmetic.
  auto load_w = [&](int k0, uint2 (&r)[FC_WCLS][2], float4 (&f)[FC_WCLS][2]) {
    const int ka = k0 + lane * 4, kb = ka + 128;   // this lane's 8 contraction indices of the step
#pragma unroll
    for (int c = 0; c < FC_WCLS; ++c) {
      const int cls = min(clsw + c, classes - 1);  // clamp: rows past the end are computed but never stored
      if constexpr (sizeof(TW) == 2) {
        r[c][0] = __ldg(reinterpret_cast<const uint2*>(w + (long)cls * k + ka));
        r[c][1] = __ldg(reinterpret_cast<const uint2*>(w + (long)cls * k + kb));
      } else {
        f[c][0] = __ldg(reinterpret_cast<const float4*>(w + (long)cls * k + ka));
        f[c][1] = __ldg(reinterpret_cast<const float4*>(w + (long)cls * k + kb));
      }
    }
  };
  uint2 rn[FC_WCLS][2];
  float4 fn[FC_WCLS][2];
  load_w(0, rn, fn);
  for (int k0 = 0; k0 < k; k0 += 256) {
    const int ka = k0 + lane * 4, kb = ka + 128;
    float wv[FC_WCLS][8];
#pragma unroll
    for (int c = 0; c < FC_WCLS; ++c) {
      if constexpr (sizeof(TW) == 2) {
        const uint2 r0 = rn[c][0], r1 = rn[c][1];
        wv[c][0] = bf16lo_to_f32(r0.x); wv[c][1] = bf16hi_to_f32(r0.x); wv[c][2] = bf16lo_to_f32(r0.y); wv[c][3] = bf16hi_to_f32(r0.y);
        wv[c][4] = bf16lo_to_f32(r1.x); wv[c][5] = bf16hi_to_f32(r1.x); wv[c][6] = bf16lo_to_f32(r1.y); wv[c][7] = bf16hi_to_f32(r1.y);
      } else {
        const float4 r0 = fn[c][0], r1 = fn[c][1];
        wv[c][0] = r0.x; wv[c][1] = r0.y; wv[c][2] = r0.z; wv[c][3] = r0.w;
        wv[c][4] = r1.x; wv[c][5] = r1.y; wv[c][6] = r1.z; wv[c][7] = r1.w;
      }
    }
    if (k0 + 256 < k) load_w(k0 + 256, rn, fn);
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

// FC for bf16 contexts on the tensor cores without giving up the fp32 pooled means: every fp32 activation
// is split exactly into three bf16 pieces (x = hi + mid + lo, 8 mantissa bits each, by truncation), the
// bf16 filter is used as is, and three mma.sync.m16n8k16 per filter fragment accumulate the exact
// bf16 x bf16 products in fp32.  A warp owns 16 images x 32 classes; fragments come straight from L2
// with 128-bit loads (the contraction index is permuted the same way for both operands, so each lane
// reads 8 consecutive k per 32-wide step) and the loads of step i+1 are issued before the MMAs of step i.
// The four warps of a CTA split the contraction (k/4 each, summed in a fixed order through shared memory):
// 512 CTAs of short dependent chains instead of 128 long ones — the kernel is latency-, not math-bound.
// The legacy mma.sync path is deliberate: M = 256 is two tcgen05 tiles — not enough CTAs to matter.
#ifndef MNV1_FCM_NT
#define MNV1_FCM_NT 4
#endif
#ifndef MNV1_FCM_WARPS
#define MNV1_FCM_WARPS 4
#endif
constexpr int FCM_NT = MNV1_FCM_NT;        // 8-class fragments per warp (a CTA owns 16 images x 32 classes)
constexpr int FCM_WARPS = MNV1_FCM_WARPS;     // warps per CTA = k-splits
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x0, x1 -> packed bf16 pairs of the three exact pieces
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
  const uint32_t h0 = u0 & 0xffff0000u, h1 = u1 & 0xffff0000u;
  const float r0 = x0 - __uint_as_float(h0), r1 = x1 - __uint_as_float(h1);          // exact
  const uint32_t m0 = __float_as_uint(r0) & 0xffff0000u, m1 = __float_as_uint(r1) & 0xffff0000u;
  const float s0 = r0 - __uint_as_float(m0), s1 = r1 - __uint_as_float(m1);          // exact, <= 8 significant bits
  hi = (h0 >> 16) | h1;
  mid = (m0 >> 16) | m1;
  lo = (__float_as_uint(s0) >> 16) | (__float_as_uint(s1) & 0xffff0000u);
}
// FCM_DEPTH k-steps of operand loads are kept in flight per warp.  Measured inside the graph (tools/prefix_times.py, head =
// pool + FC + softmax): depth 4 at 227 registers = two CTAs per SM, the 512 CTAs need a second, mostly empty wave: 21.5 us;
// depth 1 under a 128-register bound = four CTAs per SM, one wave: 19.4 us (depth 2 at 128 registers spills: 21.6 us).
#ifndef MNV1_FCM_DEPTH
#define MNV1_FCM_DEPTH 1
#endif
constexpr int FCM_DEPTH = MNV1_FCM_DEPTH;
#ifndef MNV1_FCM_MINB
#define MNV1_FCM_MINB 4
#endif
__global__ void __launch_bounds__(FCM_WARPS * 32, MNV1_FCM_MINB) fc_mma_kernel(float* __restrict__ out, const float* __restrict__ pooled,
                                                                const bf16* __restrict__ w, const float* __restrict__ bias,
                                                                int n, int k, int classes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __shared__ float s_part[FCM_WARPS][16][FCM_NT * 8 + 1];
  pdl_trigger();
  const int img0 = blockIdx.y * 16, cls0 = blockIdx.x * (FCM_NT * 8);
  const int kw = k / FCM_WARPS, k_lo = warp * kw, k_hi = k_lo + kw;   // this warp's share of the contraction
  const int r0 = min(img0 + g, n - 1), r1 = min(img0 + g + 8, n - 1);       // clamped rows are computed, never stored
  const float* a0p = pooled + (long)r0 * k + 8 * t;
  const float* a1p = pooled + (long)r1 * k + 8 * t;
  const bf16* bp[FCM_NT];
#pragma unroll
  for (int j = 0; j < FCM_NT; ++j) bp[j] = w + (long)min(cls0 + 8 * j + g, classes - 1) * k + 8 * t;
  float acc[FCM_NT][4];
#pragma unroll
  for (int j = 0; j < FCM_NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  float4 an[FCM_DEPTH][4];
  uint4 bn[FCM_DEPTH][FCM_NT];
  // the filter rows do not depend on the previous kernel: their first FCM_DEPTH steps are requested before the wait
  auto load_b = [&](int slot, int kb) {   // 32 contraction indices per step: this lane's are kb + 8t .. 8t+7
#pragma unroll
    for (int j = 0; j < FCM_NT; ++j) bn[slot][j] = __ldg(reinterpret_cast<const uint4*>(bp[j] + kb));
  };
  auto load_a = [&](int slot, int kb) {
    an[slot][0] = __ldg(reinterpret_cast<const float4*>(a0p + kb)); an[slot][1] = __ldg(reinterpret_cast<const float4*>(a0p + kb + 4));
    an[slot][2] = __ldg(reinterpret_cast<const float4*>(a1p + kb)); an[slot][3] = __ldg(reinterpret_cast<const float4*>(a1p + kb + 4));
  };
#pragma unroll
  for (int d = 0; d < FCM_DEPTH; ++d) load_b(d, k_lo + 32 * d);      // k / FCM_WARPS is a multiple of 32 * FCM_DEPTH (launch_fc)
  pdl_wait();
#pragma unroll
  for (int d = 0; d < FCM_DEPTH; ++d) load_a(d, k_lo + 32 * d);
  for (int kb0 = k_lo; kb0 < k_hi; kb0 += 32 * FCM_DEPTH) {
#pragma unroll
    for (int d = 0; d < FCM_DEPTH; ++d) {
      const int kb = kb0 + 32 * d;
      const float4 a00 = an[d][0], a01 = an[d][1], a10 = an[d][2], a11 = an[d][3];
      uint4 b[FCM_NT];
#pragma unroll
      for (int j = 0; j < FCM_NT; ++j) b[j] = bn[d][j];
      if (kb + 32 * FCM_DEPTH < k_hi) { load_b(d, kb + 32 * FCM_DEPTH); load_a(d, kb + 32 * FCM_DEPTH); }
      // MMA slot k = (2t, 2t+1 | 2t+8, 2t+9) of step s <- this lane's elements (4s, 4s+1 | 4s+2, 4s+3)
      uint32_t ah[2][4], am[2][4], al[2][4];
      split3(a00.x, a00.y, ah[0][0], am[0][0], al[0][0]); split3(a10.x, a10.y, ah[0][1], am[0][1], al[0][1]);
      split3(a00.z, a00.w, ah[0][2], am[0][2], al[0][2]); split3(a10.z, a10.w, ah[0][3], am[0][3], al[0][3]);
      split3(a01.x, a01.y, ah[1][0], am[1][0], al[1][0]); split3(a11.x, a11.y, ah[1][1], am[1][1], al[1][1]);
      split3(a01.z, a01.w, ah[1][2], am[1][2], al[1][2]); split3(a11.z, a11.w, ah[1][3], am[1][3], al[1][3]);
#pragma unroll
      for (int j = 0; j < FCM_NT; ++j) {
        mma_bf16_16816(acc[j], al[0], b[j].x, b[j].y); mma_bf16_16816(acc[j], al[1], b[j].z, b[j].w);   // small pieces first
        mma_bf16_16816(acc[j], am[0], b[j].x, b[j].y); mma_bf16_16816(acc[j], am[1], b[j].z, b[j].w);
        mma_bf16_16816(acc[j], ah[0], b[j].x, b[j].y); mma_bf16_16816(acc[j], ah[1], b[j].z, b[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < FCM_NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) s_part[warp][g + 8 * (e >> 1)][8 * j + 2 * t + (e & 1)] = acc[j][e];
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * FCM_NT * 8; i += FCM_WARPS * 32) {
    const int r = i / (FCM_NT * 8), c = i % (FCM_NT * 8);
    const int img = img0 + r, cls = cls0 + c;
    if (img < n && cls < classes)
      out[(long)img * classes + cls] = ((s_part[0][r][c] + s_part[1][r][c]) + (s_part[2][r][c] + s_part[3][r][c])) + (bias ? __ldg(bias + cls) : 0.f);
  }
}

cudaError_t launch_fc(float* out, const float* pooled, const float* w_f32, const bf16* w_bf16, const float* bias, int n,
                      int k, int classes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (k % 256) return cudaErrorNotSupported;
  dim3 grid((classes + FC_CLS - 1) / FC_CLS, (n + FC_IMGS - 1) / FC_IMGS);
  const size_t smem = (size_t)FC_IMGS * k * sizeof(float);
  if (k % 256 || smem > 96 * 1024) return cudaErrorNotSupported;
  {
    cudaError_t e = ensure_dyn_smem((const void*)fc_kernel<bf16>, 96 * 1024);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)fc_kernel<float>, 96 * 1024);
    if (e != cudaSuccess) return e;
  }
  if (smem > 96 * 1024) return cudaErrorNotSupported;
  if (w_bf16 && k % (32 * FCM_WARPS * FCM_DEPTH) == 0) {   // bf16 filter: tensor cores, exact fp32 activations (three bf16 pieces)
    dim3 g2((classes + FCM_NT * 8 - 1) / (FCM_NT * 8), (n + 15) / 16);
    return launch_pdl(fc_mma_kernel, g2, dim3(FCM_WARPS * 32), 0, st, out, pooled, w_bf16, bias, n, k, classes);
  }
  if (w_bf16) fc_kernel<bf16><<<grid, 256, smem, st>>>(out, pooled, w_bf16, bias, n, k, classes);
  else fc_kernel<float><<<grid, 256, smem, st>>>(out, pooled, w_f32, bias, n, k, classes);
  return cudaGetLastError();
}

// one 128-thread CTA per image: every thread holds up to 8 logits in registers (all loads in flight
// at once), warp shuffles + a 4-entry shared fold for max / argmax and the sum of exp.  First maximum
// wins (strict `>` at MobileNet.c:2786).
constexpr int SM_THREADS = 128, SM_PER = 8;
__global__ void __launch_bounds__(SM_THREADS) softmax_kernel(const float* __restrict__ logits, int n, int classes,
                                                             float* __restrict__ prob, int* __restrict__ top1,
                                                             float* __restrict__ top1_prob, const HeadGather g) {
  __shared__ float s_mx[4], s_sum[4];
  __shared__ int s_arg[4];
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* z = logits + (long)img * classes;
  float v[SM_PER];
#pragma unroll
  for (int i = 0; i < SM_PER; ++i) {
    const int k = threadIdx.x + i * SM_THREADS;
    v[i] = k < classes ? z[k] : -INFINITY;
  }
  // logits gather of the data-parallel mode: this image's row goes straight into the gather block of every
  // rank (peer-mapped pointers, NVLink) — fire-and-forget stores, no collective kernel in the step
  for (int d = 0; d < g.n_dst; ++d) {
    float* dst = g.logits[d] + (g.row0 + img) * (long)classes;
#pragma unroll
    for (int i = 0; i < SM_PER; ++i) { const int k = threadIdx.x + i * SM_THREADS; if (k < classes) dst[k] = v[i]; }
  }
  float mx = -INFINITY;
  int arg = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < SM_PER; ++i)
    if (v[i] > mx) { mx = v[i]; arg = threadIdx.x + i * SM_THREADS; }   // indices increase with i: strict '>' keeps the first
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, off);
    const int oarg = __shfl_xor_sync(0xffffffffu, arg, off);
    if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
  }
  if (lane == 0) { s_mx[warp] = mx; s_arg[warp] = arg; }
  __syncthreads();
  mx = s_mx[0]; arg = s_arg[0];
#pragma unroll
  for (int w = 1; w < 4; ++w)
    if (s_mx[w] > mx || (s_mx[w] == mx && s_arg[w] < arg)) { mx = s_mx[w]; arg = s_arg[w]; }
  float e[SM_PER], sum = 0.f;
#pragma unroll
  for (int i = 0; i < SM_PER; ++i) { e[i] = __expf(v[i] - mx); sum += e[i]; }   // exp(-inf) = 0 for the padding
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) s_sum[warp] = sum;
  __syncthreads();
  sum = (s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]);
  const float inv = 1.0f / sum;
  if (prob) {
#pragma unroll
    for (int i = 0; i < SM_PER; ++i) {
      const int k = threadIdx.x + i * SM_THREADS;
      if (k < classes) prob[(long)img * classes + k] = e[i] * inv;
    }
  }
  if (threadIdx.x == 0) {
    if (top1) top1[img] = arg;
    if (top1_prob) top1_prob[img] = inv;
    for (int d = 0; d < g.n_dst; ++d) {
      if (g.top1[d]) g.top1[d][g.row0 + img] = arg;
      if (g.prob[d]) g.prob[d][g.row0 + img] = inv;
    }
  }
}

// generic fallback (more than SM_THREADS * SM_PER classes): one warp per image
__global__ void __launch_bounds__(128) softmax_warp_kernel(const float* __restrict__ logits, int n, int classes,
                                                           float* __restrict__ prob, int* __restrict__ top1,
                                                           float* __restrict__ top1_prob) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* z = logits + (long)warp * classes;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int k = lane; k < classes; k += 32) {
    const float v = z[k];
    if (v > mx) { mx = v; arg = k; }
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
                           cudaStream_t st, const HeadGather* gather) {
  if (n <= 0) return cudaSuccess;
  const HeadGather g = gather ? *gather : HeadGather{};
  if (classes <= SM_THREADS * SM_PER)
    return launch_pdl(softmax_kernel, dim3((unsigned)n), dim3(SM_THREADS), 0, st, logits, n, classes, prob, top1, top1_prob, g);
  if (g.n_dst) return cudaErrorNotSupported;
  else softmax_warp_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(logits, n, classes, prob, top1, top1_prob);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------
// Fused head (bf16 contexts): pool -> FC -> softmax/argmax as ONE kernel on thread-block clusters.
// A cluster of 8 CTAs owns 16 images.  The three stages need all-to-all data inside the group (the FC
// contracts over all 1024 pooled channels of an image, softmax over all 1000 classes), so the stages are
// separated by hardware cluster barriers (release / acquire: the small intermediate arrays travel through
// L2) instead of kernel boundaries — two launch gaps and two grid ramps less (22 us -> ~8 us per batch):
//   stage 1  CTA r pools images 2r, 2r+1: thread = 4 channels of one image, 49 independent 8-byte loads,
//            the same summation order as pool_kernel (4 pixel phases, folded left to right) -> fp32 means
//   stage 2  CTA r computes classes [125 r, 125 r + 125) for the 16 images with the exact 3-piece bf16
//            split of fc_mma_kernel (16 warps = 4 class groups x 4 k-splits, folded in a fixed order)
//   stage 3  CTA r: softmax / argmax of images 2r, 2r+1 (256 threads each, first maximum wins,
//            MobileNet.c:2786), and the logits-gather: each row is also stored straight into the gather
//            buffers of the peer GPUs (HeadGather: peer-mapped pointers over NVLink) — no collective kernel.
constexpr int HF_CL = 8, HF_IMGS = 16, HF_THREADS = 512, HF_CLS = 125;
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(HF_CL, 1, 1) __launch_bounds__(HF_THREADS, 1)
head_fused_kernel(const bf16* __restrict__ in, const bf16* __restrict__ w, const float* __restrict__ bias,
                  float* pooled, float* logits, int* __restrict__ top1, float* __restrict__ top1_prob, int n, int hw,
                  int classes, const HeadGather g) {
  constexpr int K = 1024;
  __shared__ float s_part[4][4][16][33];     // [class group][k-split][image][class]
  __shared__ float s_red[2][8];
  __shared__ int s_arg[2][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int img_base = (int)(blockIdx.x / HF_CL) * HF_IMGS;
  pdl_trigger();
  pdl_wait();

  // ---- stage 1: global average pool of this CTA's two images
  {
    const int img = img_base + 2 * (int)rank + (tid >> 8), q = tid & 255;
    if (img < n) {
      const bf16* p = in + (long)img * hw * K + 4 * q;
      float s[4][4];
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) s[ph][0] = s[ph][1] = s[ph][2] = s[ph][3] = 0.f;
      // groups of 16 pixels: 16 independent 8-byte loads in flight, the phase (i & 3) a compile-time constant
      for (int i0 = 0; i0 < hw; i0 += 16) {
        uint2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = i0 + j < hw ? __ldg(reinterpret_cast<const uint2*>(p + (long)(i0 + j) * K)) : make_uint2(0u, 0u);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (i0 + j < hw) {
            float* a = s[j & 3];
            a[0] += bf16lo_to_f32(v[j].x); a[1] += bf16hi_to_f32(v[j].x); a[2] += bf16lo_to_f32(v[j].y); a[3] += bf16hi_to_f32(v[j].y);
          }
        }
      }
      const float inv = 1.0f / (float)hw;
      float4 o;
      o.x = (s[0][0] + s[1][0] + s[2][0] + s[3][0]) * inv; o.y = (s[0][1] + s[1][1] + s[2][1] + s[3][1]) * inv;
      o.z = (s[0][2] + s[1][2] + s[2][2] + s[3][2]) * inv; o.w = (s[0][3] + s[1][3] + s[2][3] + s[3][3]) * inv;
      *reinterpret_cast<float4*>(pooled + (long)img * K + 4 * q) = o;
    }
  }
  cluster_barrier();

  // ---- stage 2: FC for classes [cls_lo, cls_lo + 125) x 16 images
  const int cls_lo = (int)rank * HF_CLS, cls_hi = min(cls_lo + HF_CLS, classes);
  {
    const int cg = warp >> 2, ks = warp & 3, gq = lane >> 2, t = lane & 3;
    const int cls0 = cls_lo + cg * 32;
    const int r0 = min(img_base + gq, n - 1), r1 = min(img_base + gq + 8, n - 1);
    const int k_lo = ks * (K / 4), k_hi = k_lo + K / 4;
    const float* a0p = pooled + (long)r0 * K + 8 * t;
    const float* a1p = pooled + (long)r1 * K + 8 * t;
    const bf16* bp[FCM_NT];
#pragma unroll
    for (int j = 0; j < FCM_NT; ++j) bp[j] = w + (long)min(cls0 + 8 * j + gq, classes - 1) * K + 8 * t;
    float acc[FCM_NT][4];
#pragma unroll
    for (int j = 0; j < FCM_NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    float4 an[4];
    uint4 bn[FCM_NT];
    auto load = [&](int kb) {
      // the pooled means were written by other SMs of the cluster in this kernel: plain (coherent) loads
      an[0] = *reinterpret_cast<const float4*>(a0p + kb); an[1] = *reinterpret_cast<const float4*>(a0p + kb + 4);
      an[2] = *reinterpret_cast<const float4*>(a1p + kb); an[3] = *reinterpret_cast<const float4*>(a1p + kb + 4);
#pragma unroll
      for (int j = 0; j < FCM_NT; ++j) bn[j] = __ldg(reinterpret_cast<const uint4*>(bp[j] + kb));
    };
    load(k_lo);
    for (int kb = k_lo; kb < k_hi; kb += 32) {
      const float4 a00 = an[0], a01 = an[1], a10 = an[2], a11 = an[3];
      uint4 b[FCM_NT];
#pragma unroll
      for (int j = 0; j < FCM_NT; ++j) b[j] = bn[j];
      if (kb + 32 < k_hi) load(kb + 32);
      uint32_t ah[2][4], am[2][4], al[2][4];
      split3(a00.x, a00.y, ah[0][0], am[0][0], al[0][0]); split3(a10.x, a10.y, ah[0][1], am[0][1], al[0][1]);
      split3(a00.z, a00.w, ah[0][2], am[0][2], al[0][2]); split3(a10.z, a10.w, ah[0][3], am[0][3], al[0][3]);
      split3(a01.x, a01.y, ah[1][0], am[1][0], al[1][0]); split3(a11.x, a11.y, ah[1][1], am[1][1], al[1][1]);
      split3(a01.z, a01.w, ah[1][2], am[1][2], al[1][2]); split3(a11.z, a11.w, ah[1][3], am[1][3], al[1][3]);
#pragma unroll
      for (int j = 0; j < FCM_NT; ++j) {
        mma_bf16_16816(acc[j], al[0], b[j].x, b[j].y); mma_bf16_16816(acc[j], al[1], b[j].z, b[j].w);   // small pieces first
        mma_bf16_16816(acc[j], am[0], b[j].x, b[j].y); mma_bf16_16816(acc[j], am[1], b[j].z, b[j].w);
        mma_bf16_16816(acc[j], ah[0], b[j].x, b[j].y); mma_bf16_16816(acc[j], ah[1], b[j].z, b[j].w);
      }
    }
#pragma unroll
    for (int j = 0; j < FCM_NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s_part[cg][ks][gq + 8 * (e >> 1)][8 * j + 2 * t + (e & 1)] = acc[j][e];
  }
  __syncthreads();
  for (int i = tid; i < 4 * 16 * 32; i += HF_THREADS) {
    const int cg = i >> 9, r = (i >> 5) & 15, c = i & 31;
    const int img = img_base + r, cls = cls_lo + cg * 32 + c;
    if (img < n && cls < cls_hi)
      logits[(long)img * classes + cls] =
          ((s_part[cg][0][r][c] + s_part[cg][1][r][c]) + (s_part[cg][2][r][c] + s_part[cg][3][r][c])) + (bias ? __ldg(bias + cls) : 0.f);
  }
  cluster_barrier();

  // ---- stage 3: softmax / argmax of this CTA's two images (+ the gather stores)
  {
    const int sub = tid >> 8, t = tid & 255, w8 = (tid >> 5) & 7;
    const int img = img_base + 2 * (int)rank + sub;
    const bool live = img < n;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = t + i * 256;
      v[i] = (live && k < classes) ? logits[(long)img * classes + k] : -INFINITY;
    }
    for (int d = 0; d < g.n_dst; ++d) {          // peer GPUs' gather buffers (row = global image index)
      float* dst = g.logits[d] + (g.row0 + img) * (long)classes;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int k = t + i * 256; if (live && k < classes) dst[k] = v[i]; }
    }
    float mx = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (v[i] > mx) { mx = v[i]; arg = t + i * 256; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, off);
      const int oarg = __shfl_xor_sync(0xffffffffu, arg, off);
      if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
    }
    if (lane == 0) { s_red[sub][w8] = mx; s_arg[sub][w8] = arg; }
    __syncthreads();
    mx = s_red[sub][0]; arg = s_arg[sub][0];
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (s_red[sub][k] > mx || (s_red[sub][k] == mx && s_arg[sub][k] < arg)) { mx = s_red[sub][k]; arg = s_arg[sub][k]; }
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) sum += __expf(v[i] - mx);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) s_red[sub][w8] = sum;
    __syncthreads();
    if (t == 0 && live) {
      sum = ((s_red[sub][0] + s_red[sub][1]) + (s_red[sub][2] + s_red[sub][3])) + ((s_red[sub][4] + s_red[sub][5]) + (s_red[sub][6] + s_red[sub][7]));
      const float p1 = 1.0f / sum;
      if (top1) top1[img] = arg;
      if (top1_prob) top1_prob[img] = p1;
      for (int d = 0; d < g.n_dst; ++d) {
        if (g.top1[d]) g.top1[d][g.row0 + img] = arg;
        if (g.prob[d]) g.prob[d][g.row0 + img] = p1;
      }
    }
  }
}

cudaError_t launch_head_fused(const bf16* in, int n, int hw, int c, const mnv1_filter* fc, float* pooled_scratch,
                              float* logits, int* top1, float* top1_prob, int classes, const HeadGather& g,
                              cudaStream_t st) {
  if (c != 1024 || classes > HF_CL * HF_CLS || !fc->w_bf16 || !switches().fused_head) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(((n + HF_IMGS - 1) / HF_IMGS) * HF_CL));
  cfg.blockDim = dim3(HF_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, head_fused_kernel, in, (const bf16*)fc->w_bf16, (const float*)fc->shift, pooled_scratch,
                            logits, top1, top1_prob, n, hw, classes, g);
}

cudaError_t launch_head(mnv1_dtype dt, const void* in, int n, int hw, int c, const mnv1_filter* fc,
                        float* pooled_scratch, float* logits, int* top1, float* top1_prob, int classes,
                        const HeadGather& g, cudaStream_t st, int* launches) {
  if (dt == MNV1_BF16) {
    cudaError_t fe = launch_head_fused((const bf16*)in, n, hw, c, fc, pooled_scratch, logits, top1, top1_prob, classes, g, st);
    if (fe != cudaErrorNotSupported) { if (launches) *launches = 1; return fe; }
  }
  cudaError_t e = launch_pool(dt, pooled_scratch, in, n, hw, c, /*out_f32=*/true, st);
  if (e != cudaSuccess) return e;
  e = launch_fc(logits, pooled_scratch, fc->w_f32, dt == MNV1_BF16 ? fc->w_bf16 : nullptr, fc->shift, n, c, classes, st);
  if (e == cudaErrorNotSupported) {
    Epilogue ep{nullptr, fc->shift, MNV1_ACT_NONE};
    e = launch_fc_f32in(logits, pooled_scratch, fc->w_f32, dt == MNV1_BF16 ? fc->w_bf16 : nullptr, n, c, classes, ep, st);
  }
  if (e != cudaSuccess) return e;
  if (launches) *launches = 2;
  if (top1 || top1_prob || g.n_dst) {
    e = launch_softmax(logits, n, classes, nullptr, top1, top1_prob, st, &g);
    if (launches) *launches = 3;
  }
  return e;
}

}  // namespace mnv1
