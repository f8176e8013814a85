// int8.cu — the integer "reference-faithful" mode (MNV1_U8 contexts).
//
// The only arithmetic the reference actually has is integer: `unsigned char` feature maps times `int`
// filter values summed in an `int`, `if (sum <= 0) sum = 0`, and the sum stored back into an
// `unsigned char` (kernel.cl:2-3,10,52-56,62,87-90,94,109-112; the host keeps the filters in int8_t buffers,
// MobileNet.c:116,248).  This file runs exactly that on the GPU:
//   activations u8 (NHWC on the device), filters s8, accumulation s32,
//   out = store_u8( max(acc + bias[c], 0) >> rshift )      (ReLU optional per filter)
//   store_u8 = the C conversion to unsigned char (wrap modulo 256, what kernel.cl does) or saturation to 255,
//   selectable per context (mnv1_ctx_set_u8_store).
// With bias = 0, rshift = 0 and the wrapping store every layer is bit-identical to the literal kernel.cl
// launched once per output channel (tests/test_gpu_int8.py pins full-size layers against oracle/_ref).
//
//   pointwise / FC   tcgen05.mma.kind::i8 (u8 x s8 -> s32 in TMEM), TMA-fed 128B-swizzled K-major tiles
//   depthwise, stem  DP4A stencils: one u8 x s8 product per lane-byte (masked words), s32 accumulators
//   pool             integer sum / (f*f), the truncating division of kernel.cl:129
#include <cstdio>

#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

__device__ __forceinline__ uint32_t store_u8(int v, int wrap) {
  // v >= 0 after ReLU; without ReLU a negative value wraps like the C conversion, or saturates to 0
  if (wrap) return (uint32_t)v & 0xffu;
  return (uint32_t)min(max(v, 0), 255);
}
// d = c + sum_i a.u8[i] * b.s8[i]  (mixed signedness: the CUDA intrinsic only offers s32.s32 / u32.u32)
__device__ __forceinline__ int dp4a_u8s8(uint32_t a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int finish(int acc, int bias, int relu, int rshift) {
  int v = acc + bias;
  if (relu) v = max(v, 0);
  return v >> rshift;                      // arithmetic shift: floor((acc + bias) / 2^rshift)
}

// ------------------------------------------------------------------------------------------ pointwise / FC
constexpr int I8_BM = 128, I8_BK = 128, I8_STAGES = 4, I8_THREADS = 192;

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::i8 instruction descriptor: D = s32 (bits 4-5 = 2), A = u8 (bits 7-9 = 0), B = s8 (bits 10-12 = 1), both
// K-major, N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_u8s8_m128(int n) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct I8Params {
  uint8_t* out;          // [M][Cout] u8
  const int* bias;       // [Cout] or nullptr
  long M;
  int K, Cout, relu, rshift, wrap;
};

// out[M][Cout] = store_u8(finish(in[M][K] . w[Cout][K]^T)).  One CTA per (128-row tile, BN-column tile):
// warp 0 = TMA producer (4-stage ring of [128 x 128 B] A and [BN x 128 B] B tiles; rows / columns / K past
// the tensor are zero-filled by the TMA unit), warp 1 = TMEM allocator + single-thread MMA issuer
// (UMMA 128 x BN x 32, u8 x s8 -> s32), warps 2-5 = epilogue (tcgen05.ld of their 32 lanes, integer finish,
// 4 bytes per word, 128-bit stores).
template <int BN>
__global__ void __launch_bounds__(I8_THREADS) pointwise_i8_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b,
                                                                  const I8Params p) {
  constexpr uint32_t A_BYTES = I8_BM * I8_BK, B_BYTES = BN * I8_BK, STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t TM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem + I8_STAGES * STAGE;
  const uint32_t full = bars, empty = bars + 8u * I8_STAGES, tm_full = empty + 8u * I8_STAGES, tmem_slot = tm_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_idx = blockIdx.x * I8_BM, n_idx = blockIdx.y * BN;
  const int num_kb = (p.K + I8_BK - 1) / I8_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b);
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(full + 8u * s, 1); mbar_init(empty + 8u * s, 1); }
    mbar_init(tm_full, 1);
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % I8_STAGES;
        mbar_wait(empty + 8u * s, ((kb / I8_STAGES) & 1) ^ 1u);
        mbar_expect_tx(full + 8u * s, STAGE);
        tma_load_2d(smem + s * STAGE, &tmap_a, full + 8u * s, kb * I8_BK, m_idx);
        tma_load_2d(smem + s * STAGE + A_BYTES, &tmap_b, full + 8u * s, kb * I8_BK, n_idx);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_u8s8_m128(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % I8_STAGES;
        mbar_wait(full + 8u * s, (kb / I8_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem + s * STAGE), db = umma_desc_sw128(smem + s * STAGE + A_BYTES);
#pragma unroll
        for (int k = 0; k < I8_BK / 32; ++k) umma_i8(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit(empty + 8u * s);
      }
      umma_commit(tm_full);
    }
  } else {
    const int quarter = warp & 3;
    const long row = (long)m_idx + quarter * 32 + lane;
    mbar_wait(tm_full, 0);
    tc_fence_after();
    const bool vec = (p.Cout & 15) == 0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      const int col0 = n_idx + c0;
      uint32_t q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int col = col0 + 4 * j + b;
          const int bias = (p.bias && col < p.Cout) ? __ldg(p.bias + col) : 0;
          w |= store_u8(finish((int)v[4 * j + b], bias, p.relu, p.rshift), p.wrap) << (8 * b);
        }
        q[j] = w;
      }
      if (row < p.M) {
        uint8_t* o = p.out + row * p.Cout + col0;
        if (vec && col0 + 32 <= p.Cout) {
          reinterpret_cast<uint4*>(o)[0] = make_uint4(q[0], q[1], q[2], q[3]);
          reinterpret_cast<uint4*>(o)[1] = make_uint4(q[4], q[5], q[6], q[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.Cout) o[j] = (uint8_t)(q[j >> 2] >> (8 * (j & 3)));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

cudaError_t encode_u8(CUtensorMap* map, const void* base, uint64_t rows, uint64_t k, uint32_t box_rows, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  cuuint64_t gdim[2] = {k, rows};
  cuuint64_t gstride[1] = {k};
  cuuint32_t box[2] = {(cuuint32_t)I8_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { if (err) *err = "pointwise_i8: tensor map encode failed"; return cudaErrorInvalidValue; }
  return cudaSuccess;
}

template <int BN>
cudaError_t launch_i8(const CUtensorMap& ta, const CUtensorMap& tb, const I8Params& p, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)I8_STAGES * (I8_BM * I8_BK + BN * I8_BK) + 16 * I8_STAGES + 32;
  cudaError_t e = ensure_dyn_smem((const void*)pointwise_i8_kernel<BN>, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((p.M + I8_BM - 1) / I8_BM), (unsigned)((p.Cout + BN - 1) / BN));
  pointwise_i8_kernel<BN><<<grid, I8_THREADS, smem, st>>>(ta, tb, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ depthwise
// NHWC u8.  A thread owns one output pixel x 4 channels (one 32-bit word): for every tap the input word
// is masked to one byte at a time, so that dp4a's 4-way dot product degenerates to the single u8 x s8
// product of that channel.  taps: [9][C] s8 (tap-major, like the bf16 path's [9][C] floats).
__global__ void __launch_bounds__(256) depthwise_u8_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                                           const int8_t* __restrict__ taps, const int* __restrict__ bias,
                                                           int n, int H, int W, int C, int stride, int pad_lo, int relu,
                                                           int rshift, int wrap) {
  const int Ho = H / stride, Wo = W / stride, C4 = C >> 2;
  const long total = (long)n * Ho * Wo * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long r = i / C4;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const int img = (int)(r / Ho);
    int acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ty = 0; ty < 3; ++ty) {
      const int yy = y * stride + ty - pad_lo;
#pragma unroll
      for (int tx = 0; tx < 3; ++tx) {
        const int xx = x * stride + tx - pad_lo;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;          // zero padding on every border
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(in + (((long)img * H + yy) * W + xx) * C) + c4);
        const int t = __ldg(reinterpret_cast<const int*>(taps + (long)(ty * 3 + tx) * C) + c4);
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[b] = dp4a_u8s8(v & (0xffu << (8 * b)), t, acc[b]);   // u8 x s8, one channel
      }
    }
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      w |= store_u8(finish(acc[b], bias ? __ldg(bias + 4 * c4 + b) : 0, relu, rshift), wrap) << (8 * b);
    reinterpret_cast<uint32_t*>(out + (((long)img * Ho + y) * Wo + x) * C)[c4] = w;
  }
}

// ------------------------------------------------------------------------------------------ stem
// A thread owns one output pixel and all 32 filters.  Its 27 input bytes (R taps, G taps, B taps in the
// `findex` order of kernel.cl:15-51) are packed into 7 words; every filter's 27 s8 values are packed the same
// way, so one filter costs 7 DP4A.  wq: [32][7] words in shared memory.
__global__ void __launch_bounds__(128) stem_u8_kernel(uint8_t* __restrict__ out, const StemArgs a, const int* __restrict__ wq,
                                                      const int* __restrict__ bias, int relu, int rshift, int wrap) {
  __shared__ int s_w[32 * 7];
  __shared__ int s_b[32];
  for (int i = threadIdx.x; i < 32 * 7; i += blockDim.x) s_w[i] = wq[i];
  if (threadIdx.x < 32) s_b[threadIdx.x] = bias ? bias[threadIdx.x] : 0;
  __syncthreads();
  const int Ho = a.rows / a.stride, Wo = a.cols / a.stride;
  const long total = (long)a.n * Ho * Wo;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int img = (int)(i / ((long)Wo * Ho));
    uint32_t pk[7] = {0, 0, 0, 0, 0, 0, 0};
    const uint8_t* planes[3] = {a.r + (long)img * a.img_stride, a.g + (long)img * a.img_stride, a.b + (long)img * a.img_stride};
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
      for (int ty = 0; ty < 3; ++ty)
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          const int yy = y * a.stride + ty - a.pad_lo, xx = x * a.stride + tx - a.pad_lo;
          uint32_t v = 0;
          if (yy >= 0 && yy < a.rows && xx >= 0 && xx < a.cols) v = planes[pl][((long)yy * a.cols + xx) * a.pix_stride];
          const int t = pl * 9 + ty * 3 + tx;
          pk[t >> 2] |= v << (8 * (t & 3));
        }
    uint32_t* o = reinterpret_cast<uint32_t*>(out + i * 32);
#pragma unroll
    for (int f4 = 0; f4 < 8; ++f4) {
      uint32_t w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int f = 4 * f4 + b;
        int acc = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) acc = dp4a_u8s8(pk[k], s_w[f * 7 + k], acc);
        w |= store_u8(finish(acc, s_b[f], relu, rshift), wrap) << (8 * b);
      }
      o[f4] = w;
    }
  }
}

// ------------------------------------------------------------------------------------------ pool, layout, logits
// kernel.cl:116-131: integer sum of the f*f values of a channel, truncating division.
__global__ void pool_u8_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in, int n, int hw, int c, int wrap) {
  const long total = (long)n * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const long img = i / c;
    int sum = 0;
    for (int px = 0; px < hw; ++px) sum += in[(img * hw + px) * c + ch];
    out[i] = (uint8_t)store_u8(sum / hw, wrap);
  }
}
template <typename TI>
__global__ void nchw_to_nhwc_u8_kernel(uint8_t* __restrict__ out, const TI* __restrict__ in, int n, int c, int hw) {
  const long total = (long)n * c * hw;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const long r = i / c;
    const int px = (int)(r % hw);
    const long img = r / hw;
    const TI v = in[(img * c + ch) * hw + px];
    if constexpr (sizeof(TI) == 1) out[i] = (uint8_t)v;
    else out[i] = (uint8_t)fminf(fmaxf(rintf((float)v), 0.f), 255.f);
  }
}
template <typename TO>
__global__ void nhwc_u8_to_nchw_kernel(TO* __restrict__ out, const uint8_t* __restrict__ in, int n, int c, int hw) {
  const long total = (long)n * c * hw;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = (int)(i % hw);
    const long r = i / hw;
    const int ch = (int)(r % c);
    const long img = r / c;
    out[i] = (TO)in[(img * hw + px) * c + ch];
  }
}
__global__ void u8_to_f32_kernel(float* __restrict__ out, const uint8_t* __restrict__ in, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

unsigned grid_for(long total, int block) {
  long g = (total + block - 1) / block;
  return (unsigned)(g < 1 ? 1 : (g > 148L * 32 ? 148L * 32 : g));
}

}  // namespace

// ---------------------------------------------------------------------------------------------- launchers
cudaError_t launch_pointwise_i8(uint8_t* out, const uint8_t* in, const mnv1_filter* f, long m, int k, int cout, int wrap,
                                cudaStream_t st, std::string* err) {
  if (!f->w_s8 || k % 16 || k < 16) { if (err) *err = "pointwise (u8): Cin must be a multiple of 16"; return cudaErrorNotSupported; }
  if (m <= 0) return cudaSuccess;
  const int bn = cout >= 256 ? 256 : cout > 64 ? 128 : cout > 32 ? 64 : cout > 16 ? 32 : 16;
  CUtensorMap ta, tb;
  cudaError_t e = encode_u8(&ta, in, (uint64_t)m, (uint64_t)k, I8_BM, err);
  if (e == cudaSuccess) e = encode_u8(&tb, f->w_s8, (uint64_t)cout, (uint64_t)k, (uint32_t)bn, err);
  if (e != cudaSuccess) return e;
  I8Params p{out, f->bias_i32, m, k, cout, f->act != MNV1_ACT_NONE ? 1 : 0, f->rshift, wrap};
  switch (bn) {
    case 256: return launch_i8<256>(ta, tb, p, st);
    case 128: return launch_i8<128>(ta, tb, p, st);
    case 64: return launch_i8<64>(ta, tb, p, st);
    case 32: return launch_i8<32>(ta, tb, p, st);
    default: return launch_i8<16>(ta, tb, p, st);
  }
}

cudaError_t launch_depthwise_u8(uint8_t* out, const uint8_t* in, const mnv1_filter* f, int n, int rows, int cols, int stride,
                                int c, int pad_lo, int wrap, cudaStream_t st) {
  if (c % 4) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
  const long total = (long)n * (rows / stride) * (cols / stride) * (c / 4);
  depthwise_u8_kernel<<<grid_for(total, 256), 256, 0, st>>>(out, in, f->w_s8, f->bias_i32, n, rows, cols, c, stride, pad_lo,
                                                            f->act != MNV1_ACT_NONE ? 1 : 0, f->rshift, wrap);
  return cudaGetLastError();
}

cudaError_t launch_stem_u8(uint8_t* out, const StemArgs& a, const mnv1_filter* f, int wrap, cudaStream_t st) {
  if (a.cout != 32 || !f->w_q32) return cudaErrorNotSupported;
  if (a.n <= 0) return cudaSuccess;
  const long total = (long)a.n * (a.rows / a.stride) * (a.cols / a.stride);
  stem_u8_kernel<<<grid_for(total, 128), 128, 0, st>>>(out, a, f->w_q32, f->bias_i32, f->act != MNV1_ACT_NONE ? 1 : 0, f->rshift, wrap);
  return cudaGetLastError();
}

cudaError_t launch_pool_u8(uint8_t* out, const uint8_t* in, int n, int hw, int c, int wrap, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  pool_u8_kernel<<<grid_for((long)n * c, 256), 256, 0, st>>>(out, in, n, hw, c, wrap);
  return cudaGetLastError();
}

cudaError_t launch_u8_layout(int dir, void* out, const void* in, bool host_is_u8, int n, int c, int hw, cudaStream_t st) {
  const long total = (long)n * c * hw;
  if (total <= 0) return cudaSuccess;
  const unsigned g = grid_for(total, 256);
  if (dir == 0) {   // planar host order -> NHWC u8
    if (host_is_u8) nchw_to_nhwc_u8_kernel<uint8_t><<<g, 256, 0, st>>>((uint8_t*)out, (const uint8_t*)in, n, c, hw);
    else nchw_to_nhwc_u8_kernel<float><<<g, 256, 0, st>>>((uint8_t*)out, (const float*)in, n, c, hw);
  } else {          // NHWC u8 -> planar
    if (host_is_u8) nhwc_u8_to_nchw_kernel<uint8_t><<<g, 256, 0, st>>>((uint8_t*)out, (const uint8_t*)in, n, c, hw);
    else nhwc_u8_to_nchw_kernel<float><<<g, 256, 0, st>>>((float*)out, (const uint8_t*)in, n, c, hw);
  }
  return cudaGetLastError();
}

cudaError_t launch_u8_to_f32(float* out, const uint8_t* in, long count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  u8_to_f32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(out, in, count);
  return cudaGetLastError();
}

}  // namespace mnv1
