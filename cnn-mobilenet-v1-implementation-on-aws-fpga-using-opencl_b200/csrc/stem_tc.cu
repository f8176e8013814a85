// stem_tc.cu — layer 1 (3x3 conv, 3 colour planes -> 32 maps) on tcgen05 for bf16 contexts.
//
// Same contract as stem.cu (`convolute`, kernel.cl:2-60): tap order R,G,B planes x row offset x
// column offset (kernel.cl:15-51), intended semantics.  At 21.7 MFLOP per image the CUDA-core
// version is FFMA-bound (5.5 GFLOP per 256-image batch vs a 37 us HBM roofline), so the layer
// runs as an implicit GEMM  D[pixel][32] = A[pixel][27] . B[32][27]^T :
//   * A: each thread gathers the 27 u8 taps of ONE output pixel and widens them EXACTLY to fp16
//     (0x6400|b is 1024+b in fp16; one HSUB2 per pair), written as a K-major 128B-swizzled row;
//   * B: the filter bank times the input scale, rounded to fp16, built once per CTA;
//   * D: 128 pixels x 32 channels, fp32 in TMEM, two accumulator/A-tile buffers so the gather of
//     tile i+1 overlaps the MMA + epilogue of tile i;
//   * epilogue: tcgen05.ld -> fma(scale, shift') -> ReLU on the bf16 convert, 6-cap as
//     min.bf16x2 -> 64B-swizzled staging -> TMA store of the contiguous 8 KB NHWC tile.
// The input transform x' = s*x + b is folded:  conv(x') = sum wq*(x - p0), wq = fp16(w*s),
// p0 = -b/s (127.5 for the Keras preprocessing, exact in fp16); border taps read p0, so the
// zero padding of x' is exact and the constant -p0*sum(wq) moves into the shift.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>

#include "common.cuh"

namespace mnv1 {
namespace {

constexpr int ST_THREADS = 288;   // 4 gather warps + 4 epilogue warps + 1 MMA/TMEM warp
constexpr int ST_C = 32;
constexpr uint32_t ST_A_BYTES = 128 * 128;   // 128 pixels x 128-byte swizzle rows (first 64 B = 32 fp16 used)
constexpr uint32_t ST_B_BYTES = 32 * 128;
constexpr uint32_t ST_O_BYTES = 128 * 64;    // 128 pixels x 32 bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ST_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra ST_DONE;\n"
      "bra ST_WAIT;\n"
      "ST_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// K-major, 128B-swizzled operand descriptor (see pointwise_tc.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16, D = f32, A = B = fp16 (format 0), K-major, N = 32, M = 128
constexpr uint32_t ST_IDESC = (1u << 4) | ((uint32_t)(ST_C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(ST_IDESC), "r"(accumulate)
      : "memory");
}
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi, uint32_t cap2) {
  uint32_t d;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  asm("min.bf16x2 %0, %0, %1;" : "+r"(d) : "r"(cap2));
  return d;
}
// two u8 values (low bytes of a, b) -> exact fp16 pair
__device__ __forceinline__ uint32_t u8x2_to_f16x2(uint32_t a, uint32_t b) {
  uint32_t v = __byte_perm(a, b, 0x5410) | 0x64006400u;   // (1024 + a, 1024 + b)
  __half2 h = __hsub2(*reinterpret_cast<__half2*>(&v), __half2half2(__ushort_as_half((unsigned short)0x6400)));
  return *reinterpret_cast<uint32_t*>(&h);
}

struct StemTcParams {
  const uint8_t *r, *g, *b;
  int pix_stride; long img_stride;
  int n, rows, cols, orows, ocols, pad_lo;
  const __half* wq;      // [32][32] fp16: wq[o][k] = fp16(w[o][k] * in_scale), k >= 27 zero
  const float* scale;    // [32] or nullptr
  const float* shift2;   // [32] shift + scale * (-p0 * sum_k wq[o][k])
  uint32_t cap2;
  uint32_t pad_f16x2;    // fp16(p0) in both halves
  long m_total;
};

// IL: the image is the interleaved RGB payload (pix_stride 3), so the 9 bytes of a window row are
// contiguous and every tap is [row pointer + immediate]; otherwise three separate planes.
//
// Warp roles (288 threads): warps 0-3 gather (thread g builds row g of the A tile), warps 4-7
// epilogue (warp 4+q owns TMEM lanes 32q..), warp 8 = TMEM allocator + MMA issuer.  Two A tiles,
// two accumulators and two staging buffers; mbarriers a_full[2] (128 gather arrivals),
// mma_done[2] (tcgen05.commit: "D ready" for the epilogue AND "A free" for the gatherers),
// tmem_free[2] (4 epilogue-warp arrivals).  Gather of tile i+1, MMA of tile i and epilogue of
// tile i-1 therefore run concurrently instead of back to back.
template <int S, bool RELU, bool IL>
__global__ void __launch_bounds__(ST_THREADS, 3)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmap_out, const StemTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem;                         // 2 x 16 KB
  const uint32_t sB = smem + 2 * ST_A_BYTES;        // 4 KB
  const uint32_t sO = sB + ST_B_BYTES;              // 2 x 8 KB
  const uint32_t bars = sO + 2 * ST_O_BYTES;        // a_full[2] | mma_done[2] | tmem_free[2]
  const uint32_t tmem_slot = bars + 48;
  __shared__ __align__(16) float s_scale[ST_C];
  __shared__ __align__(16) float s_shift[ST_C];
  const uint32_t a_full = bars, mma_done = bars + 16, tmem_free = bars + 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid < ST_C) {
    s_scale[tid] = p.scale ? p.scale[tid] : 1.f;
    s_shift[tid] = p.shift2[tid];
    // B tile: thread t < 32 writes filter row t (32 fp16 = 4 chunks) with the 128B swizzle
    const uint4* src = reinterpret_cast<const uint4*>(p.wq + tid * 32);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 v = __ldg(src + c);
      sts128(sB + tid * 128 + ((c ^ (tid & 7)) << 4), v.x, v.y, v.z, v.w);
    }
  }
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    for (int b = 0; b < 2; ++b) { mbar_init(a_full + 8 * b, 128); mbar_init(mma_done + 8 * b, 1); mbar_init(tmem_free + 8 * b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();   // the previous step's kernels are done with the activation arena; the images are in place

  const int Wo = p.ocols, HoWo = p.orows * p.ocols, W = p.cols, H = p.rows, ps = p.pix_stride;
  const long tiles = (p.m_total + 127) / 128;

  if (warp < 4) {
    // ======================= gather warps =======================
    int i = 0;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int buf = i & 1, k = i >> 1;
      const long m = t * 128 + tid;
      uint32_t a[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = 0u;
      if (m < p.m_total) {
        // 32-bit index math (m_total < 2^31 is checked on the host): a 64-bit divide is a subroutine
        const unsigned mu = (unsigned)m;
        const int img = (int)(mu / (unsigned)HoWo);
        const int rem = (int)(mu - (unsigned)img * (unsigned)HoWo);
        const int oy = (int)((unsigned)rem / (unsigned)Wo), ox = rem - oy * Wo;
        const int iy0 = oy * S - p.pad_lo, ix0 = ox * S - p.pad_lo;
        const uint8_t* base[3] = {p.r + (long)img * p.img_stride, p.g + (long)img * p.img_stride,
                                  p.b + (long)img * p.img_stride};
        if (iy0 >= 0 && ix0 >= 0 && iy0 + 2 < H && ix0 + 2 < W) {
          uint32_t tap[28];
          tap[27] = 0;
#ifdef ST_EXP_NOLOAD
          for (int q = 0; q < 27; ++q) tap[q] = (tid + q) & 255;
          if (false) {
#else
          if (IL) {
#endif
            const uint8_t* r0 = base[0] + ((long)iy0 * W + ix0) * 3;
            const uint8_t* r1 = r0 + (long)W * 3;
            const uint8_t* r2 = r1 + (long)W * 3;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                tap[pl * 9 + 0 + jj] = __ldg(r0 + jj * 3 + pl);
                tap[pl * 9 + 3 + jj] = __ldg(r1 + jj * 3 + pl);
                tap[pl * 9 + 6 + jj] = __ldg(r2 + jj * 3 + pl);
              }
          } else {
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
              const uint8_t* r0 = base[pl] + (long)iy0 * W + ix0;
              const uint8_t* r1 = r0 + W;
              const uint8_t* r2 = r1 + W;
#pragma unroll
              for (int jj = 0; jj < 3; ++jj) {
                tap[pl * 9 + 0 + jj] = __ldg(r0 + jj);
                tap[pl * 9 + 3 + jj] = __ldg(r1 + jj);
                tap[pl * 9 + 6 + jj] = __ldg(r2 + jj);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 14; ++q) a[q] = u8x2_to_f16x2(tap[2 * q], tap[2 * q + 1]);
          a[13] &= 0x0000ffffu;
        } else {
          // border pixel: out-of-range taps read p0 (the raw value whose transform is zero)
          unsigned short hv[28];
          const unsigned short padv = (unsigned short)(p.pad_f16x2 & 0xffffu);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
#pragma unroll
              for (int jj = 0; jj < 3; ++jj) {
                const int y = iy0 + ii, x = ix0 + jj;
                unsigned short h = padv;
                if (y >= 0 && y < H && x >= 0 && x < W)
                  h = __half_as_ushort(__ushort2half_rn((unsigned short)__ldg(base[pl] + ((long)y * W + x) * ps)));
                hv[pl * 9 + ii * 3 + jj] = h;
              }
          hv[27] = 0;
#pragma unroll
          for (int q = 0; q < 14; ++q) a[q] = (uint32_t)hv[2 * q] | ((uint32_t)hv[2 * q + 1] << 16);
        }
      }
      // A[buf] was last read by the MMAs of tile i-2: wait for their commit before overwriting
      if (k > 0) mbar_wait(mma_done + 8 * buf, (uint32_t)(k - 1) & 1u);
      const uint32_t arow = sA + buf * ST_A_BYTES + tid * 128;
#ifndef ST_EXP_NOGATHERSTS
#pragma unroll
      for (int c = 0; c < 4; ++c) sts128(arow + ((c ^ (tid & 7)) << 4), a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#else
      if (a[0] == 0x12345678u) sts128(arow, a[0], a[1], a[2], a[3]);
#endif
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_full + 8 * buf) : "memory");
    }
  } else if (warp < 8) {
    // ======================= epilogue warps =======================
    const int q = warp & 3, row = q * 32 + lane;       // TMEM lane quarter q, tile row
    const bool leader = tid == 128;
    int i = 0;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int buf = i & 1, k = i >> 1;
      mbar_wait(mma_done + 8 * buf, (uint32_t)k & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[32];
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ST_C);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_free + 8 * buf) : "memory");
#ifdef ST_EXP_NOEPI
      if (v[0] == 0x12345678u) p.shift2 ? (void)0 : (void)0;
      continue;
#endif
      // staging buffer `buf` was read by the TMA store of tile i-2
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t orow = sO + buf * ST_O_BYTES + row * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 s0 = *reinterpret_cast<const float4*>(&s_scale[c * 8]), s1 = *reinterpret_cast<const float4*>(&s_scale[c * 8 + 4]);
        const float4 t0 = *reinterpret_cast<const float4*>(&s_shift[c * 8]), t1 = *reinterpret_cast<const float4*>(&s_shift[c * 8 + 4]);
        const uint32_t p0 = pack2<RELU>(fmaf(__uint_as_float(v[c * 8 + 0]), s0.x, t0.x), fmaf(__uint_as_float(v[c * 8 + 1]), s0.y, t0.y), p.cap2);
        const uint32_t p1 = pack2<RELU>(fmaf(__uint_as_float(v[c * 8 + 2]), s0.z, t0.z), fmaf(__uint_as_float(v[c * 8 + 3]), s0.w, t0.w), p.cap2);
        const uint32_t p2 = pack2<RELU>(fmaf(__uint_as_float(v[c * 8 + 4]), s1.x, t1.x), fmaf(__uint_as_float(v[c * 8 + 5]), s1.y, t1.y), p.cap2);
        const uint32_t p3 = pack2<RELU>(fmaf(__uint_as_float(v[c * 8 + 6]), s1.z, t1.z), fmaf(__uint_as_float(v[c * 8 + 7]), s1.w, t1.w), p.cap2);
        sts128(orow + ((c ^ ((row >> 1) & 3)) << 4), p0, p1, p2, p3);   // SWIZZLE_64B
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
#ifndef ST_EXP_NOSTORE
      if (leader) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_out),
                     "r"(sO + buf * ST_O_BYTES), "r"(0), "r"((int)(t * 128))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
#endif
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (lane == 0) {
    // ======================= MMA issuer =======================
    const uint64_t descB = make_smem_desc(sB);
    int i = 0;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int buf = i & 1, k = i >> 1;
      if (k > 0) mbar_wait(tmem_free + 8 * buf, (uint32_t)(k - 1) & 1u);   // accumulator drained by the epilogue
      mbar_wait(a_full + 8 * buf, (uint32_t)k & 1u);                        // A tile written
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t descA = make_smem_desc(sA + buf * ST_A_BYTES);
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * ST_C);
      umma_f16(tmem_d, descA, descB, 0u);            // k = 0..15
      umma_f16(tmem_d, descA + 2, descB + 2, 1u);    // k = 16..31
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mma_done + 8 * buf) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

}  // namespace

// Host-side preparation of the fp16 filter bank and folded shift for a given input transform.
// wq_dev: [32][32] fp16, shift2_dev: [32] fp32 (both device, caller-allocated).
cudaError_t stem_tc_prepare(const float* w_oihw_host, const float* scale_host, const float* shift_host,
                            float in_scale, float in_bias, __half* wq_dev, float* shift2_dev, float* p0_out,
                            float* shift2_host_out) {
  if (in_scale == 0.f) return cudaErrorInvalidValue;
  const float p0 = -in_bias / in_scale;
  const float p0h = __half2float(__float2half_rn(p0));
  __half wq[32 * 32];
  float shift2[32];
  for (int o = 0; o < 32; ++o) {
    double sum = 0.0;
    for (int k = 0; k < 32; ++k) {
      const float v = k < 27 ? w_oihw_host[o * 27 + k] * in_scale : 0.f;
      wq[o * 32 + k] = __float2half_rn(v);
      sum += (double)__half2float(wq[o * 32 + k]);
    }
    const double sc = scale_host ? scale_host[o] : 1.0, sh = shift_host ? shift_host[o] : 0.0;
    shift2[o] = (float)(sh + sc * (-(double)p0h * sum));
  }
  cudaError_t e = cudaMemcpy(wq_dev, wq, sizeof wq, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(shift2_dev, shift2, sizeof shift2, cudaMemcpyHostToDevice);
  if (p0_out) *p0_out = p0h;
  if (shift2_host_out) for (int o = 0; o < 32; ++o) shift2_host_out[o] = shift2[o];
  return e;
}

cudaError_t launch_stem_tc(bf16* out, const StemArgs& a, const __half* wq_dev, const float* scale_dev,
                           const float* shift2_dev, float p0, int act, int num_sms, cudaStream_t st,
                           std::string* err) {
  if (a.cout != ST_C || (a.stride != 1 && a.stride != 2)) return cudaErrorNotSupported;
  if (a.n <= 0) return cudaSuccess;
  if ((long)a.n * (a.rows / a.stride) * (a.cols / a.stride) >= (1L << 31)) return cudaErrorNotSupported;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  StemTcParams p{};
  p.r = a.r; p.g = a.g; p.b = a.b; p.pix_stride = a.pix_stride; p.img_stride = a.img_stride;
  p.n = a.n; p.rows = a.rows; p.cols = a.cols; p.orows = a.rows / a.stride; p.ocols = a.cols / a.stride;
  p.pad_lo = a.pad_lo; p.wq = wq_dev; p.scale = scale_dev; p.shift2 = shift2_dev;
  p.cap2 = act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  const unsigned short ph = __half_as_ushort(__float2half_rn(p0));
  p.pad_f16x2 = (uint32_t)ph | ((uint32_t)ph << 16);
  p.m_total = (long)a.n * p.orows * p.ocols;
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)ST_C, (cuuint64_t)p.m_total};
  cuuint64_t gstr[1] = {(cuuint64_t)ST_C * 2};
  cuuint32_t box[2] = {(cuuint32_t)ST_C, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) { char b[128]; snprintf(b, sizeof b, "stem output tensor map encode failed (CUresult %d)", (int)r); *err = b; }
    return cudaErrorInvalidValue;
  }
  const size_t smem = 1024 + 2 * ST_A_BYTES + ST_B_BYTES + 2 * ST_O_BYTES + 64;
  const long tiles = (p.m_total + 127) / 128;
  long grid = (long)num_sms * 3;
  if (grid > tiles) grid = tiles;
  const bool relu = act != MNV1_ACT_NONE;
  {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f) { if (e == cudaSuccess) e = ensure_dyn_smem((const void*)f, (int)smem); };
    set((const void*)stem_tc_kernel<1, true, true>); set((const void*)stem_tc_kernel<1, false, true>);
    set((const void*)stem_tc_kernel<2, true, true>); set((const void*)stem_tc_kernel<2, false, true>);
    set((const void*)stem_tc_kernel<1, true, false>); set((const void*)stem_tc_kernel<1, false, false>);
    set((const void*)stem_tc_kernel<2, true, false>); set((const void*)stem_tc_kernel<2, false, false>);
    if (e != cudaSuccess) return e;
  }
  // interleaved = one base pointer with g = r + 1, b = r + 2 and pixel stride 3
  const bool il = a.pix_stride == 3 && a.g == a.r + 1 && a.b == a.r + 2;
  if (!il && a.pix_stride != 1) return cudaErrorNotSupported;
  cudaError_t le = cudaSuccess;
#define ST_LAUNCH(S, R)                                                                          \
  do {                                                                                            \
    if (il) le = launch_pdl(stem_tc_kernel<S, R, true>, dim3((unsigned)grid), dim3(ST_THREADS), smem, st, tm, p);  \
    else    le = launch_pdl(stem_tc_kernel<S, R, false>, dim3((unsigned)grid), dim3(ST_THREADS), smem, st, tm, p); \
  } while (0)
  if (a.stride == 2) { if (relu) ST_LAUNCH(2, true); else ST_LAUNCH(2, false); }
  else               { if (relu) ST_LAUNCH(1, true); else ST_LAUNCH(1, false); }
#undef ST_LAUNCH
  return le;
}

}  // namespace mnv1
