// stem_rows.cu — layer 1 (3x3/2 conv, interleaved RGB u8 -> 32 maps) with the image rows staged
// in shared memory by bulk copies.
//
// Same contract and arithmetic as stem_tc.cu (`convolute`, kernel.cl:2-60 on the tensor cores:
// exact u8 -> fp16 widening, fp16 filter bank, fp32 accumulation, folded input transform); what
// changes is how the 27 taps reach the A operand.  stem_tc.cu has every thread issue 27 byte
// loads from global memory per output pixel and wait for them, so the layer is bound by load
// latency (0.38 of its HBM roofline).  Here one tile is ONE output row of one image:
//   * producer (1 thread): the three input rows a tile needs are contiguous in the interleaved
//     image (3 x cols x 3 bytes = 2016 B at 224), so one cp.async.bulk per tile lands them in a
//     ring slot and signals an mbarrier; the ring runs NI tiles ahead of the gatherers;
//   * gather (128 threads, thread = output column): 3 aligned 32-bit shared loads per input row
//     cover the 9 window bytes, a funnel shift realigns them, PRMT + HSUB2 widen pairs exactly
//     (0x6400|b = 1024+b in fp16).  The K order of the GEMM is chosen so that pairs never straddle
//     words: k = 8*row + j for window bytes j < 8, k = 24 + row for byte 8; the filter bank is
//     permuted to match when the B tile is built;
//   * MMA as in stem_tc.cu (M = 128 with Wo valid rows, N = 32, K = 32, NA accumulators in flight), the
//     epilogue on packed FFMA2 with scale/shift in the constant bank; a pixel's 32 channels are 64
//     contiguous bytes of the NHWC map, so every thread stores its pixel with two 256-bit stores
//     (no staging in shared memory, no proxy fence, no TMA store on the warp's critical path:
//     65.1 vs 66.7 us).  Tried and not kept: the producer folded into the first gather warp, 9 warps
//     and four CTAs per SM (67.2 vs 66.3 us on the same box).
// Padding: missing rows (tile-uniform) and columns (first/last thread) are replaced by p0 after the
// widening, as stem_tc.cu does.
#include <cuda_fp16.h>

#include <cstdio>

#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

constexpr int SR_THREADS = 320;   // 4 gather + 4 epilogue + MMA/TMEM + producer warps
constexpr int SR_C = 32;
constexpr int SR_NI = 8;          // input ring depth (tiles in flight per CTA)
constexpr int SR_NA = 4;          // A tiles / accumulators in flight (two tiles share one 128B-swizzled 16 KB block)
constexpr uint32_t SR_A_BYTES = 128 * 128;
constexpr uint32_t SR_B_BYTES = 32 * 128;
constexpr int SR_LEAD = 16;       // bytes before row 0 in a slot (column -1 of the REF padding)

// kind::f16, D = f32, A = B = fp16, K-major, N = 32, M = 128
constexpr uint32_t SR_IDESC = (1u << 4) | ((uint32_t)(SR_C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(SR_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// (1024 + lo, 1024 + hi) as fp16 bit patterns -> exact (lo, hi)
__device__ __forceinline__ uint32_t unbias(uint32_t v, uint32_t bias) {
  __half2 h = __hsub2(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<__half2*>(&bias));
  return *reinterpret_cast<uint32_t*>(&h);
}

struct StemRowsParams {
  const uint8_t* img;    // interleaved RGB, n x rows x cols x 3
  long img_stride;
  int rows, cols, orows, ocols, pad_lo, tiles, slot_bytes;
  const __half* wq;      // [32][32] fp16 in stem_tc.cu's K order (plane, row, column)
  bf16* out;             // NHWC map
  float scale[SR_C], shift[SR_C];
  uint32_t cap2, pad_f16x2;
};

template <bool RELU>
__global__ void __launch_bounds__(SR_THREADS, 3)
stem_rows_kernel(const __grid_constant__ StemRowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem;                         // 2 x 16 KB
  const uint32_t sB = smem + 2 * SR_A_BYTES;        // 4 KB
  const uint32_t bars = sB + SR_B_BYTES;            // a_full[NA] mma_done[NA] tmem_free[NA] in_full[NI] in_empty[NI]
  const uint32_t a_full = bars, mma_done = a_full + 8 * SR_NA, tmem_free = mma_done + 8 * SR_NA,
                 in_full = tmem_free + 8 * SR_NA, in_empty = in_full + 8 * SR_NI;
  const uint32_t tmem_slot = in_empty + 8 * SR_NI;
  const uint32_t sIn = bars + 256;                  // NI slots of slot_bytes

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid < SR_C) {
    // B tile: thread t writes filter row t in the row-major-window K order, 128B-swizzled
    const __half* src = p.wq + tid * 32;
    uint32_t wv[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      uint32_t h2[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * q + e;
        int old = -1;
        if (k < 24) { const int r = k >> 3, j = k & 7; old = (j % 3) * 9 + r * 3 + j / 3; }
        else if (k < 27) old = 2 * 9 + (k - 24) * 3 + 2;
        h2[e] = old >= 0 ? (uint32_t)__half_as_ushort(__ldg(src + old)) : 0u;
      }
      wv[q] = h2[0] | (h2[1] << 16);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(sB + tid * 128 + ((c ^ (tid & 7)) << 4), wv[4 * c], wv[4 * c + 1], wv[4 * c + 2], wv[4 * c + 3]);
  }
  if (tid == 0) {
    for (int b = 0; b < SR_NA; ++b) { mbar_init(a_full + 8 * b, 128); mbar_init(mma_done + 8 * b, 1); mbar_init(tmem_free + 8 * b, 4); }
    for (int s = 0; s < SR_NI; ++s) { mbar_init(in_full + 8 * s, 1); mbar_init(in_empty + 8 * s, 4); }
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)(SR_NA * SR_C));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  pdl_wait();   // previous step's kernels are done with the activation arena; the images are in place

  const int Wo = p.ocols, Ho = p.orows, H = p.rows, RB = p.cols * 3;
  const int tiles = p.tiles;

  if (warp < 4) {
    // ======================= gather warps (thread = output column) =======================
    const int ox = tid;
    const int start = SR_LEAD + 6 * ox - 3 * p.pad_lo;        // byte offset of the window in a slot row
    const uint32_t sh8 = (uint32_t)(start & 3) * 8;
    const uint32_t col_off = (uint32_t)(start & ~3);
    const bool pad_left = 2 * ox - p.pad_lo < 0, pad_right = 2 * ox - p.pad_lo + 2 >= p.cols;
    const uint32_t pad2 = p.pad_f16x2;
    const uint32_t bias2 = 0x64006400u, bias1 = 0x00006400u;
    // row of the tile inside its image, stepped without a divide
    const int oy_step = (int)(gridDim.x % (unsigned)Ho);
    int oy = (int)(blockIdx.x % (unsigned)Ho);
    const int sw = tid & 7;
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int slot = i % SR_NI, kin = i / SR_NI, buf = i % SR_NA, k = i / SR_NA;
      mbar_wait(in_full + 8 * slot, (uint32_t)kin & 1u);
      uint32_t a[14];
      if (ox < Wo) {
        const uint32_t base = sIn + (uint32_t)slot * (uint32_t)p.slot_bytes + col_off;
        uint32_t e8[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint32_t w0 = lds32(base + r * RB), w1 = lds32(base + r * RB + 4), w2 = lds32(base + r * RB + 8);
          const uint32_t b0 = __funnelshift_r(w0, w1, sh8), b1 = __funnelshift_r(w1, w2, sh8);
          e8[r] = (w2 >> sh8) & 0xffu;
          a[4 * r + 0] = unbias(__byte_perm(b0, bias2, 0x5150), bias2);
          a[4 * r + 1] = unbias(__byte_perm(b0, bias2, 0x5352), bias2);
          a[4 * r + 2] = unbias(__byte_perm(b1, bias2, 0x5150), bias2);
          a[4 * r + 3] = unbias(__byte_perm(b1, bias2, 0x5352), bias2);
        }
        a[12] = unbias(e8[0] | (e8[1] << 16) | bias2, bias2);
        a[13] = unbias(e8[2] | bias1, bias1);
        // padding: rows outside the image are tile-uniform, columns touch the first / last thread
        const int iy0 = 2 * oy - p.pad_lo;
        const bool row_lo = iy0 < 0, row_hi = iy0 + 2 >= H;
        if (row_lo | row_hi | pad_left | pad_right) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const bool rp = (r == 0 && row_lo) || (r == 2 && row_hi);
            if (rp) { a[4 * r] = pad2; a[4 * r + 1] = pad2; a[4 * r + 2] = pad2; a[4 * r + 3] = pad2; }
            if (pad_left) { a[4 * r] = pad2; a[4 * r + 1] = (a[4 * r + 1] & 0xffff0000u) | (pad2 & 0xffffu); }
            if (pad_right) a[4 * r + 3] = pad2;
            if (rp || pad_right) {
              if (r == 0) a[12] = (a[12] & 0xffff0000u) | (pad2 & 0xffffu);
              if (r == 1) a[12] = (a[12] & 0x0000ffffu) | (pad2 & 0xffff0000u);
              if (r == 2) a[13] = pad2 & 0xffffu;
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 14; ++q) a[q] = 0u;
      }
      // the slot's bytes are in registers: hand it back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(in_empty + 8 * slot);
      oy += oy_step;
      if (oy >= Ho) oy -= Ho;
      // A[buf] was last read by the MMAs of tile i-NA.  Tile `buf` is the 64-byte half (buf & 1) of
      // the 128-byte rows of block buf >> 1: chunks 4h..4h+3 before the swizzle.
      if (k > 0) mbar_wait(mma_done + 8 * buf, (uint32_t)(k - 1) & 1u);
      const uint32_t arow = sA + (uint32_t)(buf >> 1) * SR_A_BYTES + tid * 128;
      const int c0 = (buf & 1) * 4;
      sts128(arow + (((c0 + 0) ^ sw) << 4), a[0], a[1], a[2], a[3]);
      sts128(arow + (((c0 + 1) ^ sw) << 4), a[4], a[5], a[6], a[7]);
      sts128(arow + (((c0 + 2) ^ sw) << 4), a[8], a[9], a[10], a[11]);
      sts128(arow + (((c0 + 3) ^ sw) << 4), a[12], a[13], 0u, 0u);
      fence_proxy_async();
      mbar_arrive(a_full + 8 * buf);
    }
  } else if (warp < 8) {
    // ======================= epilogue warps =======================
    // Each warp owns 32 pixels of the row and works on its own — no block-level barrier between the four warps.
    const int q = warp & 3, row = q * 32 + lane;
    const int vr = Wo - 32 * q < 0 ? 0 : (Wo - 32 * q > 32 ? 32 : Wo - 32 * q);   // valid pixels of this warp
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int buf = i % SR_NA, k = i / SR_NA;
      mbar_wait(mma_done + 8 * buf, (uint32_t)k & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * SR_C), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_free + 8 * buf);
      if (vr == 0) continue;
      uint32_t o[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int ch = 2 * e;
        const f32x2 acc = f2_pack(__uint_as_float(v[ch]), __uint_as_float(v[ch + 1]));
        o[e] = pack2_f2<RELU>(f2_fma(acc, f2_pack(p.scale[ch], p.scale[ch + 1]), f2_pack(p.shift[ch], p.shift[ch + 1])), p.cap2);
      }
      if (lane < vr) {
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + ((long)t * Wo + row) * 64;
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                     "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 32), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
                     "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]) : "memory");
      }
    }
  } else if (warp == 8) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint64_t descB = umma_desc_sw128(sB);
      int i = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
        const int buf = i % SR_NA, k = i / SR_NA;
        if (k > 0) mbar_wait(tmem_free + 8 * buf, (uint32_t)(k - 1) & 1u);
        mbar_wait(a_full + 8 * buf, (uint32_t)k & 1u);
        tc_fence_after();
        // K-slices of 32 bytes inside the 128-byte swizzle atom: tile half h starts at byte 64 h
        const uint64_t descA = umma_desc_sw128(sA + (uint32_t)(buf >> 1) * SR_A_BYTES) + (uint64_t)((buf & 1) * 4);
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * SR_C);
        umma_f16(tmem_d, descA, descB, 0u);            // k = 0..15
        umma_f16(tmem_d, descA + 2, descB + 2, 1u);    // k = 16..31
        umma_commit(mma_done + 8 * buf);
      }
    }
  } else if (lane == 0) {
    // ======================= producer: one bulk copy (the tile's input rows) per tile ==========
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int slot = i % SR_NI, kin = i / SR_NI;
      if (kin > 0) mbar_wait(in_empty + 8 * slot, (uint32_t)(kin - 1) & 1u);
      const int img = t / Ho, oy = t - img * Ho;
      const int iy0 = 2 * oy - p.pad_lo;
      const int r_lo = iy0 < 0 ? 0 : iy0, r_hi = iy0 + 3 > H ? H : iy0 + 3;
      const uint32_t bytes = (uint32_t)((r_hi - r_lo) * RB);
      const uint32_t dst = sIn + (uint32_t)slot * (uint32_t)p.slot_bytes + SR_LEAD + (uint32_t)((r_lo - iy0) * RB);
      mbar_expect_tx(in_full + 8 * slot, bytes);
      bulk_load(dst, p.img + (long)img * p.img_stride + (long)r_lo * RB, bytes, in_full + 8 * slot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(SR_NA * SR_C));
  }
}

}  // namespace

// Row-staged stem for interleaved RGB, stride 2.  scale_host / shift2_host: the 32 folded epilogue
// constants (host copies; they travel in the constant bank).  cudaErrorNotSupported = use stem_tc.
cudaError_t launch_stem_rows(bf16* out, const StemArgs& a, const __half* wq_dev, const float* scale_host,
                             const float* shift2_host, float p0, int act, int num_sms, cudaStream_t st,
                             std::string* err) {
  const bool il = a.pix_stride == 3 && a.g == a.r + 1 && a.b == a.r + 2;
  const int rb = a.cols * 3;
  if (!il || a.stride != 2 || a.cout != SR_C || (a.rows & 1) || (a.cols & 1) || a.cols / 2 > 128 || rb % 16 ||
      a.img_stride % 16 || (reinterpret_cast<uintptr_t>(a.r) & 15) || !shift2_host)
    return cudaErrorNotSupported;
  if (a.n <= 0) return cudaSuccess;
  const long m_total = (long)a.n * (a.rows / 2) * (a.cols / 2);
  if (m_total >= (1L << 31)) return cudaErrorNotSupported;
  if (reinterpret_cast<uintptr_t>(out) & 31) { if (err) *err = "stem output must be 32-byte aligned"; return cudaErrorNotSupported; }
  StemRowsParams p{};
  p.img = a.r; p.img_stride = a.img_stride; p.rows = a.rows; p.cols = a.cols; p.orows = a.rows / 2; p.ocols = a.cols / 2;
  p.pad_lo = a.pad_lo; p.tiles = a.n * p.orows; p.wq = wq_dev; p.out = out;
  p.slot_bytes = (SR_LEAD + 3 * rb + 16 + 127) & ~127;
  for (int c = 0; c < SR_C; ++c) { p.scale[c] = scale_host ? scale_host[c] : 1.f; p.shift[c] = shift2_host[c]; }
  p.cap2 = act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  const unsigned short ph = __half_as_ushort(__float2half_rn(p0));
  p.pad_f16x2 = (uint32_t)ph | ((uint32_t)ph << 16);
  const size_t smem = 1024 + 2 * SR_A_BYTES + SR_B_BYTES + 256 + (size_t)SR_NI * p.slot_bytes;
  if (smem > 75 * 1024) return cudaErrorNotSupported;   // 3 CTAs per SM
  long grid = (long)num_sms * 3;
  if (grid > p.tiles) grid = p.tiles;
  {
    cudaError_t e = ensure_dyn_smem((const void*)stem_rows_kernel<true>, 75 * 1024);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)stem_rows_kernel<false>, 75 * 1024);
    if (e != cudaSuccess) return e;
  }
  if (act != MNV1_ACT_NONE) return launch_pdl(stem_rows_kernel<true>, dim3((unsigned)grid), dim3(SR_THREADS), smem, st, p);
  return launch_pdl(stem_rows_kernel<false>, dim3((unsigned)grid), dim3(SR_THREADS), smem, st, p);
}

}  // namespace mnv1
