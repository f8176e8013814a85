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
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// bytes (sat_u8(a), sat_u8(b), sat_u8(c), sat_u8(d)), a in the low byte: two cvt.pack.sat.u8.s32
__device__ __forceinline__ uint32_t pack4_sat_u8(int a, int b, int c, int d) {
  uint32_t hi, r;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
  return r;
}
__device__ __forceinline__ int finish(int acc, int bias, int relu, int rshift) {
  int v = acc + bias;
  if (relu) v = max(v, 0);
  return v >> rshift;                      // arithmetic shift: floor((acc + bias) / 2^rshift)
}

// ------------------------------------------------------------------------------------------ pointwise / FC
constexpr int I8_BM = 128, I8_STAGES = 4, I8_THREADS = 320;   // TMA producer, MMA issuer, 8 epilogue warps

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
  int res;               // the CTA's filter tile (all k-blocks) stays in shared memory; CTAs are bound to an n-tile
};

// out[M][Cout] = store_u8(finish(in[M][K] . w[Cout][K]^T)).  Persistent CTAs loop over (128-row tile, BN-column tile)
// units (n-tile innermost, so the A rows of a unit are re-read from L2):
//   warp 0   TMA producer: 4-stage ring of [128 x 128 B] A and [BN x 128 B] B tiles, running ahead across units; rows /
//            columns / K past the tensor are zero-filled by the TMA unit
//   warp 1   TMEM allocator + single-thread MMA issuer (UMMA 128 x BN x 32, u8 x s8 -> s32) into TWO accumulator stages,
//            so the epilogue of unit i overlaps the MMAs of unit i+1
//   warps 2-5 epilogue: tcgen05.ld of their 32 lanes, integer finish, 4 bytes per word, 128-bit stores
// KB = bytes (= u8 elements) of K per k-block and shared-memory row: 128 with the 128B swizzle, or — for the first layers,
// whose whole contraction is 32 or 64 channels — 64 / 32 with the matching narrower swizzle, so that no zero-filled
// columns are staged and multiplied (a [128 x 128 B] box over a 32-byte-wide tensor cost 3.6 k cycles per tile).
template <int KB> __device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
  constexpr uint64_t layout = KB == 128 ? 2 : KB == 64 ? 4 : 6;            // SWIZZLE_128B / _64B / _32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * KB) >> 4) << 32;                                      // SBO: 8 rows
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
template <int BN, int KB>
__global__ void __launch_bounds__(I8_THREADS) pointwise_i8_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b,
                                                                  const I8Params p) {
  constexpr int I8_BK = KB;
  constexpr uint32_t A_BYTES = I8_BM * I8_BK, B_BYTES = BN * I8_BK, STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t ACC = BN < 32 ? 32 : BN, TM_COLS = 2 * ACC;
  constexpr int EPI_HALVES = BN >= 64 ? 2 : 1;           // two epilogue warps per TMEM lane quarter, each takes half the columns
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + I8_BK - 1) / I8_BK;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const long m_tiles = (p.M + I8_BM - 1) / I8_BM;
  const long units = m_tiles * n_tiles;
  // Shared memory: p.res = 0: a ring of (A, B) k-block pairs.  p.res = 1: the CTA works on ONE n-tile for its whole life
  // (CTA c: n-tile c % n_tiles, m-tiles c / n_tiles, + gridDim.x / n_tiles, ...), its filter tile — all k-blocks —
  // is loaded once and the ring carries A only: a third of the L2 -> SM traffic of a 512 -> 512 layer, and half the
  // TMA issues (a thread starts a bulk copy every ~380 cycles).
  const bool res = p.res != 0;
  const uint32_t ring_bytes = res ? (uint32_t)I8_STAGES * A_BYTES + (uint32_t)num_kb * B_BYTES : (uint32_t)I8_STAGES * STAGE;
  const uint32_t bars = smem + ring_bytes;
  const uint32_t full = bars, empty = bars + 8u * I8_STAGES, tm_full = empty + 8u * I8_STAGES, tm_empty = tm_full + 16,
                 b_full = tm_empty + 16, tmem_slot = b_full + 16;   // sBias stays 16-byte aligned
  const uint32_t sBias = tmem_slot + 16;                    // [n_tiles * BN] s32, zero past Cout: the epilogue reads it with broadcast LDS.128
  auto a_addr = [&](int s) { return smem + (uint32_t)s * (res ? A_BYTES : STAGE); };
  auto b_addr = [&](int s, int kb) { return res ? smem + (uint32_t)I8_STAGES * A_BYTES + (uint32_t)kb * B_BYTES : smem + (uint32_t)s * STAGE + A_BYTES; };
  // this CTA's it-th unit
  const long my_first = res ? blockIdx.x / n_tiles : blockIdx.x, my_step = res ? gridDim.x / n_tiles : gridDim.x;
  const long my_end = res ? m_tiles : units;
  const int my_n = res ? (int)(blockIdx.x % n_tiles) * BN : 0;
  auto unit_m = [&](long u) { return (int)(res ? u : u / n_tiles) * I8_BM; };
  auto unit_n = [&](long u) { return res ? my_n : (int)(u % n_tiles) * BN; };
  if (p.bias) {
    int* sb = reinterpret_cast<int*>(smem_raw + (sBias - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < n_tiles * BN; i += I8_THREADS) sb[i] = i < p.Cout ? __ldg(p.bias + i) : 0;
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b);
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(full + 8u * s, 1); mbar_init(empty + 8u * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tm_full + 8u * a, 1); mbar_init(tm_empty + 8u * a, 4 * EPI_HALVES); }
    mbar_init(b_full, 1);
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      if (res && my_first < my_end) {
        mbar_expect_tx(b_full, (uint32_t)num_kb * B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(b_addr(0, kb), &tmap_b, b_full, kb * I8_BK, my_n);
      }
      int s = 0; uint32_t ph = 0;
      for (long u = my_first; u < my_end; u += my_step) {
        const int m_idx = unit_m(u), n_idx = unit_n(u);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty + 8u * s, ph ^ 1u);
          mbar_expect_tx(full + 8u * s, res ? A_BYTES : STAGE);
          tma_load_2d(a_addr(s), &tmap_a, full + 8u * s, kb * I8_BK, m_idx);
          if (!res) tma_load_2d(b_addr(s, kb), &tmap_b, full + 8u * s, kb * I8_BK, n_idx);
          if (++s == I8_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_u8s8_m128(BN);
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      if (res && my_first < my_end) mbar_wait(b_full, 0u);
      for (long u = my_first; u < my_end; u += my_step) {
        mbar_wait(tm_empty + 8u * as, aph ^ 1u);          // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * ACC;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full + 8u * s, ph);
          tc_fence_after();
          const uint64_t da = umma_desc_kmajor<KB>(a_addr(s));
          const uint64_t db = umma_desc_kmajor<KB>(b_addr(s, kb));
#pragma unroll
          for (int k = 0; k < I8_BK / 32; ++k) umma_i8(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(empty + 8u * s);
          if (++s == I8_STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(tm_full + 8u * as);
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else if ((warp - 2) / 4 < EPI_HALVES) {
    // two warps per TMEM lane quarter (warp % 4): warps 2..5 take the first half of the tile's columns, 6..9 the second —
    // one warp per scheduler could not keep up with the tensor pipe (every instruction waited out its own latency)
    const int quarter = warp & 3, half = (warp - 2) / 4;
    constexpr int C_LO_STEP = BN / EPI_HALVES < 32 ? 32 : BN / EPI_HALVES;
    const int c_lo = half * C_LO_STEP, c_hi = c_lo + C_LO_STEP < BN ? c_lo + C_LO_STEP : (BN < 32 ? 32 : BN);
    const bool vec = (p.Cout & 31) == 0;                    // 32-byte rows pieces: one 256-bit store = one full sector
    int as = 0; uint32_t aph = 0;
    for (long u = my_first; u < my_end; u += my_step) {
      const int m_idx = unit_m(u), n_idx = unit_n(u);
      const long row = (long)m_idx + quarter * 32 + lane;
      mbar_wait(tm_full + 8u * as, aph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)as * ACC + (uint32_t)c0, v);
        tmem_ld_wait();
        const int col0 = n_idx + c0;
        int x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = (int)v[j];
        if (p.bias) {                                         // uniform; staged in shared memory, zero past Cout
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int4 b4;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(b4.x), "=r"(b4.y), "=r"(b4.z), "=r"(b4.w) : "r"(sBias + (uint32_t)(col0 + 4 * j) * 4u));
            x[4 * j] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
          }
        }
        uint32_t q[8];
        if (!p.wrap) {
          // saturating store: sat_u8(x >> s) — the ReLU is implied (a negative value stays negative under the arithmetic
          // shift and saturates to 0); two cvt.pack.sat per four bytes
          if (p.rshift) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] >>= p.rshift;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) q[j] = pack4_sat_u8(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        } else {
          // kernel.cl's store: ReLU, shift, then the C conversion to unsigned char (low byte)
#pragma unroll
          for (int j = 0; j < 32; ++j) { if (p.relu) x[j] = max(x[j], 0); x[j] >>= p.rshift; }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            q[j] = prmt(prmt((uint32_t)x[4 * j], (uint32_t)x[4 * j + 1], 0x0040), prmt((uint32_t)x[4 * j + 2], (uint32_t)x[4 * j + 3], 0x0040), 0x5410);
        }
        if (row < p.M) {
          uint8_t* o = p.out + row * p.Cout + col0;
          if (vec && col0 + 32 <= p.Cout) {
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]),
                         "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Cout) o[j] = (uint8_t)(q[j >> 2] >> (8 * (j & 3)));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tm_empty + 8u * as);
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

cudaError_t encode_u8(CUtensorMap* map, const void* base, uint64_t rows, uint64_t k, uint32_t box_rows, int kb, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  cuuint64_t gdim[2] = {k, rows};
  cuuint64_t gstride[1] = {k};
  cuuint32_t box[2] = {(cuuint32_t)kb, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = kb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : kb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { if (err) *err = "pointwise_i8: tensor map encode failed"; return cudaErrorInvalidValue; }
  return cudaSuccess;
}

template <int BN, int KB>
cudaError_t launch_i8(const CUtensorMap& ta, const CUtensorMap& tb, I8Params p, int num_sms, cudaStream_t st) {
  constexpr int I8_BK = KB;
  const int n_tiles = (p.Cout + BN - 1) / BN, num_kb = (p.K + KB - 1) / KB;
  const long units = ((p.M + I8_BM - 1) / I8_BM) * n_tiles;
  const size_t tail = 16 * I8_STAGES + 96 + (size_t)n_tiles * BN * 4;
  const size_t smem_ring = 1024 + (size_t)I8_STAGES * (I8_BM * I8_BK + BN * I8_BK) + tail;
  const size_t smem_res = 1024 + (size_t)I8_STAGES * I8_BM * I8_BK + (size_t)num_kb * BN * I8_BK + tail;
  // resident filter tile: when it fits and there are CTAs for every n-tile
  p.res = smem_res <= 227 * 1024 && (long)num_sms >= n_tiles && units >= n_tiles ? 1 : 0;
  const size_t smem = p.res ? smem_res : smem_ring;
  cudaError_t e = ensure_dyn_smem((const void*)pointwise_i8_kernel<BN, KB>, 227 * 1024);
  if (e != cudaSuccess) return e;
  long per_sm = (227 * 1024) / (long)smem < 1 ? 1 : (227 * 1024) / (long)smem;     // resident CTAs per SM (shared memory)
  const long tm = 512 / (2 * (BN < 32 ? 32 : BN));                                   // ... and by TMEM columns
  if (per_sm > tm) per_sm = tm;
  if (per_sm > 2) per_sm = 2;                                                        // 320 threads each
  long grid = (long)num_sms * per_sm;
  if (grid > units) grid = units;
  if (p.res) grid -= grid % n_tiles;
  pointwise_i8_kernel<BN, KB><<<(unsigned)grid, I8_THREADS, smem, st>>>(ta, tb, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ depthwise
// NHWC u8.  A thread owns 4 adjacent output pixels x 4 channels (one 32-bit word per pixel).  DP4A is a 4-way dot
// product over the BYTES of a word, and a 3x3 depthwise tap row is a 3-way dot product over adjacent PIXELS of one
// channel — so the thread transposes its input words (pixel-major: 4 channels of one pixel) into channel-major words
// (4 adjacent pixels of one channel, PRMT), forms the 3-pixel window of every output with a funnel shift (the 4th byte
// meets a zero tap) and spends ONE dp4a.u32.s32 per (output, tap row): 3 instead of 9 multiply-adds per output.
// rows32: [C][3] words, word = (w[ty][0], w[ty][1], w[ty][2], 0) as s8 — built once per filter (api.cu).
// 4 words (pixels p0..p3, bytes = channels) -> 4 words (channels, bytes = pixels p0..p3)
__device__ __forceinline__ void transpose4(const uint32_t (&w)[4], uint32_t (&c)[4]) {
  const uint32_t t0 = prmt(w[0], w[1], 0x5140), t1 = prmt(w[2], w[3], 0x5140);   // (a.b0 b.b0 a.b1 b.b1)
  const uint32_t t2 = prmt(w[0], w[1], 0x7362), t3 = prmt(w[2], w[3], 0x7362);   // (a.b2 b.b2 a.b3 b.b3)
  c[0] = prmt(t0, t1, 0x5410); c[1] = prmt(t0, t1, 0x7632);
  c[2] = prmt(t2, t3, 0x5410); c[3] = prmt(t2, t3, 0x7632);
}
// The thread walks DOWN a band of RB output rows: every input row is loaded, transposed and windowed once and feeds the
// (up to) three output rows it belongs to, whose accumulators live in a ring of 3 (stride 1) / 2 (stride 2) slots —
// the row loop is fully unrolled, so the slots are compile-time registers.
template <int S>
__global__ void __launch_bounds__(256, 2) depthwise_u8_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                                           const int* __restrict__ rows32, const int* __restrict__ bias,
                                                           int n, int H, int W, int C, int pad_lo, int relu, int rshift,
                                                           int wrap) {
  constexpr int RB = 7;                       // output rows per thread (every map of the schedule is a multiple of 7 high)
  constexpr int HR = (RB - 1) * S + 3;        // input rows of a band
  constexpr int RING = S == 1 ? 3 : 2;
  constexpr int NPX = 3 * S + 3;              // input pixels a strip of 4 outputs needs: 6 (stride 1) or 9 (stride 2)
  constexpr int NW = (NPX + 3) / 4;           // channel-major words per channel: 2 or 3
  const int Ho = H / S, Wo = W / S, C4 = C >> 2, strips = (Wo + 3) >> 2, bands = (Ho + RB - 1) / RB;
  const long total = (long)n * bands * strips * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long r = i / C4;
    const int strip = (int)(r % strips); r /= strips;
    const int band = (int)(r % bands);
    const int img = (int)(r / bands);
    const int x0 = strip * 4, y0 = band * RB;
    int taps[4][3], bs[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) taps[b][ty] = __ldg(rows32 + (long)(4 * c4 + b) * 3 + ty);
      bs[b] = bias ? __ldg(bias + 4 * c4 + b) : 0;
    }
    int acc[RING][4][4];                      // [slot][channel][output pixel]
    // The band's input rows are requested PF rows ahead of the row being folded in: with one row at a time every thread
    // sat out a full L2 / HBM round trip per row (ncu: 3.4 warps stalled on the long scoreboard per issued instruction
    // at 16 warps per SM).  A row outside the image is the zero padding: its words are zero and contribute nothing.
    constexpr int PF = S == 1 ? 2 : 1;         // stride 2 reads 9 words per row: a second row in flight would spill
    uint32_t buf[PF + 1][NW * 4];
    auto load_row = [&](int q, uint32_t (&px)[NW * 4]) {
      const int yy = y0 * S + q - pad_lo;
      const bool ok = yy >= 0 && yy < H;
      const uint32_t* row = reinterpret_cast<const uint32_t*>(in + ((long)img * H + (ok ? yy : 0)) * W * C) + c4;
#pragma unroll
      for (int k = 0; k < NW * 4; ++k) {
        const int xx = x0 * S + k - pad_lo;
        px[k] = (ok && k < NPX && xx >= 0 && xx < W) ? __ldg(row + (long)xx * C4) : 0u;
      }
    };
#pragma unroll
    for (int q = 0; q < PF; ++q) load_row(q, buf[q]);
#pragma unroll
    for (int q = 0; q < HR; ++q) {
      if (q % S == 0 && q / S < RB) {         // output row q / S starts with this input row: fresh accumulators (+ bias)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[(q / S) % RING][b][j] = bs[b];
      }
      if (q + PF < HR) load_row(q + PF, buf[(q + PF) % (PF + 1)]);
      {
        const uint32_t (&px)[NW * 4] = buf[q % (PF + 1)];
        uint32_t ch[NW][4];                   // [group of 4 pixels][channel]
#pragma unroll
        for (int g = 0; g < NW; ++g) {
          const uint32_t w4[4] = {px[4 * g], px[4 * g + 1], px[4 * g + 2], px[4 * g + 3]};
          transpose4(w4, ch[g]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint32_t win[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int start = j * S;          // first input pixel of output j's window
            const uint32_t lo = ch[start >> 2][b], hi = ch[(start >> 2) + 1 < NW ? (start >> 2) + 1 : NW - 1][b];
            win[j] = (start & 3) ? __funnelshift_r(lo, hi, 8 * (start & 3)) : lo;
          }
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {    // input row q is tap row ty of output row (q - ty) / S
            if (q - ty >= 0 && (q - ty) % S == 0 && (q - ty) / S < RB) {
              const int slot = ((q - ty) / S) % RING;
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[slot][b][j] = dp4a_u8s8(win[j], taps[b][ty], acc[slot][b][j]);
            }
          }
        }
      }
      if (q - 2 >= 0 && (q - 2) % S == 0 && (q - 2) / S < RB) {     // output row (q - 2) / S is complete
        const int o = (q - 2) / S, slot = o % RING, y = y0 + o;
        if (y < Ho) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (x0 + j >= Wo) break;
            int v[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) v[b] = acc[slot][b][j];
            uint32_t w;
            if (!wrap) {                      // sat_u8(v >> s): the ReLU is implied by the saturation at 0
#pragma unroll
              for (int b = 0; b < 4; ++b) v[b] >>= rshift;
              w = pack4_sat_u8(v[0], v[1], v[2], v[3]);
            } else {                          // kernel.cl: ReLU, then the C conversion to unsigned char
#pragma unroll
              for (int b = 0; b < 4; ++b) { if (relu) v[b] = max(v[b], 0); v[b] >>= rshift; }
              w = prmt(prmt((uint32_t)v[0], (uint32_t)v[1], 0x0040), prmt((uint32_t)v[2], (uint32_t)v[3], 0x0040), 0x5410);
            }
            reinterpret_cast<uint32_t*>(out + (((long)img * Ho + y) * Wo + x0 + j) * C)[c4] = w;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ depthwise, tiled
// The same arithmetic with the band of input rows staged in shared memory.  The kernel above takes every input word
// straight from global memory: 16 warps per SM (128 registers for the row ring and the words in flight) cannot cover the
// round trip, and a third of its instructions are bounds checks and 64-bit addresses (ncu: issue slots 46 %, 3.4 warps
// stalled on the long scoreboard per issued instruction).  Here one CTA owns a tile of RB output rows x 4 STRIPS output
// columns x CB channels: all threads copy the (RB-1) S + 3 input rows with cp.async (zero-filled outside the image: the
// padding), then a thread (4 output pixels x 4 channels, as above) walks down the rows reading its words with
// immediate-offset shared loads.  The pixel pitch is padded so that the strips of a warp fall into different banks
// (CB = 32: 10 words, CB = 64: 18 words at stride 2; CB >= 128: a warp is one strip and reads one pixel's channels).
// Three CTAs per SM overlap one tile's copy with the arithmetic of the others.
template <int S_, int RB_, int CB_, int STRIPS_>
struct DtCfg {
  static constexpr int S = S_, RB = RB_, CB = CB_, STRIPS = STRIPS_;
  static constexpr int LANES = CB / 4;                 // threads across the channels of a pixel
  static constexpr int THREADS = STRIPS * LANES;
  static constexpr int TWO = STRIPS * 4;               // output columns of a tile
  static constexpr int HR = (RB - 1) * S + 3;          // input rows of a tile
  static constexpr int BW = (TWO - 1) * S + 3;         // input columns of a tile
  static constexpr int PITCH = CB == 32 ? 10 : CB == 64 ? 18 : CB / 4;   // words per pixel in shared memory
  static constexpr int ROWP = BW * PITCH;              // words per row
  static constexpr int UNIT = CB >= 128 ? 16 : 8;      // bytes per cp.async
  static constexpr int UPP = CB / UNIT;                // copies per pixel
  static constexpr int UNITS = HR * BW * UPP;
  static constexpr size_t SMEM = (size_t)HR * ROWP * 4;
  static_assert(CB % 32 == 0 && (CB >= 128 || (CB == 32 && S == 1) || (CB == 64 && S == 2)), "pitch table");
  static_assert(THREADS % 32 == 0 && THREADS <= 256, "whole warps, at most 256 threads");
};

template <typename Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::S == 1 ? 2 : 3) depthwise_u8_tile_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                                                          const int* __restrict__ rows32, const int* __restrict__ bias,
                                                                          int H, int W, int C, int pad_lo, int relu, int rshift,
                                                                          int wrap, int bands, int col_tiles, int slabs) {
  constexpr int S = Cfg::S, RB = Cfg::RB, HR = Cfg::HR, RING = S == 1 ? 3 : 2, NPX = 3 * S + 3, NW = (NPX + 3) / 4;
  extern __shared__ __align__(16) uint32_t s_tile[];
  const int Ho = H / S, Wo = W / S;
  int t = blockIdx.x;
  const int slab = t % slabs; t /= slabs;
  const int ct = t % col_tiles; t /= col_tiles;
  const int band = t % bands;
  const int img = t / bands;
  const int cb0 = slab * Cfg::CB, x0t = ct * Cfg::TWO, y0 = band * RB;
  // ---- copy the tile's input rows (all threads), zero-filled outside the image
  {
    const int iy0 = y0 * S - pad_lo, ix0 = x0t * S - pad_lo;
    const uint8_t* src0 = in + (long)img * H * W * C + cb0;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_tile);
    for (int u = threadIdx.x; u < Cfg::UNITS; u += Cfg::THREADS) {
      const int cu = u % Cfg::UPP, pr = u / Cfg::UPP, px = pr % Cfg::BW, row = pr / Cfg::BW;
      const int yy = iy0 + row, xx = ix0 + px;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const uint8_t* src = src0 + ((long)(ok ? yy : 0) * W + (ok ? xx : 0)) * C + cu * Cfg::UNIT;
      const uint32_t dst = dst0 + (uint32_t)(row * Cfg::ROWP + px * Cfg::PITCH) * 4u + (uint32_t)(cu * Cfg::UNIT);
      const int bytes = ok ? Cfg::UNIT : 0;       // src-size 0: the destination is filled with zeros
      if (Cfg::UNIT == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // ---- this thread's strip and channels; the taps travel while the copy is in flight
  const int c4l = threadIdx.x % Cfg::LANES, strip = threadIdx.x / Cfg::LANES;
  const int c4 = (cb0 >> 2) + c4l, x0 = x0t + strip * 4;
  int taps[4][3], bs[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
#pragma unroll
    for (int ty = 0; ty < 3; ++ty) taps[b][ty] = __ldg(rows32 + (long)(4 * c4 + b) * 3 + ty);
    bs[b] = bias ? __ldg(bias + 4 * c4 + b) : 0;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint32_t* tp = s_tile + strip * 4 * S * Cfg::PITCH + c4l;
  const int C4 = C >> 2;
  int acc[RING][4][4];                        // [slot][channel][output pixel]
#pragma unroll
  for (int q = 0; q < HR; ++q) {
    if (q % S == 0 && q / S < RB) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[(q / S) % RING][b][j] = bs[b];
    }
    uint32_t px[NW * 4];
#pragma unroll
    for (int k = 0; k < NW * 4; ++k) px[k] = k < NPX ? tp[q * Cfg::ROWP + k * Cfg::PITCH] : 0u;
    uint32_t ch[NW][4];                       // [group of 4 pixels][channel]
#pragma unroll
    for (int g = 0; g < NW; ++g) {
      const uint32_t w4[4] = {px[4 * g], px[4 * g + 1], px[4 * g + 2], px[4 * g + 3]};
      transpose4(w4, ch[g]);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t win[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int start = j * S;
        const uint32_t lo = ch[start >> 2][b], hi = ch[(start >> 2) + 1 < NW ? (start >> 2) + 1 : NW - 1][b];
        win[j] = (start & 3) ? __funnelshift_r(lo, hi, 8 * (start & 3)) : lo;
      }
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        if (q - ty >= 0 && (q - ty) % S == 0 && (q - ty) / S < RB) {
          const int slot = ((q - ty) / S) % RING;
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[slot][b][j] = dp4a_u8s8(win[j], taps[b][ty], acc[slot][b][j]);
        }
      }
    }
    if (q - 2 >= 0 && (q - 2) % S == 0 && (q - 2) / S < RB) {       // output row (q - 2) / S is complete
      const int o = (q - 2) / S, slot = o % RING, y = y0 + o;
      if (y < Ho) {
        uint32_t* orow = reinterpret_cast<uint32_t*>(out + ((long)img * Ho + y) * Wo * C) + c4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (x0 + j >= Wo) break;
          int v[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) v[b] = acc[slot][b][j];
          uint32_t w;
          if (!wrap) {
#pragma unroll
            for (int b = 0; b < 4; ++b) v[b] >>= rshift;
            w = pack4_sat_u8(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) { if (relu) v[b] = max(v[b], 0); v[b] >>= rshift; }
            w = prmt(prmt((uint32_t)v[0], (uint32_t)v[1], 0x0040), prmt((uint32_t)v[2], (uint32_t)v[3], 0x0040), 0x5410);
          }
          orow[(long)(x0 + j) * C4] = w;
        }
      }
    }
  }
}

template <typename Cfg>
cudaError_t launch_dw_tile(uint8_t* out, const uint8_t* in, const mnv1_filter* f, int n, int rows, int cols, int c, int pad_lo,
                           int wrap, cudaStream_t st) {
  const int Ho = rows / Cfg::S, Wo = cols / Cfg::S;
  const int bands = (Ho + Cfg::RB - 1) / Cfg::RB, col_tiles = (Wo + Cfg::TWO - 1) / Cfg::TWO, slabs = c / Cfg::CB;
  const long grid = (long)n * bands * col_tiles * slabs;
  if (grid <= 0 || grid >= (1L << 31)) return cudaErrorNotSupported;
  cudaError_t e = ensure_dyn_smem((const void*)depthwise_u8_tile_kernel<Cfg>, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  depthwise_u8_tile_kernel<Cfg><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, st>>>(
      out, in, f->w_q32, f->bias_i32, rows, cols, c, pad_lo, f->act != MNV1_ACT_NONE ? 1 : 0, f->rshift, wrap, bands, col_tiles, slabs);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ stem
// A thread owns one output pixel and all 32 filters.  Its 27 input bytes are packed into 7 words in the order
// k = 9*ty + 3*tx + plane — the order in which an interleaved RGB row delivers them: the 9 window bytes of a row are
// contiguous, so each row costs three aligned 32-bit loads and a funnel shift instead of nine byte loads.  Every filter's
// 27 s8 values are packed in the same order (api.cu), so one filter costs 7 DP4A.  Planar inputs (the reference's
// three-plane calling convention) place their bytes in the same order one at a time.  wq: [32][7] words.
// The packed filter bank and the bias travel as a kernel parameter: the DP4As take them straight from the constant bank
// (from shared memory every DP4A needed a load of its own and the kernel spent half its issue slots on them).
struct StemU8Consts { int w[32 * 7]; int b[32]; };
__global__ void __launch_bounds__(128) stem_u8_kernel(uint8_t* __restrict__ out, const __grid_constant__ StemArgs a,
                                                      const __grid_constant__ StemU8Consts cw, int relu, int rshift, int wrap) {
  const int Ho = a.rows / a.stride, Wo = a.cols / a.stride;
  const long total = (long)a.n * Ho * Wo;
  const bool interleaved = a.pix_stride == 3 && a.g == a.r + 1 && a.b == a.r + 2 && (a.cols * 3) % 4 == 0 && a.img_stride % 4 == 0 &&
                           (reinterpret_cast<uintptr_t>(a.r) & 3) == 0;
  const int rowbytes = a.cols * 3;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int img = (int)(i / ((long)Wo * Ho));
    uint32_t rw[3][3];                                   // the 9 window bytes of each row: two words + one byte
    if (interleaved) {
      const int start = (x * a.stride - a.pad_lo) * 3;   // may be negative (left padding) or run past the row (right padding)
      const int al = start & ~3, sh = 8 * (start & 3);
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        const int yy = y * a.stride + ty - a.pad_lo;
        uint32_t w[3] = {0u, 0u, 0u};
        if (yy >= 0 && yy < a.rows) {
          const uint8_t* row = a.r + (long)img * a.img_stride + (long)yy * rowbytes;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int off = al + 4 * k;
            if (off >= 0 && off + 4 <= rowbytes) w[k] = __ldg(reinterpret_cast<const uint32_t*>(row + off));
          }
        }
        rw[ty][0] = __funnelshift_r(w[0], w[1], sh);
        rw[ty][1] = __funnelshift_r(w[1], w[2], sh);
        rw[ty][2] = (w[2] >> sh) & 0xffu;
      }
    } else {
      const uint8_t* planes[3] = {a.r + (long)img * a.img_stride, a.g + (long)img * a.img_stride, a.b + (long)img * a.img_stride};
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        rw[ty][0] = rw[ty][1] = rw[ty][2] = 0u;
#pragma unroll
        for (int tx = 0; tx < 3; ++tx)
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            const int yy = y * a.stride + ty - a.pad_lo, xx = x * a.stride + tx - a.pad_lo;
            uint32_t v = 0;
            if (yy >= 0 && yy < a.rows && xx >= 0 && xx < a.cols) v = planes[pl][((long)yy * a.cols + xx) * a.pix_stride];
            const int k = 3 * tx + pl;
            rw[ty][k >> 2] |= v << (8 * (k & 3));
          }
      }
    }
    // 27 bytes -> 7 words: row 0 at byte 0, row 1 at byte 9, row 2 at byte 18
    uint32_t pk[7];
    pk[0] = rw[0][0];
    pk[1] = rw[0][1];
    pk[2] = rw[0][2] | (rw[1][0] << 8);                                      // r0[8], r1[0..2]
    pk[3] = __funnelshift_r(rw[1][0], rw[1][1], 24);                         // r1[3..6]
    pk[4] = (rw[1][1] >> 24) | (rw[1][2] << 8) | (rw[2][0] << 16);           // r1[7], r1[8], r2[0..1]
    pk[5] = __funnelshift_r(rw[2][0], rw[2][1], 16);                         // r2[2..5]
    pk[6] = (rw[2][1] >> 16) | (rw[2][2] << 16);                             // r2[6..7], r2[8], 0
    uint32_t* o = reinterpret_cast<uint32_t*>(out + i * 32);
    uint32_t q[8];
#pragma unroll
    for (int f4 = 0; f4 < 8; ++f4) {
      int v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int f = 4 * f4 + b;
        int acc = cw.b[f];
#pragma unroll
        for (int t = 0; t < 7; ++t) acc = dp4a_u8s8(pk[t], cw.w[f * 7 + t], acc);
        v[b] = acc;
      }
      if (!wrap) {
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] >>= rshift;
        q[f4] = pack4_sat_u8(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) { if (relu) v[b] = max(v[b], 0); v[b] >>= rshift; }
        q[f4] = prmt(prmt((uint32_t)v[0], (uint32_t)v[1], 0x0040), prmt((uint32_t)v[2], (uint32_t)v[3], 0x0040), 0x5410);
      }
    }
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]),
                 "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ stem on the tensor cores
// DP4A issues at a quarter of the FP32 rate on this part (one warp instruction every four cycles per scheduler), so the
// 224 DP4As per output pixel above cost 80 us per 256 images whatever else the kernel does.  This is stem_rows.cu for
// bytes: a tile is one output row of one image; a producer thread lands the tile's three input rows (contiguous in the
// interleaved image) in a ring slot with one bulk copy; the gather threads (thread = output column) read the nine window
// bytes of each row with three aligned shared loads + a funnel shift and write them AS THEY ARE into the 32-byte-swizzled
// K-major A operand (k = 8 row + j for window bytes j < 8, k = 24 + row for byte 8: no byte ever crosses a word; the
// filter bank is permuted to match); ONE tcgen05.mma kind::i8 (M 128, N 32, K 32) per tile; the epilogue warps add the
// bias, shift, saturate and store 32 bytes per pixel with one 256-bit store.  Zero padding: missing rows are tile-uniform,
// missing columns touch the first / last thread.
constexpr int SI_THREADS = 320;   // 4 gather + 4 epilogue + MMA/TMEM + producer warps
constexpr int SI_NI = 8;          // input ring depth
constexpr int SI_NA = 4;          // A tiles / accumulators in flight
constexpr int SI_LEAD = 16;       // bytes before row 0 in a slot (column -1 of the REF padding)
constexpr uint32_t SI_A_BYTES = 128 * 32, SI_B_BYTES = 32 * 32;
struct StemI8Params {
  const uint8_t* img;    // interleaved RGB, n x rows x cols x 3
  uint8_t* out;          // n x orows x ocols x 32
  long img_stride;
  int rows, cols, orows, ocols, pad_lo, tiles, slot_bytes, relu, rshift, wrap;
  uint32_t b[32][8];     // filter bank: [filter][k / 4] words of 4 s8 in the A operand's k order
  int bias[32];
};

__global__ void __launch_bounds__(SI_THREADS, 3) stem_i8_rows_kernel(const __grid_constant__ StemI8Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem;                               // NA x 4 KB
  const uint32_t sB = smem + SI_NA * SI_A_BYTES;          // 1 KB
  const uint32_t bars = sB + SI_B_BYTES;                  // a_full[NA] mma_done[NA] tmem_free[NA] in_full[NI] in_empty[NI]
  const uint32_t a_full = bars, mma_done = a_full + 8 * SI_NA, tmem_free = mma_done + 8 * SI_NA,
                 in_full = tmem_free + 8 * SI_NA, in_empty = in_full + 8 * SI_NI;
  const uint32_t tmem_slot = in_empty + 8 * SI_NI;
  const uint32_t sIn = bars + 256;                        // NI slots of slot_bytes
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 32) {                                         // B tile: filter `tid`, 32-byte swizzle (16-byte chunk ^= bit 2 of the row)
    const uint32_t row = sB + (uint32_t)tid * 32u, sw = (uint32_t)(tid >> 2) & 1u;
    sts128(row + ((0u ^ sw) << 4), p.b[tid][0], p.b[tid][1], p.b[tid][2], p.b[tid][3]);
    sts128(row + ((1u ^ sw) << 4), p.b[tid][4], p.b[tid][5], p.b[tid][6], p.b[tid][7]);
  }
  if (tid == 0) {
    for (int b = 0; b < SI_NA; ++b) { mbar_init(a_full + 8 * b, 128); mbar_init(mma_done + 8 * b, 1); mbar_init(tmem_free + 8 * b, 4); }
    for (int s = 0; s < SI_NI; ++s) { mbar_init(in_full + 8 * s, 1); mbar_init(in_empty + 8 * s, 4); }
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)(SI_NA * 32));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  const int Wo = p.ocols, Ho = p.orows, H = p.rows, RB = p.cols * 3;
  const int tiles = p.tiles;

  if (warp < 4) {
    // ======================= gather warps (thread = output column) =======================
    const int ox = tid;
    const int start = SI_LEAD + 6 * ox - 3 * p.pad_lo;      // byte offset of the window in a slot row
    const uint32_t sh8 = (uint32_t)(start & 3) * 8;
    const uint32_t col_off = (uint32_t)(start & ~3);
    const bool pad_left = 2 * ox - p.pad_lo < 0, pad_right = 2 * ox - p.pad_lo + 2 >= p.cols;
    const int oy_step = (int)(gridDim.x % (unsigned)Ho);
    int oy = (int)(blockIdx.x % (unsigned)Ho);
    const uint32_t sw = (uint32_t)(tid >> 2) & 1u;
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int slot = i % SI_NI, kin = i / SI_NI, buf = i % SI_NA, k = i / SI_NA;
      mbar_wait(in_full + 8 * slot, (uint32_t)kin & 1u);
      uint32_t a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = 0u;
      if (ox < Wo) {
        const uint32_t base = sIn + (uint32_t)slot * (uint32_t)p.slot_bytes + col_off;
        const int iy0 = 2 * oy - p.pad_lo;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (iy0 + r < 0 || iy0 + r >= H) continue;        // a row outside the image is the zero padding (tile-uniform)
          const uint32_t w0 = lds32(base + r * RB), w1 = lds32(base + r * RB + 4), w2 = lds32(base + r * RB + 8);
          uint32_t b0 = __funnelshift_r(w0, w1, sh8), b1 = __funnelshift_r(w1, w2, sh8), e8 = (w2 >> sh8) & 0xffu;
          if (pad_left) b0 &= 0xff000000u;                  // window bytes 0..2 = column -1
          if (pad_right) { b1 &= 0x0000ffffu; e8 = 0u; }    // window bytes 6..8 = column `cols`
          a[2 * r] = b0; a[2 * r + 1] = b1;
          a[6] |= e8 << (8 * r);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(in_empty + 8 * slot);      // the slot's bytes are in registers
      oy += oy_step;
      if (oy >= Ho) oy -= Ho;
      if (k > 0) mbar_wait(mma_done + 8 * buf, (uint32_t)(k - 1) & 1u);   // A[buf] was read by the MMA of tile i - NA
      const uint32_t arow = sA + (uint32_t)buf * SI_A_BYTES + (uint32_t)tid * 32u;
      sts128(arow + ((0u ^ sw) << 4), a[0], a[1], a[2], a[3]);
      sts128(arow + ((1u ^ sw) << 4), a[4], a[5], a[6], a[7]);
      fence_proxy_async();
      mbar_arrive(a_full + 8 * buf);
    }
  } else if (warp < 8) {
    // ======================= epilogue warps =======================
    const int q = warp & 3, col = q * 32 + lane;
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int buf = i % SI_NA, k = i / SI_NA;
      mbar_wait(mma_done + 8 * buf, (uint32_t)k & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_free + 8 * buf);
      if (col >= Wo) continue;
      int x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = (int)v[j] + p.bias[j];
      uint32_t w[8];
      if (!p.wrap) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] >>= p.rshift;
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pack4_sat_u8(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) { if (p.relu) x[j] = max(x[j], 0); x[j] >>= p.rshift; }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          w[j] = prmt(prmt((uint32_t)x[4 * j], (uint32_t)x[4 * j + 1], 0x0040), prmt((uint32_t)x[4 * j + 2], (uint32_t)x[4 * j + 3], 0x0040), 0x5410);
      }
      uint8_t* o = p.out + ((long)t * Wo + col) * 32;
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                   "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
  } else if (warp == 8) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_u8s8_m128(32);
      const uint64_t descB = umma_desc_kmajor<32>(sB);
      int i = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
        const int buf = i % SI_NA, k = i / SI_NA;
        if (k > 0) mbar_wait(tmem_free + 8 * buf, (uint32_t)(k - 1) & 1u);
        mbar_wait(a_full + 8 * buf, (uint32_t)k & 1u);
        tc_fence_after();
        umma_i8(tmem_base + (uint32_t)(buf * 32), umma_desc_kmajor<32>(sA + (uint32_t)buf * SI_A_BYTES), descB, idesc, 0u);
        umma_commit(mma_done + 8 * buf);
      }
    }
  } else if (lane == 0) {
    // ======================= producer: one bulk copy (the tile's input rows) per tile ==========
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int slot = i % SI_NI, kin = i / SI_NI;
      if (kin > 0) mbar_wait(in_empty + 8 * slot, (uint32_t)(kin - 1) & 1u);
      const int img = t / Ho, oy = t - img * Ho;
      const int iy0 = 2 * oy - p.pad_lo;
      const int r_lo = iy0 < 0 ? 0 : iy0, r_hi = iy0 + 3 > H ? H : iy0 + 3;
      const uint32_t bytes = (uint32_t)((r_hi - r_lo) * RB);
      const uint32_t dst = sIn + (uint32_t)slot * (uint32_t)p.slot_bytes + SI_LEAD + (uint32_t)((r_lo - iy0) * RB);
      mbar_expect_tx(in_full + 8 * slot, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(p.img + (long)img * p.img_stride + (long)r_lo * RB), "r"(bytes), "r"(in_full + 8 * slot) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(SI_NA * 32));
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
// Four channels per thread (c % 4 == 0, hw <= 257): one 32-bit load per pixel, the four byte sums carried in two words of
// two 16-bit lanes each (hw * 255 < 65536), all of a thread's loads of a batch of 7 pixels in flight at once.
__global__ void __launch_bounds__(128) pool_u8_vec_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in, int n, int hw,
                                                          int c, int wrap) {
  const int c4n = c >> 2;
  const long total = (long)n * c4n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long img = i / c4n;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(in + img * hw * c) + c4;
    uint32_t s02 = 0u, s13 = 0u;                 // (ch0, ch2) and (ch1, ch3) as 16-bit lanes
    int px = 0;
    for (; px + 7 <= hw; px += 7) {
      uint32_t w[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) w[k] = __ldg(p + (long)(px + k) * c4n);
#pragma unroll
      for (int k = 0; k < 7; ++k) { s02 += w[k] & 0x00ff00ffu; s13 += (w[k] >> 8) & 0x00ff00ffu; }
    }
    for (; px < hw; ++px) { const uint32_t w = __ldg(p + (long)px * c4n); s02 += w & 0x00ff00ffu; s13 += (w >> 8) & 0x00ff00ffu; }
    const int s0 = (int)(s02 & 0xffffu), s2 = (int)(s02 >> 16), s1 = (int)(s13 & 0xffffu), s3 = (int)(s13 >> 16);
    const uint32_t r = store_u8(s0 / hw, wrap) | (store_u8(s1 / hw, wrap) << 8) | (store_u8(s2 / hw, wrap) << 16) |
                       (store_u8(s3 / hw, wrap) << 24);
    reinterpret_cast<uint32_t*>(out + img * c)[c4] = r;
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
                                int num_sms, cudaStream_t st, std::string* err) {
  if (!f->w_s8 || k % 16 || k < 16) { if (err) *err = "pointwise (u8): Cin must be a multiple of 16"; return cudaErrorNotSupported; }
  if (m <= 0) return cudaSuccess;
  int bn = cout >= 256 ? 256 : cout > 64 ? 128 : cout > 32 ? 64 : cout > 16 ? 32 : 16;
  // few rows (the FC layer: M = batch): narrower column tiles spread the contraction over more SMs
  if (bn > 64 && ((m + I8_BM - 1) / I8_BM) * ((cout + bn - 1) / bn) * 2 < num_sms) bn = 64;
  const int kb = k == 32 ? 32 : k == 64 ? 64 : 128;
  CUtensorMap ta, tb;
  cudaError_t e = encode_u8(&ta, in, (uint64_t)m, (uint64_t)k, I8_BM, kb, err);
  if (e == cudaSuccess) e = encode_u8(&tb, f->w_s8, (uint64_t)cout, (uint64_t)k, (uint32_t)bn, kb, err);
  if (e != cudaSuccess) return e;
  I8Params p{out, f->bias_i32, m, k, cout, f->act != MNV1_ACT_NONE ? 1 : 0, f->rshift, wrap};
  if (kb == 32) {          // layer 3 of the schedule (32 -> 64)
    if (bn == 64) return launch_i8<64, 32>(ta, tb, p, num_sms, st);
    if (bn == 128) return launch_i8<128, 32>(ta, tb, p, num_sms, st);
  } else if (kb == 64) {   // layer 5 (64 -> 128)
    if (bn == 128) return launch_i8<128, 64>(ta, tb, p, num_sms, st);
    if (bn == 64) return launch_i8<64, 64>(ta, tb, p, num_sms, st);
  }
  if (kb != 128) {         // other shapes with a narrow contraction: 128-byte boxes, zero-filled by the TMA unit
    e = encode_u8(&ta, in, (uint64_t)m, (uint64_t)k, I8_BM, 128, err);
    if (e == cudaSuccess) e = encode_u8(&tb, f->w_s8, (uint64_t)cout, (uint64_t)k, (uint32_t)bn, 128, err);
    if (e != cudaSuccess) return e;
  }
  switch (bn) {
    case 256: return launch_i8<256, 128>(ta, tb, p, num_sms, st);
    case 128: return launch_i8<128, 128>(ta, tb, p, num_sms, st);
    case 64: return launch_i8<64, 128>(ta, tb, p, num_sms, st);
    case 32: return launch_i8<32, 128>(ta, tb, p, num_sms, st);
    default: return launch_i8<16, 128>(ta, tb, p, num_sms, st);
  }
}

cudaError_t launch_depthwise_u8(uint8_t* out, const uint8_t* in, const mnv1_filter* f, int n, int rows, int cols, int stride,
                                int c, int pad_lo, int wrap, cudaStream_t st) {
  if (c % 4 || !f->w_q32) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
  // 128 channels and more go through shared memory (32 / 64 channels measured faster on the direct kernel: 89 vs 99 us on
  // layer 2, 78 vs 82 us on layer 4 — four strips of a warp share the copy's 8-byte granularity and the padded pitch)
  if ((reinterpret_cast<uintptr_t>(in) & 15) == 0 && rows % stride == 0 && cols % stride == 0) {
    const int wo = cols / stride;
    if (stride == 1) {
      if (c % 512 == 0 && wo <= 8) return launch_dw_tile<DtCfg<1, 7, 512, 2>>(out, in, f, n, rows, cols, c, pad_lo, wrap, st);
      if (c % 256 == 0 && wo <= 16) return launch_dw_tile<DtCfg<1, 7, 256, 4>>(out, in, f, n, rows, cols, c, pad_lo, wrap, st);
      if (c % 128 == 0) return launch_dw_tile<DtCfg<1, 7, 128, 7>>(out, in, f, n, rows, cols, c, pad_lo, wrap, st);
    } else if (stride == 2) {
      if (c % 512 == 0 && wo <= 8) return launch_dw_tile<DtCfg<2, 4, 512, 2>>(out, in, f, n, rows, cols, c, pad_lo, wrap, st);
      if (c % 128 == 0 && wo <= 16) return launch_dw_tile<DtCfg<2, 4, 128, 4>>(out, in, f, n, rows, cols, c, pad_lo, wrap, st);
      if (c % 128 == 0) return launch_dw_tile<DtCfg<2, 4, 128, 7>>(out, in, f, n, rows, cols, c, pad_lo, wrap, st);
    }
  }
  const long total = (long)n * ((rows / stride + 6) / 7) * ((cols / stride + 3) / 4) * (c / 4);
  const int relu = f->act != MNV1_ACT_NONE ? 1 : 0;
  if (stride == 1)
    depthwise_u8_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(out, in, f->w_q32, f->bias_i32, n, rows, cols, c, pad_lo, relu, f->rshift, wrap);
  else
    depthwise_u8_kernel<2><<<grid_for(total, 256), 256, 0, st>>>(out, in, f->w_q32, f->bias_i32, n, rows, cols, c, pad_lo, relu, f->rshift, wrap);
  return cudaGetLastError();
}

cudaError_t launch_stem_u8(uint8_t* out, const StemArgs& a, const mnv1_filter* f, int wrap, int num_sms, cudaStream_t st,
                           const char** kernel_name) {
  if (a.cout != 32 || !f->w_q32) return cudaErrorNotSupported;
  if (a.n <= 0) return cudaSuccess;
  const long total = (long)a.n * (a.rows / a.stride) * (a.cols / a.stride);
  if (f->h_q32.size() != 32 * 7) return cudaErrorNotSupported;
  const bool il = a.pix_stride == 3 && a.g == a.r + 1 && a.b == a.r + 2;
  const int rb = a.cols * 3;
  if (il && a.stride == 2 && !(a.rows & 1) && !(a.cols & 1) && a.cols / 2 <= 128 && rb % 16 == 0 && a.img_stride % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(a.r) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0 && f->rshift < 32 &&
      (long)a.n * (a.rows / 2) < (1L << 31)) {
    StemI8Params p{};
    p.img = a.r; p.out = out; p.img_stride = a.img_stride; p.rows = a.rows; p.cols = a.cols; p.orows = a.rows / 2; p.ocols = a.cols / 2;
    p.pad_lo = a.pad_lo; p.tiles = a.n * p.orows; p.relu = f->act != MNV1_ACT_NONE ? 1 : 0; p.rshift = f->rshift; p.wrap = wrap;
    p.slot_bytes = (SI_LEAD + 3 * rb + 16 + 127) & ~127;
    // h_q32 holds filter o's 27 s8 values at byte 9 ty + 3 tx + plane; the A operand's order is k = 8 ty + j (j = 3 tx + plane < 8),
    // k = 24 + ty for j = 8
    for (int o = 0; o < 32; ++o) {
      for (int w = 0; w < 8; ++w) p.b[o][w] = 0u;
      for (int ty = 0; ty < 3; ++ty)
        for (int j = 0; j < 9; ++j) {
          const int src = 9 * ty + j, k = j < 8 ? 8 * ty + j : 24 + ty;
          const uint32_t v = ((uint32_t)f->h_q32[o * 7 + (src >> 2)] >> (8 * (src & 3))) & 0xffu;
          p.b[o][k >> 2] |= v << (8 * (k & 3));
        }
      p.bias[o] = f->h_bias.empty() ? 0 : f->h_bias[o];
    }
    const size_t smem = 1024 + SI_NA * SI_A_BYTES + SI_B_BYTES + 256 + (size_t)SI_NI * p.slot_bytes;
    if (smem <= 75 * 1024) {   // 3 CTAs per SM
      cudaError_t e = ensure_dyn_smem((const void*)stem_i8_rows_kernel, 75 * 1024);
      if (e != cudaSuccess) return e;
      long grid = (long)num_sms * 3;
      if (grid > p.tiles) grid = p.tiles;
      if (kernel_name) *kernel_name = "stem_i8_rows_kernel";
      stem_i8_rows_kernel<<<(unsigned)grid, SI_THREADS, smem, st>>>(p);
      return cudaGetLastError();
    }
  }
  StemU8Consts k;
  for (int i = 0; i < 32 * 7; ++i) k.w[i] = f->h_q32[i];
  for (int i = 0; i < 32; ++i) k.b[i] = f->h_bias.empty() ? 0 : f->h_bias[i];
  stem_u8_kernel<<<grid_for(total, 128), 128, 0, st>>>(out, a, k, f->act != MNV1_ACT_NONE ? 1 : 0, f->rshift, wrap);
  return cudaGetLastError();
}

cudaError_t launch_pool_u8(uint8_t* out, const uint8_t* in, int n, int hw, int c, int wrap, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (c % 4 == 0 && hw <= 257 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3) == 0)
    pool_u8_vec_kernel<<<grid_for((long)n * (c / 4), 128), 128, 0, st>>>(out, in, n, hw, c, wrap);
  else
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
