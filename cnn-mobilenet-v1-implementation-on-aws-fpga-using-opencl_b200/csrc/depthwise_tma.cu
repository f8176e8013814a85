// depthwise_tma.cu — the bandwidth path of the 3x3 depthwise stencil for bf16 contexts.
//
// Same contract as depthwise.cu (`depthwise`, kernel.cl:62-92) but built around TMA:
//   * the NHWC feature map is a 4-D tensor map (C, W, H, N); a producer warp streams halo
//     row-chunks [RC rows][(TWO-1)*S+3 cols][CB channels] into a ring of shared-memory stages
//     (mbarrier full/empty pairs).  Out-of-bounds rows/cols — the zero padding of the layer —
//     are filled by the TMA unit, so the stencil has no border branches at all.
//   * persistent CTAs: a CTA keeps ONE channel block (its 9 x 4 taps per thread stay in
//     registers for the whole kernel) and walks (image, column strip, row segment) items; the
//     bytes in flight per SM are stages x chunk x resident CTAs, independent of registers.
//   * a consumer thread owns 4 channels of TW adjacent output columns; every input element is
//     read from shared memory (ld.shared.v2) and widened to fp32 once per thread, then
//     accumulated input-row-major into a ring of 3 (stride 1) or 2 (stride 2) output-row
//     accumulators.  The folded-BN scale is pre-multiplied into the taps and the shift seeds the
//     accumulator; ReLU rides on the bf16 pack (cvt.rn.relu.bf16x2.f32), the 6-cap is one FMNMX.
//     At 4.5 FLOP/B the stencil is instruction-issue bound on B200 unless the loop is this lean
//     (profiles/r01_dw_v1_ncu.txt: 29 instr/element before, FFMA only 31 % of issue).
#include <cstdio>

#include "common.cuh"

namespace mnv1 {
namespace {

constexpr int DT_THREADS = 256;        // warp 0 = TMA producer, warps 1..7 = consumers
constexpr int DT_CONSUMERS = DT_THREADS - 32;
constexpr int DT_MAX_STAGES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware instead of burning issue slots
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DW_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DW_DONE;\n"
      "bra DW_WAIT;\n"
      "DW_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
// fp32 pair -> bf16x2 with ReLU folded into the convert and the upper clamp applied on the packed
// pair (min.bf16x2): rounding is monotonic and the cap is exactly representable, so this equals
// round(min(max(x, 0), cap)).
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi, uint32_t cap2) {
  uint32_t d;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  asm("min.bf16x2 %0, %0, %1;" : "+r"(d) : "r"(cap2));
  return d;
}
__device__ __forceinline__ void stg64_pred(void* ptr, uint32_t a, uint32_t b, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %3, 0;\n"
      "@p st.global.v2.u32 [%0], {%1, %2};\n"
      "}\n" ::"l"(ptr),
      "r"(a), "r"(b), "r"((uint32_t)pred)
      : "memory");
}

struct DwParams {
  bf16* out;
  const float* w9xC;   // taps, [9][C], already multiplied by the folded-BN scale
  const float* shift;  // [C] or nullptr
  float hi;            // upper clamp (6 for ReLU6, +inf otherwise)
  int n, H, W, C, Ho, Wo, pad_lo;
  int strips, cblocks, segs, seg_rows;  // work decomposition
  int chunks;                           // row-chunks per item
  int stages;
  int rest_total;                       // n * strips * segs : items per channel block
};

// S: stride; CB: channels per CTA; TWO: output columns per item; TW: output columns per thread;
// RC: input rows per chunk (a multiple of the accumulator-ring period: 3 for S=1, 4 for S=2).
template <int S, int CB, int TWO, int TW, int RC, bool RELU>
__global__ void __launch_bounds__(DT_THREADS, 2)
depthwise_tma_kernel(const __grid_constant__ CUtensorMap tmap_in, const DwParams p) {
  static_assert((S == 1 && RC % 3 == 0) || (S == 2 && RC % 4 == 0), "RC must be a multiple of the ring period");
  constexpr int BW = (TWO - 1) * S + 3;           // input columns per chunk
  constexpr int CQ = CB / 4;                      // channel quads (threads) per pixel
  constexpr int PG = TWO / TW;                    // pixel groups per item row
  constexpr int NCOL = (TW - 1) * S + 3;          // input columns a thread reads per row
  constexpr int RING = S == 1 ? 3 : 2;
  constexpr uint32_t STAGE_BYTES = (uint32_t)RC * BW * CB * 2;
  constexpr uint32_t STAGE_PITCH = (STAGE_BYTES + 1023u) & ~1023u;
  static_assert(TWO % TW == 0 && CQ * PG <= DT_CONSUMERS, "tile does not fit the consumer threads");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  const uint32_t bar_full = smem + (uint32_t)stages * STAGE_PITCH;   // stages x 8 bytes
  const uint32_t bar_empty = bar_full + 8u * DT_MAX_STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_in) : "memory");
    for (int s = 0; s < stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, DT_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();   // the previous layer's output is complete and visible

  // a CTA owns channel block cb and every (gridDim/cblocks)-th of the remaining items
  const int cblocks = p.cblocks;
  const int cb = blockIdx.x % cblocks;
  const int j0 = blockIdx.x / cblocks, jstride = gridDim.x / cblocks;
  const int segs = p.segs, strips = p.strips, seg_rows = p.seg_rows, chunks = p.chunks;
  const int rest_total = p.rest_total;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int pad_lo = p.pad_lo;
      for (int rest = j0; rest < rest_total; rest += jstride) {
        int t = rest;
        const int seg = t % segs; t /= segs;
        const int strip = t % strips;
        const int img = t / strips;
        const int x0 = strip * TWO * S - pad_lo;
        const int y0 = seg * seg_rows * S - pad_lo;
        for (int k = 0; k < chunks; ++k) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          mbar_expect_tx(bar_full + 8u * stage, STAGE_BYTES);
          tma_load_4d(smem + (uint32_t)stage * STAGE_PITCH, &tmap_in, bar_full + 8u * stage, cb * CB, x0, y0 + k * RC, img);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int ct = threadIdx.x - 32;
  const bool active = ct < CQ * PG;
  const int cq = active ? ct % CQ : 0;
  const int pg = active ? ct / CQ : 0;
  const int c0 = cb * CB + cq * 4;
  const int C = p.C, Wo = p.Wo, Ho = p.Ho;
  // upper clamp as a packed bf16 pair (+inf = 0x7f80 when there is no cap)
  const uint32_t cap2 = p.hi < 1e30f ? 0x40c040c0u : 0x7f807f80u;

  float w[9][4], sh[4];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.w9xC + (long)k * C + c0));
    w[k][0] = a.x; w[k][1] = a.y; w[k][2] = a.z; w[k][3] = a.w;
  }
  if (p.shift) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    sh[0] = a.x; sh[1] = a.y; sh[2] = a.z; sh[3] = a.w;
  } else {
    sh[0] = sh[1] = sh[2] = sh[3] = 0.f;
  }
  // byte offset of this thread's first column / channel quad inside a stage
  const uint32_t thr_off = (uint32_t)((pg * TW * S) * CB + cq * 4) * 2u;
  const long row_pitch_b = (long)Wo * C * 2;  // output bytes per row

  float acc[RING][TW][4];
#pragma unroll
  for (int a = 0; a < RING; ++a)
#pragma unroll
    for (int t = 0; t < TW; ++t)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[a][t][v] = 0.f;

  int stage = 0; uint32_t phase = 0;
  for (int rest = j0; rest < rest_total; rest += jstride) {
    int tt = rest;
    const int seg = tt % segs; tt /= segs;
    const int strip = tt % strips;
    const int img = tt / strips;
    const int ox = strip * TWO + pg * TW;
    bool col_ok[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) col_ok[t] = active && (ox + t) < Wo;
    uint8_t* out_col = reinterpret_cast<uint8_t*>(p.out + (((long)img * Ho + (long)seg * seg_rows) * Wo + ox) * C + c0);
    // S=1: input row q closes output row q-2; S=2: even q closes q/2-1.  `o` tracks that row.
    int o = S == 1 ? -2 : -1;

    for (int k = 0; k < chunks; ++k) {
      mbar_wait(bar_full + 8u * stage, phase);
      const uint32_t sbase = smem + (uint32_t)stage * STAGE_PITCH + thr_off;
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        // q = qk + r : input row relative to the item's first (padded) row
        float x[NCOL][4];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
          const uint2 raw = lds64(sbase + (uint32_t)((r * BW + j) * CB * 2));
          x[j][0] = bf16lo_to_f32(raw.x); x[j][1] = bf16hi_to_f32(raw.x);
          x[j][2] = bf16lo_to_f32(raw.y); x[j][3] = bf16hi_to_f32(raw.y);
        }
        int slot_done = -1;
        if (S == 1) {
          const int a0 = r % 3, a1 = (r + 2) % 3, a2 = (r + 1) % 3;  // output rows q, q-1, q-2 (RC % 3 == 0)
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc[a0][t][v] = fmaf(x[t + 2][v], w[2][v], fmaf(x[t + 1][v], w[1][v], fmaf(x[t][v], w[0][v], sh[v])));
              acc[a1][t][v] = fmaf(x[t + 2][v], w[5][v], fmaf(x[t + 1][v], w[4][v], fmaf(x[t][v], w[3][v], acc[a1][t][v])));
              acc[a2][t][v] = fmaf(x[t + 2][v], w[8][v], fmaf(x[t + 1][v], w[7][v], fmaf(x[t][v], w[6][v], acc[a2][t][v])));
            }
          slot_done = a2;
        } else {
          if ((r & 1) == 0) {  // q even: closes output row q/2-1 (tap row 2), opens q/2 (tap row 0); RC % 4 == 0
            const int a0 = (r / 2) % 2, a2 = (r / 2 + 1) % 2;
#pragma unroll
            for (int t = 0; t < TW; ++t)
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                acc[a2][t][v] = fmaf(x[2 * t + 2][v], w[8][v], fmaf(x[2 * t + 1][v], w[7][v], fmaf(x[2 * t][v], w[6][v], acc[a2][t][v])));
                acc[a0][t][v] = fmaf(x[2 * t + 2][v], w[2][v], fmaf(x[2 * t + 1][v], w[1][v], fmaf(x[2 * t][v], w[0][v], sh[v])));
              }
            slot_done = a2;
          } else {             // q odd: middle tap row of output row (q-1)/2
            const int a1 = (r / 2) % 2;
#pragma unroll
            for (int t = 0; t < TW; ++t)
#pragma unroll
              for (int v = 0; v < 4; ++v)
                acc[a1][t][v] = fmaf(x[2 * t + 2][v], w[5][v], fmaf(x[2 * t + 1][v], w[4][v], fmaf(x[2 * t][v], w[3][v], acc[a1][t][v])));
          }
        }
        if (slot_done >= 0) {  // compile-time: this input row closes an output row
          const bool row_ok = (unsigned)o < (unsigned)seg_rows;
          uint8_t* orow = out_col + (long)o * row_pitch_b;
#pragma unroll
          for (int t = 0; t < TW; ++t)
            stg64_pred(orow + (long)t * C * 2, pack2<RELU>(acc[slot_done][t][0], acc[slot_done][t][1], cap2),
                       pack2<RELU>(acc[slot_done][t][2], acc[slot_done][t][3], cap2), row_ok && col_ok[t]);
          ++o;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8u * stage);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
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

template <int S, int CB, int TWO, int TW, int RC, bool RELU>
cudaError_t launch_variant2(const bf16* in, DwParams p, int num_sms, cudaStream_t st, std::string* err) {
  constexpr int BW = (TWO - 1) * S + 3;
  constexpr uint32_t STAGE_BYTES = (uint32_t)RC * BW * CB * 2;
  constexpr uint32_t STAGE_PITCH = (STAGE_BYTES + 1023u) & ~1023u;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.n};
  cuuint64_t gstr[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
  cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)BW, (cuuint32_t)RC, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) { char b[128]; snprintf(b, sizeof b, "depthwise tensor map encode failed (CUresult %d)", (int)r); *err = b; }
    return cudaErrorInvalidValue;
  }
  p.strips = (p.Wo + TWO - 1) / TWO;
  p.cblocks = p.C / CB;
  // grid: 2 CTAs per SM, a multiple of the channel blocks (each CTA keeps one block's taps)
  long grid = (long)num_sms * 2;
  grid = grid / p.cblocks * p.cblocks;
  if (grid < p.cblocks) return cudaErrorNotSupported;
  const long per_cb = grid / p.cblocks;
  // row segments: pick the divisor of Ho that minimises the rows the busiest CTA walks
  // (waves x chunks x RC): tall segments re-read less halo, short ones balance better.
  int seg_rows = p.Ho;
  long best = -1;
  for (int sr = p.Ho; sr >= 1; --sr) {
    if (p.Ho % sr) continue;
    const long items = (long)p.n * p.strips * (p.Ho / sr);
    const long waves = (items + per_cb - 1) / per_cb;
    const long ch = ((sr - 1) * S + 3 + RC - 1) / RC;
    const long cost = waves * ch * RC;
    if (best < 0 || cost < best) { best = cost; seg_rows = sr; }
  }
  p.seg_rows = seg_rows;
  p.segs = p.Ho / seg_rows;
  p.chunks = ((seg_rows - 1) * S + 3 + RC - 1) / RC;
  p.rest_total = p.n * p.strips * p.segs;
  if (per_cb > p.rest_total) grid = (long)p.rest_total * p.cblocks;
  int stages = (int)((110 * 1024 - 1024 - 16 * DT_MAX_STAGES) / STAGE_PITCH);   // 2 CTAs/SM
  if (stages > DT_MAX_STAGES) stages = DT_MAX_STAGES;
  if (stages < 2) return cudaErrorNotSupported;
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * STAGE_PITCH + 16 * DT_MAX_STAGES;
  {
    cudaError_t e = ensure_dyn_smem((const void*)depthwise_tma_kernel<S, CB, TWO, TW, RC, RELU>, 112 * 1024);
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(depthwise_tma_kernel<S, CB, TWO, TW, RC, RELU>, dim3((unsigned)grid), dim3(DT_THREADS), smem, st, tm, p);
}

template <int S, int CB, int TWO, int TW, int RC>
cudaError_t launch_variant(const bf16* in, const DwParams& p, bool relu, int num_sms, cudaStream_t st,
                           std::string* err) {
  return relu ? launch_variant2<S, CB, TWO, TW, RC, true>(in, p, num_sms, st, err)
              : launch_variant2<S, CB, TWO, TW, RC, false>(in, p, num_sms, st, err);
}

}  // namespace

// Returns cudaErrorNotSupported (without launching) when no TMA variant fits the shape: the
// caller then uses the direct kernel of depthwise.cu.
cudaError_t launch_depthwise_tma(bf16* out, const bf16* in, const float* w9xC_scaled, const float* shift, int act,
                                 int n, int rows, int cols, int stride, int c, int pad_lo, int num_sms,
                                 cudaStream_t st, std::string* err) {
  DwParams p{};
  p.out = out; p.w9xC = w9xC_scaled; p.shift = shift;
  p.hi = act == MNV1_ACT_RELU6 ? 6.f : INFINITY;
  p.n = n; p.H = rows; p.W = cols; p.C = c; p.Ho = rows / stride; p.Wo = cols / stride; p.pad_lo = pad_lo;
  if (n <= 0) return cudaSuccess;
  const bool relu = act != MNV1_ACT_NONE;
  const int Wo = p.Wo;
  if (stride == 1) {
    // 7x7 and 14x14 maps: 9-row items (7 output rows + halo) = three 3-row chunks, nothing wasted
    if (c % 128 == 0 && Wo <= 7) return launch_variant<1, 128, 7, 1, 3>(in, p, relu, num_sms, st, err);
    if (c % 128 == 0 && Wo == 14) return launch_variant<1, 128, 14, 2, 3>(in, p, relu, num_sms, st, err);
    if (c % 64 == 0 && Wo % 28 == 0) return launch_variant<1, 64, 28, 2, 3>(in, p, relu, num_sms, st, err);
    if (c % 32 == 0 && Wo % 56 == 0) return launch_variant<1, 32, 56, 2, 6>(in, p, relu, num_sms, st, err);
  } else if (stride == 2) {
    if (c % 128 == 0 && Wo <= 7) return launch_variant<2, 128, 7, 1, 4>(in, p, relu, num_sms, st, err);
    if (c % 64 == 0 && Wo % 14 == 0) return launch_variant<2, 64, 14, 1, 4>(in, p, relu, num_sms, st, err);
  }
  return cudaErrorNotSupported;
}

}  // namespace mnv1
