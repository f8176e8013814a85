// depthwise_ring.cu — the 3x3 depthwise stencil for the 14x14x512 maps (layers 14-24 of the
// MobileNet.c schedule), stride 1 and 2.
//
// Same contract as depthwise_tma.cu (`depthwise`, kernel.cl:62-92) and bit-identical results (same
// tap order, fp32 accumulation).  At 14x14 the layer is not HBM-bound — the whole map sits in
// L2 — but instruction-issue bound, so this kernel is the stencil half of fused_rb.cu on its own:
//   * unit = (tile of R x TWO output pixels of one image, 64-channel block);
//   * one persistent CTA per SM, NG stencil groups of 4 warps, group g takes units g, g+NG, ...;
//   * a producer thread streams the halo row-chunks [RC][(TWO-1)*S+3][64] of every unit through a
//     4-D tensor map (out-of-bounds = the layer's zero padding) into a private ring per group, in
//     wavefront order (chunk k of the NG units of a round, then chunk k+1, ...);
//   * a thread keeps 4 channels x TW columns, fetches input row q+1 while it computes row q, skips
//     the partial sums that would fall outside the tile, keeps a ring of 3 (S=1) / 2 (S=2) output-row
//     accumulators and stores every finished row with one 8-byte st.global per column.
// No border branches, no shared-memory stores, ~13 instructions per output.
#include <cstdio>

#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

constexpr int DR_GW = 4;   // warps per stencil group

template <int S_, int C_, int H_, int TWO_, int R_, int TW_, int RC_, int NG_, int NIG_>
struct DrCfg {
  static constexpr int S = S_, C = C_, H = H_, TWO = TWO_, R = R_, TW = TW_, RC = RC_, NG = NG_, NIG = NIG_;
  static constexpr int HO = H / S, WO = H / S;
  static constexpr int NKB = C / 64;
  static constexpr int HR = (R - 1) * S + 3, BW = (TWO - 1) * S + 3, NCHK = (HR + RC - 1) / RC;
  static constexpr int PG = TWO / TW, NCOL = (TW - 1) * S + 3, RING = S == 1 ? 3 : 2;
  static constexpr int BANDS = HO / R, STRIPS = WO / TWO;
  static constexpr uint32_t CHUNK_BYTES = (uint32_t)RC * BW * 128;
  static constexpr int NI = NG * NIG;
  static constexpr int THREADS = (NG * DR_GW + 1) * 32;
  static constexpr uint32_t OFF_IN = 0;
  static constexpr uint32_t OFF_TAPS = OFF_IN + NI * CHUNK_BYTES;
  static constexpr uint32_t OFF_SH = OFF_TAPS + 9u * C * 4;
  static constexpr uint32_t OFF_BAR = OFF_SH + (uint32_t)C * 4;
  static constexpr size_t SMEM = 1024 + OFF_BAR + 16u * NI;
  static_assert(C % 64 == 0 && HO % R == 0 && WO % TWO == 0 && TWO % TW == 0, "shape does not tile");
  static_assert(PG * 16 <= DR_GW * 32, "tile does not fit a stencil group");
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
};

struct DrParams {
  bf16* out;
  const float* taps;    // [9][C] taps x folded-BN scale
  const float* shift;   // [C] or nullptr
  uint32_t cap2;
  int pad_lo;
  int units;            // n * BANDS * STRIPS * NKB
};

template <class Cfg, bool RELU>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
depthwise_ring_kernel(const __grid_constant__ CUtensorMap tmap_in, const DrParams p) {
  constexpr int S = Cfg::S, C = Cfg::C, TWO = Cfg::TWO, R = Cfg::R, TW = Cfg::TW, RC = Cfg::RC, NG = Cfg::NG, NIG = Cfg::NIG;
  constexpr int HR = Cfg::HR, BW = Cfg::BW, NCHK = Cfg::NCHK, PG = Cfg::PG, NCOL = Cfg::NCOL, RING = Cfg::RING;
  constexpr int NKB = Cfg::NKB, NI = Cfg::NI, WO = Cfg::WO, HO = Cfg::HO;
  constexpr int PER_IMG = Cfg::BANDS * Cfg::STRIPS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_g = smem_raw + (smem - smem_u32(smem_raw));
  const uint32_t sIn = smem + Cfg::OFF_IN, sTaps = smem + Cfg::OFF_TAPS, sSh = smem + Cfg::OFF_SH;
  const uint32_t in_full = smem + Cfg::OFF_BAR, in_empty = in_full + 8u * NI;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();

  {
    constexpr int PER = (9 * C / 4 + Cfg::THREADS - 1) / Cfg::THREADS;
    stage_constants<PER>(reinterpret_cast<float*>(smem_g + Cfg::OFF_TAPS), p.taps, 9 * C, tid, Cfg::THREADS);
    float* sh = reinterpret_cast<float*>(smem_g + Cfg::OFF_SH);
    if (p.shift) stage_constants<1>(sh, p.shift, C, tid, Cfg::THREADS);
    else for (int i = tid; i < C; i += Cfg::THREADS) sh[i] = 0.f;
  }
  if (tid == 0) {
    prefetch_tmap(&tmap_in);
    for (int s = 0; s < NI; ++s) { mbar_init(in_full + 8u * s, 1); mbar_init(in_empty + 8u * s, DR_GW); }
    mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();                                            // the previous layer's output is complete and visible

  // this CTA's units: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int G = gridDim.x;
  const int total_units = (p.units - (int)blockIdx.x + G - 1) / G;

  if (warp == NG * DR_GW) {
    // ======================= TMA producer (top warp: highest issue priority) =======================
    if (lane == 0) {
      int stage[NG]; uint32_t phase[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) { stage[g] = 0; phase[g] = 0; }
      for (int u0 = 0; u0 < total_units; u0 += NG) {
        const int nu = total_units - u0 < NG ? total_units - u0 : NG;
        int cx[NG], cy[NG], cc[NG], ci[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const uint32_t u = blockIdx.x + (uint32_t)(u0 + (g < nu ? g : 0)) * (uint32_t)G;
          const uint32_t tile = u / NKB, kb = u - tile * NKB;
          const uint32_t img = tile / PER_IMG, rem = tile - img * PER_IMG;
          const uint32_t band = rem / Cfg::STRIPS, strip = rem - band * Cfg::STRIPS;
          cx[g] = (int)strip * TWO * S - p.pad_lo; cy[g] = (int)band * R * S - p.pad_lo; cc[g] = (int)kb * 64; ci[g] = (int)img;
        }
#pragma unroll 1
        for (int k = 0; k < NCHK; ++k) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (g < nu) {
              const uint32_t st = (uint32_t)(g * NIG + stage[g]);
              mbar_wait(in_empty + 8u * st, phase[g] ^ 1u);
              mbar_expect_tx(in_full + 8u * st, Cfg::CHUNK_BYTES);
              tma_load_4d(sIn + st * Cfg::CHUNK_BYTES, &tmap_in, in_full + 8u * st, cc[g], cx[g], cy[g] + k * RC, ci[g]);
              if (++stage[g] == NIG) { stage[g] = 0; phase[g] ^= 1u; }
            }
          }
        }
      }
    }
    return;
  }

  // ======================= stencil groups =======================
  const int g = warp / DR_GW;
  const int t = tid - g * DR_GW * 32;
  const bool active = t < PG * 16;
  const int quad = t & 15;
  const int pg = active ? t >> 4 : PG - 1;
  const uint32_t in_off = (uint32_t)(pg * TW * S) * 128u + (uint32_t)quad * 8u;
  uint32_t rstage = 0, rphase = 0;

  for (int u0 = 0; u0 < total_units; u0 += NG) {
    const int nu = total_units - u0 < NG ? total_units - u0 : NG;
    if (g >= nu) break;
    const uint32_t u = blockIdx.x + (uint32_t)(u0 + g) * (uint32_t)G;
    const uint32_t tile = u / NKB, kb = u - tile * NKB;
    const uint32_t img = tile / PER_IMG, rem = tile - img * PER_IMG;
    const uint32_t band = rem / Cfg::STRIPS, strip = rem - band * Cfg::STRIPS;
    // this k-block's taps and shift for the thread's 4 channels
    f32x2 w[9][2], sh[2];                                   // the thread's 4 channels as two packed fp32 pairs
    {
      const uint32_t ch = (kb * 64u + (uint32_t)quad * 4u) * 4u;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 a = lds128f(sTaps + (uint32_t)k * C * 4u + ch);
        w[k][0] = f2_pack(a.x, a.y); w[k][1] = f2_pack(a.z, a.w);
      }
      const float4 a = lds128f(sSh + ch);
      sh[0] = f2_pack(a.x, a.y); sh[1] = f2_pack(a.z, a.w);
    }
    uint8_t* const obase = reinterpret_cast<uint8_t*>(p.out) +
        ((((size_t)img * HO + band * R) * WO + strip * TWO + (uint32_t)(pg * TW)) * C + kb * 64u + (uint32_t)quad * 4u) * 2u;

    f32x2 acc[RING][TW][2];
    uint32_t rowbase = 0, cur_stage = 0, prev_stage = 0;
    uint2 nraw[NCOL];
    auto fetch_row = [&](int q) {                          // q is a compile-time constant at every call site
      if (q % RC == 0) {                                   // first row of the next chunk of the unit
        prev_stage = cur_stage;
        cur_stage = (uint32_t)(g * NIG) + rstage;
        mbar_wait(in_full + 8u * cur_stage, rphase);
        rowbase = sIn + cur_stage * Cfg::CHUNK_BYTES + in_off;
        if (++rstage == NIG) { rstage = 0; rphase ^= 1u; }
      }
#pragma unroll
      for (int j = 0; j < NCOL; ++j) nraw[j] = lds64(rowbase + (uint32_t)(((q % RC) * BW + j) * 128));
    };
    fetch_row(0);
#pragma unroll
    for (int q = 0; q < HR; ++q) {
      f32x2 x[NCOL][2];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) { x[j][0] = f2_from_bf16x2(nraw[j].x); x[j][1] = f2_from_bf16x2(nraw[j].y); }
      if (q + 1 < HR) fetch_row(q + 1);
      if (q % RC == RC - 1 || q == HR - 1) {              // row q was the last of its chunk: hand the stage back
        __syncwarp();
        if (lane == 0) mbar_arrive(in_empty + 8u * (((q + 1) % RC == 0 && q + 1 < HR) ? prev_stage : cur_stage));
      }
      // input row q feeds tap row tr of output row o = (q - tr) / S; tap rows accumulate in order 0, 1, 2
      // (the order of depthwise_tma.cu), the shift seeds the accumulator
#pragma unroll
      for (int tr = 2; tr >= 0; --tr) {
        if ((q - tr) >= 0 && (q - tr) % S == 0 && (q - tr) / S < R) {
          const int o = (q - tr) / S, slot = o % RING;
#pragma unroll
          for (int c = 0; c < TW; ++c)
#pragma unroll
            for (int v = 0; v < 2; ++v) {   // FFMA2: two channels per instruction
              const f32x2 init = tr == 0 ? sh[v] : acc[slot][c][v];
              acc[slot][c][v] = f2_fma(x[c * S + 2][v], w[3 * tr + 2][v],
                                     f2_fma(x[c * S + 1][v], w[3 * tr + 1][v], f2_fma(x[c * S][v], w[3 * tr][v], init)));
            }
          if (tr == 2 && active) {                          // output row o is complete
#pragma unroll
            for (int c = 0; c < TW; ++c) {
              const uint32_t lo = pack2_f2<RELU>(acc[slot][c][0], p.cap2);
              const uint32_t hi = pack2_f2<RELU>(acc[slot][c][1], p.cap2);
              asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(obase + (size_t)(o * WO + c) * C * 2u), "r"(lo), "r"(hi) : "memory");
            }
          }
        }
      }
    }
  }
}

template <class Cfg>
cudaError_t launch_dr(bf16* out, const bf16* in, const float* taps, const float* shift, int act, int n, int pad_lo,
                      int num_sms, cudaStream_t st, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  CUtensorMap tin;
  cuuint64_t gdim[4] = {(cuuint64_t)Cfg::C, (cuuint64_t)Cfg::H, (cuuint64_t)Cfg::H, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)Cfg::C * 2, (cuuint64_t)Cfg::H * Cfg::C * 2, (cuuint64_t)Cfg::H * Cfg::H * Cfg::C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::RC, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { if (err) *err = "depthwise_ring: tensor map encode failed"; return cudaErrorInvalidValue; }
  DrParams p{};
  p.out = out; p.taps = taps; p.shift = shift;
  p.cap2 = act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.pad_lo = pad_lo;
  p.units = n * Cfg::BANDS * Cfg::STRIPS * Cfg::NKB;
  int grid = num_sms < p.units ? num_sms : p.units;
  {
    cudaError_t e = ensure_dyn_smem((const void*)depthwise_ring_kernel<Cfg, true>, (int)Cfg::SMEM);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)depthwise_ring_kernel<Cfg, false>, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
  }
  if (act != MNV1_ACT_NONE) return launch_pdl(depthwise_ring_kernel<Cfg, true>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, st, tin, p);
  return launch_pdl(depthwise_ring_kernel<Cfg, false>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, st, tin, p);
}

//                   S    C   H TWO R TW RC NG NIG
using DrL14 = DrCfg<1,  512, 14, 14, 7, 2, 9, 4, 2>;   // 14x14x512
// (round 2, tools/run_layer.py --layer 14: three groups 28.2 us, four 25.9-26.7, five 28.2 — the stencil is not short of
//  warps; experiments/README.md)
using DrL24 = DrCfg<2,  512, 14,  7, 7, 1, 15, 3, 2>;   // 14x14x512 -> 7x7

}  // namespace

// cudaErrorNotSupported (nothing launched) when the shape has no variant here.
cudaError_t launch_depthwise_ring(bf16* out, const bf16* in, const float* w9xC_scaled, const float* shift, int act, int n,
                                  int rows, int cols, int stride, int c, int pad_lo, int num_sms, cudaStream_t st,
                                  std::string* err) {
  if (rows != cols) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
#define DR_TRY(CFG) \
  if (stride == CFG::S && c == CFG::C && rows == CFG::H) \
    return launch_dr<CFG>(out, in, w9xC_scaled, shift, act, n, pad_lo, num_sms, st, err)
  DR_TRY(DrL14);
  DR_TRY(DrL24);
#undef DR_TRY
  return cudaErrorNotSupported;
}

}  // namespace mnv1
