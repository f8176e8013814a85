// fused_block.cu — depthwise 3x3 (stride 1) -> pointwise 1x1 as ONE kernel for bf16 contexts.
//
// Replaces a `depthwise` launch followed by a `pointwise` launch (kernel.cl:62-92 then :94-114;
// e.g. MobileNet.c:1406-1471 + :1496-1560 for layers 14/15) when the feature map is W columns
// wide with W*R <= 128 pixels per work item.  The depthwise output never goes to HBM: the
// stencil warps write it, already bf16 and 128B-swizzled, as the K-major A operand of the
// tcgen05 GEMM.  Per pair this removes one feature-map write and one read (SURVEY App. B:
// 2 x 100 352 elements per image for the 14x14x512 pairs) and hides the issue-bound stencil
// behind the MMAs.
//
// Work item = (image, band of R output rows): M tile = R*W pixels (98 of 128 rows used at 14x14),
// all Cin channels in 64-wide k-blocks, all Cout (<= 512) channels as one or two 256-column
// accumulators in TMEM.  20 warps:
//   warp 0        TMA producer: depthwise input halo tiles [R+2][W+2][64] (4-D map, OOB = padding)
//                 and filter tiles [256][64] into two mbarrier rings
//   warp 1        TMEM allocator + tcgen05.mma issuer
//   warps 2-15    stencil (the pace-setter, so it gets most of the SM): warp = one column, lane =
//                 2 channels; input-row-major accumulation into a 3-row ring (as depthwise_tma.cu),
//                 BN shift/ReLU6, pack, st.shared into A[k-block]
//   warps 16-19   epilogue: tcgen05.ld -> fma(scale, shift) -> ReLU/cap -> swizzled staging -> TMA store
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace mnv1 {
namespace {

constexpr int FB_DW_WARPS = 14, FB_EPI_WARPS = 4;
constexpr int FB_THREADS = (2 + FB_DW_WARPS + FB_EPI_WARPS) * 32;
constexpr uint32_t FB_A_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr uint32_t FB_B_BYTES = 256 * 128;        // 256 filters x 64 bf16
constexpr uint32_t FB_O_BYTES = 128 * 128;
constexpr int FB_B_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FB_DONE;\n"
      "bra FB_WAIT;\n"
      "FB_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B (pointwise_tc.cu)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
constexpr uint32_t FB_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(FB_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi, uint32_t cap2) {
  uint32_t d;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  asm("min.bf16x2 %0, %0, %1;" : "+r"(d) : "r"(cap2));
  return d;
}

struct FbParams {
  const float* dw_taps;    // [9][C] taps x folded-BN scale
  const float* dw_shift;   // [C] or nullptr
  const float* pw_scale;   // [Cout] or nullptr
  const float* pw_shift;   // [Cout] or nullptr
  uint32_t dw_cap2, pw_cap2;
  int n, C, Cout, bands;   // bands = H / R
  long items;              // n * bands
};

// W: map width (= columns per item); R: output rows per item (R*W <= 128).
template <int W, int R, bool DW_RELU, bool PW_RELU>
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_dw_pw_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_out, const FbParams p) {
  constexpr int HR = R + 2, HW = W + 2;                       // halo tile
  constexpr uint32_t IN_BYTES = (uint32_t)HR * HW * 64 * 2;
  constexpr uint32_t IN_PITCH = (IN_BYTES + 1023u) & ~1023u;
  static_assert(R * W <= 128 && W <= FB_DW_WARPS, "item does not fit one UMMA M tile / the stencil warps");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sIn = smem;                                   // 2 halo tiles
  const uint32_t sA = sIn + 2 * IN_PITCH;                      // 2 A tiles
  const uint32_t sB = sA + 2 * FB_A_BYTES;                     // 3 filter tiles
  const uint32_t sO = sB + FB_B_STAGES * FB_B_BYTES;           // 2 staging blocks
  const uint32_t sTaps = sO + 2 * FB_O_BYTES;                  // [9][C] fp32
  const uint32_t sDwShift = sTaps + 9u * p.C * 4u;             // [C]
  const uint32_t sPwScale = sDwShift + (uint32_t)p.C * 4u;     // [Cout]
  const uint32_t sPwShift = sPwScale + (uint32_t)p.Cout * 4u;  // [Cout]
  const uint32_t bars = (sPwShift + (uint32_t)p.Cout * 4u + 15u) & ~15u;
  const uint32_t in_full = bars, in_empty = bars + 16, a_full = bars + 32, a_empty = bars + 48;
  const uint32_t b_full = bars + 64, b_empty = bars + 96, tm_full = bars + 128, tm_empty = bars + 136;
  const uint32_t tmem_slot = bars + 144;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C, Cout = p.Cout, num_kb = C / 64, n_tiles = Cout / 256;

  // constants -> shared memory (generic stores through the shared window)
  {
    float* taps = reinterpret_cast<float*>(smem_raw + (sTaps - smem_u32(smem_raw)));
    for (int i = tid; i < 9 * C; i += FB_THREADS) taps[i] = p.dw_taps[i];
    float* ds = reinterpret_cast<float*>(smem_raw + (sDwShift - smem_u32(smem_raw)));
    for (int i = tid; i < C; i += FB_THREADS) ds[i] = p.dw_shift ? p.dw_shift[i] : 0.f;
    float* ps = reinterpret_cast<float*>(smem_raw + (sPwScale - smem_u32(smem_raw)));
    float* pt = reinterpret_cast<float*>(smem_raw + (sPwShift - smem_u32(smem_raw)));
    for (int i = tid; i < Cout; i += FB_THREADS) { ps[i] = p.pw_scale ? p.pw_scale[i] : 1.f; pt[i] = p.pw_shift ? p.pw_shift[i] : 0.f; }
  }
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(in_full + 8 * s, 1); mbar_init(in_empty + 8 * s, FB_DW_WARPS);
      mbar_init(a_full + 8 * s, FB_DW_WARPS); mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < FB_B_STAGES; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
    mbar_init(tm_full, 1); mbar_init(tm_empty, FB_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int is = 0; uint32_t iph = 0; int bs = 0; uint32_t bph = 0;
      for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
        const int img = (int)(it / p.bands), band = (int)(it % p.bands);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(in_empty + 8 * is, iph ^ 1u);
          mbar_expect_tx(in_full + 8 * is, IN_BYTES);
          tma_load_4d(sIn + is * IN_PITCH, &tmap_in, in_full + 8 * is, kb * 64, -1, band * R - 1, img);
          if (++is == 2) { is = 0; iph ^= 1u; }
          for (int nt = 0; nt < n_tiles; ++nt) {
            mbar_wait(b_empty + 8 * bs, bph ^ 1u);
            mbar_expect_tx(b_full + 8 * bs, FB_B_BYTES);
            tma_load_2d(sB + bs * FB_B_BYTES, &tmap_b, b_full + 8 * bs, kb * 64, nt * 256);
            if (++bs == FB_B_STAGES) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; uint32_t tph = 0;
      for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
        mbar_wait(tm_empty, tph ^ 1u);                    // epilogue drained the accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(a_full + 8 * as, aph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(sA + as * FB_A_BYTES);
          for (int nt = 0; nt < n_tiles; ++nt) {
            mbar_wait(b_full + 8 * bs, bph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t db = make_smem_desc(sB + bs * FB_B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (uint32_t)(nt * 256), da + 2 * k, db + 2 * k, (kb | k) ? 1u : 0u);
            umma_commit(b_empty + 8 * bs);
            if (++bs == FB_B_STAGES) { bs = 0; bph ^= 1u; }
          }
          umma_commit(a_empty + 8 * as);
          if (++as == 2) { as = 0; aph ^= 1u; }
        }
        umma_commit(tm_full);
        tph ^= 1u;
      }
    }
  } else if (warp < 2 + FB_DW_WARPS) {
    // ======================= stencil warps =======================
    const int col = warp - 2;                             // one column per warp
    const int cp = lane;                                  // channel pair inside the 64-wide k-block
    const bool active = col < W;
    const int colc = active ? col : 0;
    int is = 0; uint32_t iph = 0; int as = 0; uint32_t aph = 0;
    for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        // this k-block's taps and shift for the thread's 2 channels
        float w[9][2], sh[2];
        const uint32_t tap0 = sTaps + (uint32_t)(kb * 64 + cp * 2) * 4u;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float2 a = lds64f(tap0 + (uint32_t)k * C * 4u);
          w[k][0] = a.x; w[k][1] = a.y;
        }
        {
          const float2 a = lds64f(sDwShift + (uint32_t)(kb * 64 + cp * 2) * 4u);
          sh[0] = a.x; sh[1] = a.y;
        }
        mbar_wait(in_full + 8 * is, iph);                 // halo tile landed
        mbar_wait(a_empty + 8 * as, aph ^ 1u);            // A tile no longer read by the MMAs
        const uint32_t src = sIn + is * IN_PITCH + (uint32_t)(colc * 64 + cp * 2) * 2u;
        const uint32_t dstA = sA + as * FB_A_BYTES + ((uint32_t)(cp & 3) << 2);
        float acc[3][2];
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a][0] = acc[a][1] = 0.f;
#pragma unroll
        for (int q = 0; q < HR; ++q) {                    // input row q of the halo tile
          float x[3][2];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const uint32_t raw = lds32(src + (uint32_t)((q * HW + j) * 64 * 2));
            x[j][0] = bf16lo_to_f32(raw); x[j][1] = bf16hi_to_f32(raw);
          }
          const int a0 = q % 3, a1 = (q + 2) % 3, a2 = (q + 1) % 3;   // output rows q, q-1, q-2
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            acc[a0][v] = fmaf(x[2][v], w[2][v], fmaf(x[1][v], w[1][v], fmaf(x[0][v], w[0][v], sh[v])));
            acc[a1][v] = fmaf(x[2][v], w[5][v], fmaf(x[1][v], w[4][v], fmaf(x[0][v], w[3][v], acc[a1][v])));
            acc[a2][v] = fmaf(x[2][v], w[8][v], fmaf(x[1][v], w[7][v], fmaf(x[0][v], w[6][v], acc[a2][v])));
          }
          if (q >= 2 && active) {                         // output row q-2 is complete
            const int pix = (q - 2) * W + col;            // row of the A tile
            sts32(dstA + (uint32_t)pix * 128u + ((uint32_t)((cp >> 2) ^ (pix & 7)) << 4),
                  pack2<DW_RELU>(acc[a2][0], acc[a2][1], p.dw_cap2));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive(a_full + 8 * as); mbar_arrive(in_empty + 8 * is); }
        if (++is == 2) { is = 0; iph ^= 1u; }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ======================= epilogue warps =======================
    const int ew = warp - (2 + FB_DW_WARPS);
    const int quarter = warp & 3;                          // TMEM lane quarter = warp % 4
    const int row = quarter * 32 + lane;
    const bool leader = ew == 0 && lane == 0;
    const uint32_t row_off = (uint32_t)row * 128u, row_x = (uint32_t)(row & 7);
    uint32_t tph = 0, blk = 0;
    for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
      const int img = (int)(it / p.bands), band = (int)(it % p.bands);
      const int m0 = (img * p.bands + band) * (R * W);     // first output pixel of the item
      mbar_wait(tm_full, tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nblk = Cout / 64;
#pragma unroll 1
      for (int b = 0; b < nblk; ++b, ++blk) {
        const uint32_t sbuf = sO + (blk & 1u) * FB_O_BYTES;
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * FB_EPI_WARPS) : "memory");
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          {
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * 64 + 32 * half);
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
          const uint32_t colb = (uint32_t)(b * 64 + 32 * half);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 s0 = lds128f(sPwScale + (colb + j) * 4u), s1 = lds128f(sPwScale + (colb + j + 4) * 4u);
            const float4 t0 = lds128f(sPwShift + (colb + j) * 4u), t1 = lds128f(sPwShift + (colb + j + 4) * 4u);
            const uint32_t q0 = pack2<PW_RELU>(fmaf(__uint_as_float(v[j + 0]), s0.x, t0.x), fmaf(__uint_as_float(v[j + 1]), s0.y, t0.y), p.pw_cap2);
            const uint32_t q1 = pack2<PW_RELU>(fmaf(__uint_as_float(v[j + 2]), s0.z, t0.z), fmaf(__uint_as_float(v[j + 3]), s0.w, t0.w), p.pw_cap2);
            const uint32_t q2 = pack2<PW_RELU>(fmaf(__uint_as_float(v[j + 4]), s1.x, t1.x), fmaf(__uint_as_float(v[j + 5]), s1.y, t1.y), p.pw_cap2);
            const uint32_t q3 = pack2<PW_RELU>(fmaf(__uint_as_float(v[j + 6]), s1.z, t1.z), fmaf(__uint_as_float(v[j + 7]), s1.w, t1.w), p.pw_cap2);
            const uint32_t chunk = (uint32_t)(4 * half + j / 8);
            sts128(sbuf + row_off + ((chunk ^ row_x) << 4), q0, q1, q2, q3);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * FB_EPI_WARPS) : "memory");
        if (leader) {
          // box = 64 columns x (R*W) rows: the padding rows of the M tile are never written
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_out),
                       "r"(sbuf), "r"(b * 64), "r"(m0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tm_empty);
      tph ^= 1u;
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

template <int W, int R>
cudaError_t launch_fb(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n, int H,
                      int num_sms, cudaStream_t st, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  const int C = dw->cout, Cout = pw->cout;
  CUtensorMap tin, tout;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(W + 2), (cuuint32_t)(R + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "fused block: input tensor map encode failed"; return cudaErrorInvalidValue; }
  }
  {
    const uint64_t m = (uint64_t)n * H * W;
    cuuint64_t gdim[2] = {(cuuint64_t)Cout, m};
    cuuint64_t gstr[1] = {(cuuint64_t)Cout * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(R * W)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "fused block: output tensor map encode failed"; return cudaErrorInvalidValue; }
  }
  FbParams p{};
  p.dw_taps = dw->w_scaled; p.dw_shift = dw->shift; p.pw_scale = pw->scale; p.pw_shift = pw->shift;
  p.dw_cap2 = dw->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.pw_cap2 = pw->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.n = n; p.C = C; p.Cout = Cout; p.bands = H / R; p.items = (long)n * p.bands;
  constexpr uint32_t IN_PITCH = (((uint32_t)(R + 2) * (W + 2) * 64 * 2) + 1023u) & ~1023u;
  const size_t smem = 1024 + 2 * IN_PITCH + 2 * FB_A_BYTES + FB_B_STAGES * FB_B_BYTES + 2 * FB_O_BYTES +
                      (size_t)9 * C * 4 + (size_t)C * 4 + (size_t)Cout * 8 + 16 + 160;
  if (smem > 227 * 1024) { if (err) *err = "fused block: shared memory budget exceeded"; return cudaErrorNotSupported; }
  const bool dr = dw->act != MNV1_ACT_NONE, pr = pw->act != MNV1_ACT_NONE;
  long grid = num_sms;
  if (grid > p.items) grid = p.items;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); };
    set((const void*)fused_dw_pw_kernel<W, R, true, true>); set((const void*)fused_dw_pw_kernel<W, R, true, false>);
    set((const void*)fused_dw_pw_kernel<W, R, false, true>); set((const void*)fused_dw_pw_kernel<W, R, false, false>);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
#define FB_LAUNCH(A, B) fused_dw_pw_kernel<W, R, A, B><<<(unsigned)grid, FB_THREADS, smem, st>>>(tin, pw->tmap_b, tout, p)
  if (dr) { if (pr) FB_LAUNCH(true, true); else FB_LAUNCH(true, false); }
  else    { if (pr) FB_LAUNCH(false, true); else FB_LAUNCH(false, false); }
#undef FB_LAUNCH
  return cudaGetLastError();
}

}  // namespace

namespace {
// The 14x14 streaming-filter variant is kept for experiments (MNV1_FB=1): it re-reads the whole
// pointwise filter from L2 for every 98-pixel item and is slower than the two separate kernels.
bool fb_supported(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride) {
  static const bool fb_on = getenv("MNV1_FB") != nullptr;
  if (!fb_on) return false;
  if (stride != 1 || !dw->w_scaled || !pw->has_tmap || pw->tmap_bn != 256 || pw->cin != dw->cout) return false;
  if (dw->cout % 64 || pw->cout % 256 || pw->cout > 512 || dw->cout > 1024) return false;
  return rows == 14 && cols == 14;
}
}  // namespace

bool fused_dw_pw_supported(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride) {
  return fused_rb_supported(dw, pw, rows, cols, stride) || fb_supported(dw, pw, rows, cols, stride);
}

// cudaErrorNotSupported (nothing launched) when the block shape has no fused variant.
cudaError_t launch_fused_dw_pw(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n,
                               int rows, int cols, int stride, int pad_lo, int num_sms, cudaStream_t st,
                               std::string* err) {
  {  // blocks whose pointwise filter stays resident in shared memory (layers 2-11)
    cudaError_t e = launch_fused_rb(out, in, dw, pw, n, rows, cols, stride, pad_lo, num_sms, st, err);
    if (e != cudaErrorNotSupported) return e;
  }
  if (!fb_supported(dw, pw, rows, cols, stride)) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
  return launch_fb<14, 7>(out, in, dw, pw, n, rows, num_sms, st, err);
}

}  // namespace mnv1
