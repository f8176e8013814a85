// fused_pair.cu — depthwise 3x3 + pointwise 1x1 as ONE kernel for the blocks whose pointwise filter does NOT
// fit in shared memory: the 512-channel 14x14 blocks, layers 14+15 ... 22+23 of the MobileNet.c schedule
// (`depthwise` kernel.cl:62-92 followed by `pointwise` kernel.cl:94-114).
//
// fused_rb.cu keeps the whole pointwise filter resident (<= 128 KB); here the filter is 512 KB, so it is
// STREAMED, and to halve what each SM has to pull the block runs on CTA PAIRS (tcgen05.mma.cta_group::2):
//   * a CTA owns one tile of R x TWO = 7 x 14 output pixels (half an image; TMEM lane = r*16 + x) and runs the
//     depthwise stencil of fused_rb.cu for it: NG groups of 4 warps, group g takes the 64-channel k-blocks
//     g, g+NG, ... of every tile and writes each finished row as bf16 straight into the 128B-swizzled K-major
//     A operand of that k-block (NA stages) — the depthwise map never leaves the SM;
//   * the two CTAs of a cluster work on two tiles (the two halves of one image) with ONE sequence of pair
//     UMMAs, M = 256: D[256 x 256] (+)= A_kb . B_(kb,nh)^T for the two 256-column halves nh of the 512 output
//     channels.  Each CTA stages only ITS 128 filter rows of a (kb, nh) tile (16 KB) through an NB-deep ring;
//     the pair's tensor cores read the other half from the peer.  Both accumulators (2 x 256 columns = the
//     whole TMEM) are live for the 8 k-blocks of a tile; the k-loop is outermost so an A stage is consumed
//     by both halves before it is released;
//   * the leader's MMA warp issues for the pair and multicasts its commits (a_empty, b_empty, tm_full) to
//     both CTAs; a_full lives in the leader and collects the arrivals of both CTAs' stencil groups
//     (mbarrier.arrive on the peer's shared::cluster address, acquire.cluster on the wait);
//   * 4 epilogue warps per CTA (one per TMEM lane quarter) drain half 0, release it to the MMAs of the next
//     tile, then drain half 1: fma(scale, shift) from the constant bank, ReLU6, private swizzled staging,
//     4-D TMA stores.
// Same arithmetic in the same order as depthwise_ring.cu followed by pointwise_pair.cu: bit-identical results.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

constexpr int FP_EPI_WARPS = 4;     // one per TMEM lane quarter
constexpr int FP_TP = 16;           // TMEM-lane pitch of a tile row
constexpr uint32_t FP_A_BYTES = 128 * 128;    // A stage: 128 rows x 64 bf16, 128B swizzle
constexpr uint32_t FP_BH_BYTES = 128 * 128;   // B stage: this CTA's 128 filter rows x 64 bf16

template <int S_, int NKB_, int COUT_, int H_, int TWO_, int R_, int TW_, int RC_, int NG_, int GW_, int NIG_, int NA_, int NB_, int NSTG_>
struct FpCfg {
  static constexpr int GW = GW_;   // warps per stencil group
  static constexpr int S = S_, NKB = NKB_, COUT = COUT_, H = H_, TWO = TWO_, R = R_, TW = TW_, RC = RC_;
  static constexpr int NG = NG_, NIG = NIG_, NI = NG_ * NIG_, NA = NA_, NB = NB_, NSTG = NSTG_;
  static constexpr int CK = 64, C = CK * NKB, NH = COUT / 256;
  static constexpr int HO = H / S, WO = H / S;
  static constexpr int BANDS = HO / R, STRIPS = WO / TWO;
  static constexpr int HR = (R - 1) * S + 3, BW = (TWO - 1) * S + 3, NCHK = (HR + RC - 1) / RC;
  static constexpr int PG = TWO / TW, NCOL = (TW - 1) * S + 3, RING = S == 1 ? 3 : 2;
  static constexpr uint32_t CHUNK_BYTES = (uint32_t)RC * BW * 128;
  // warp roles: stencil groups first (lowest issue priority), then the epilogue, then the single-thread roles
  static constexpr int W_EPI = NG * GW, W_MMA = W_EPI + FP_EPI_WARPS, W_TMA = W_MMA + 1, W_BPROD = W_MMA + 2;
  static constexpr int WARPS = W_MMA + 4;   // the single-thread roles share the last warpgroup (one idle warp)
  static_assert(W_EPI % 4 == 0, "roles are dispatched per warpgroup");
  static constexpr int THREADS = WARPS * 32;
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_B = OFF_A + NA * FP_A_BYTES;
  static constexpr uint32_t OFF_O = OFF_B + NB * FP_BH_BYTES;
  static constexpr uint32_t OFF_IN = OFF_O + FP_EPI_WARPS * NSTG * 4096u;
  static constexpr uint32_t OFF_TAPS = OFF_IN + NI * CHUNK_BYTES;
  static constexpr uint32_t OFF_DSH = OFF_TAPS + 9u * C * 4;
  static constexpr uint32_t OFF_BAR = OFF_DSH + (uint32_t)C * 4;
  static constexpr int NBAR = 2 * NI + 2 * NA + 2 * NB + 1 + NH;
  static constexpr uint32_t OFF_END = OFF_BAR + 8u * NBAR + 16;
  static constexpr size_t SMEM = 1024 + OFF_END;
  static_assert(COUT % 256 == 0 && NH >= 1 && NH <= 2, "Cout: 256 or 512 (two 256-column TMEM accumulators)");
  static_assert(R * FP_TP <= 128 && TWO <= FP_TP, "tile does not fit one UMMA M tile");
  static_assert(HO % R == 0 && WO % TWO == 0 && TWO % TW == 0 && PG * 16 <= GW * 32, "shape does not tile");
  static_assert((BANDS * STRIPS) % 2 == 0, "the two CTAs of a pair take the two halves of an image");
  static_assert(NA >= NG && NIG >= 2 && NB >= 2, "an A stage per group at least; rings hold at least two entries");
  static_assert(CHUNK_BYTES % 128 == 0, "chunk pitch");
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
};

struct FpParams {
  const float* dw_taps;    // [9][C] taps x folded-BN scale
  const float* dw_shift;   // [C] or nullptr
  float pw_scale[512];     // folded-BN scale / shift of the pointwise layer, by value: constant-bank operands
  float pw_shift[512];
  uint32_t dw_cap2, pw_cap2;
  int pad_lo;
  int images;              // a unit = one image = the pair's two tiles
};

__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FPW_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FPW_DONE;\n"
      "bra FPW_WAIT;\n"
      "FPW_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <class Cfg, bool DW_RELU, bool PW_RELU>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
fused_pair_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
                  const __grid_constant__ FpParams p) {
  constexpr int FP_GW = Cfg::GW;
  constexpr int S = Cfg::S, NKB = Cfg::NKB, NH = Cfg::NH, TWO = Cfg::TWO, R = Cfg::R, TW = Cfg::TW, C = Cfg::C;
  constexpr int RC = Cfg::RC, NG = Cfg::NG, NIG = Cfg::NIG, NI = Cfg::NI, NA = Cfg::NA, NB = Cfg::NB, NSTG = Cfg::NSTG;
  constexpr int HR = Cfg::HR, BW = Cfg::BW, NCHK = Cfg::NCHK, PG = Cfg::PG, NCOL = Cfg::NCOL, RING = Cfg::RING;
  constexpr int W_EPI = Cfg::W_EPI, W_MMA = Cfg::W_MMA, W_TMA = Cfg::W_TMA, W_BPROD = Cfg::W_BPROD;
  constexpr int TILES_PER_IMG = Cfg::BANDS * Cfg::STRIPS;   // 2: the pair's two tiles

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_g = smem_raw + (smem - smem_u32(smem_raw));
  const uint32_t sA = smem + Cfg::OFF_A, sB = smem + Cfg::OFF_B, sO = smem + Cfg::OFF_O, sIn = smem + Cfg::OFF_IN;
  const uint32_t sTaps = smem + Cfg::OFF_TAPS, sDsh = smem + Cfg::OFF_DSH;
  const uint32_t bars = smem + Cfg::OFF_BAR;
  const uint32_t in_full = bars, in_empty = in_full + 8u * NI, a_full = in_empty + 8u * NI, a_empty = a_full + 8u * NA;
  const uint32_t b_full = a_empty + 8u * NA, b_empty = b_full + 8u * NB, tm_full = b_empty + 8u * NB, tm_empty = tm_full + 8u;
  const uint32_t tmem_slot = tm_empty + 8u * NH;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  {  // constants -> shared memory
    constexpr int PER = (9 * C / 4 + Cfg::THREADS - 1) / Cfg::THREADS;
    stage_constants<PER>(reinterpret_cast<float*>(smem_g + Cfg::OFF_TAPS), p.dw_taps, 9 * C, tid, Cfg::THREADS);
    float* ds = reinterpret_cast<float*>(smem_g + Cfg::OFF_DSH);
    if (p.dw_shift) stage_constants<1>(ds, p.dw_shift, C, tid, Cfg::THREADS);
    else for (int i = tid; i < C; i += Cfg::THREADS) ds[i] = 0.f;
  }
  if (tid == 0) {
    prefetch_tmap(&tmap_in); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_out); prefetch_tmap(&tmap_out2);
    for (int s = 0; s < NI; ++s) { mbar_init(in_full + 8u * s, 1); mbar_init(in_empty + 8u * s, FP_GW); }
    for (int s = 0; s < NA; ++s) { mbar_init(a_full + 8u * s, 2 * FP_GW); mbar_init(a_empty + 8u * s, 1); }   // a_full: both CTAs' groups
    for (int s = 0; s < NB; ++s) { mbar_init(b_full + 8u * s, 1); mbar_init(b_empty + 8u * s, 1); }
    mbar_init(tm_full, 1);
    for (int h = 0; h < NH; ++h) mbar_init(tm_empty + 8u * h, 2 * FP_EPI_WARPS);                              // both CTAs' epilogue warps
    mbar_init_fence();
  }
  if (warp == W_MMA) tmem_alloc_pair(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer's barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  pdl_wait();                                       // the previous layer's output is complete and visible

  // this pair's images: cid, cid + num_clusters, ...; this CTA's tile of image i is tile `rank`
  const int nt = (p.images - cid + num_clusters - 1) / num_clusters;     // > 0: the grid never exceeds the image count
  const int total_units = nt * NKB;                                       // (tile, k-block) units of this CTA
  const int band = (int)rank / Cfg::STRIPS, strip = (int)rank % Cfg::STRIPS;

  // Register budget: 768 threads are launched with 80 registers each = 61440, and setmaxnreg only moves registers
  // INSIDE that allocation (an .inc that asks for more than the CTA's warps have released waits forever): the epilogue
  // warpgroup goes to 72, the single-thread roles to 56, the 16 stencil warps to 88 (512 x 88 + 128 x 72 + 128 x 56 =
  // 61440).  setmaxnreg is warpgroup-wide and must be the same
  // instruction for the 4 warps of a warpgroup, so the role dispatch branches per warpgroup first.
  if (warp >= W_MMA) {
   reg_dec<56>();
   if (warp == W_TMA) {
    // ======================= input producer: halo row-chunks of every (tile, k-block) unit =======================
    if (lane == 0) {
      int stage[NG]; uint32_t phase[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) { stage[g] = 0; phase[g] = 0; }
      const int cx = strip * TWO * S - p.pad_lo, cy = band * R * S - p.pad_lo;
      for (int u0 = 0; u0 < total_units; u0 += NG) {
        const int nu = total_units - u0 < NG ? total_units - u0 : NG;
        int cc[NG], ci[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int ul = u0 + (g < nu ? g : 0);
          const int lt = ul / NKB, kb = ul - lt * NKB;
          cc[g] = kb * 64; ci[g] = cid + lt * num_clusters;
        }
#pragma unroll 1
        for (int k = 0; k < NCHK; ++k) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (g < nu) {
              const uint32_t st = (uint32_t)(g * NIG + stage[g]);
              mbar_wait(in_empty + 8u * st, phase[g] ^ 1u);
              mbar_expect_tx(in_full + 8u * st, Cfg::CHUNK_BYTES);
              tma_load_4d(sIn + st * Cfg::CHUNK_BYTES, &tmap_in, in_full + 8u * st, cc[g], cx, cy + k * RC, ci[g]);
              if (++stage[g] == NIG) { stage[g] = 0; phase[g] ^= 1u; }
            }
          }
        }
      }
    }
   } else if (warp == W_BPROD) {
    // ======================= filter producer: this CTA's 128 rows of every (k-block, half) tile =======================
    // Completion is signalled on the LEADER's b_full; the leader's thread announces the bytes of both CTAs (a
    // complete_tx that lands first only makes the count negative for a while).
    if (lane == 0) {
      const uint32_t b_full_leader = mapa_shared(b_full, 0);
      int stage = 0; uint32_t phase = 0;
      for (int lt = 0; lt < nt; ++lt)
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
          for (int nh = 0; nh < NH; ++nh) {
            mbar_wait(b_empty + 8u * stage, phase ^ 1u);
            if (rank == 0) mbar_expect_tx(b_full + 8u * stage, 2 * FP_BH_BYTES);
            tma_load_2d_pair(sB + (uint32_t)stage * FP_BH_BYTES, &tmap_b, b_full_leader + 8u * stage, kb * 64, nh * 256 + (int)rank * 128);
            if (++stage == NB) { stage = 0; phase ^= 1u; }
          }
    }
   } else if (warp == W_MMA) {
    // ======================= MMA issuer: the leader's converged warp, one elected lane =======================
    constexpr uint32_t idesc = umma_idesc_bf16_m256(256);
    const uint32_t elected = rank == 0 ? elect_one() : 0u;
    int bs = 0; uint32_t bph = 0;
    int ul = 0;
    for (int lt = 0; rank == 0 && lt < nt; ++lt) {
#pragma unroll 1
      for (int kb = 0; kb < NKB; ++kb, ++ul) {
        const uint32_t st = (uint32_t)ul % NA, ph = ((uint32_t)ul / NA) & 1u;
        mbar_wait_cluster(a_full + 8u * st, ph);            // both CTAs' stencil groups have written this A stage
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(sA + st * FP_A_BYTES);
#pragma unroll
        for (int nh = 0; nh < NH; ++nh) {
          if (kb == 0) { mbar_wait_cluster(tm_empty + 8u * nh, ((uint32_t)lt & 1u) ^ 1u); tc_fence_after(); }   // the epilogues drained this half
          mbar_wait(b_full + 8u * bs, bph);
          tc_fence_after();
          const uint64_t db = umma_desc_sw128(sB + (uint32_t)bs * FP_BH_BYTES);
          const uint32_t tmem_d = tmem_base + (uint32_t)nh * 256u;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_pair_if(elected, tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_pair_if(elected, b_empty + 8u * bs);   // frees the filter slot in both CTAs
          if (++bs == NB) { bs = 0; bph ^= 1u; }
        }
        umma_commit_pair_if(elected, a_empty + 8u * st);     // frees the A stage in both CTAs
      }
      umma_commit_pair_if(elected, tm_full);                 // both CTAs' accumulators are complete
    }
   }
  } else if (warp >= W_EPI) {
    // ======================= epilogue: one warp per TMEM lane quarter =======================
    reg_dec<72>();
    const int quarter = warp & 3;                          // TMEM lanes 32*quarter .. +31 = tile rows 2*quarter, +1
    const int rsel = lane >> 4, mx = lane & 15;
    const bool valid = mx < TWO && 2 * quarter + rsel < R;
    const uint32_t wbuf = sO + (uint32_t)(warp - W_EPI) * (NSTG * 4096u);
    const uint32_t line = (uint32_t)(rsel * TWO + mx);      // dense [2][TWO] lines of 128 B, as the 2-row TMA box reads them
    const uint32_t line_off = line * 128u, line_x = line & 7u;
    const uint32_t tm_empty_leader = mapa_shared(tm_empty, 0);
    const int row0 = 2 * quarter;
    uint32_t blk = 0;
    for (int lt = 0; lt < nt; ++lt) {
      const int img = cid + lt * num_clusters;
      mbar_wait(tm_full, (uint32_t)lt & 1u);
      tc_fence_after();
#pragma unroll
      for (int nh = 0; nh < NH; ++nh) {
#pragma unroll
        for (int b = 0; b < 4; ++b, ++blk) {
          const uint32_t sbuf = wbuf + (NSTG == 1 ? 0u : (blk % NSTG) * 4096u);
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(nh * 256 + b * 64);
          uint32_t v[32];
          tmem_ld32_nowait(taddr, v);
          if (lane == 0) tma_store_wait_read<NSTG - 1>();      // the store that last read this buffer is done with it
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld_wait();
            uint32_t q[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int col = nh * 256 + b * 64 + half * 32 + j;
              const f32x2 acc2 = f2_pack(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              q[j / 2] = pack2_f2<PW_RELU>(f2_fma(acc2, f2_pack(p.pw_scale[col], p.pw_scale[col + 1]),
                                                  f2_pack(p.pw_shift[col], p.pw_shift[col + 1])), p.pw_cap2);
            }
            if (half == 0) tmem_ld32_nowait(taddr + 32u, v);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
              if (valid) sts128(sbuf + line_off + (((uint32_t)(half * 4 + c4) ^ line_x) << 4), q[4 * c4], q[4 * c4 + 1], q[4 * c4 + 2], q[4 * c4 + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int y = band * R + row0;
            if (row0 + 1 < R) tma_store_4d(&tmap_out2, sbuf, nh * 256 + b * 64, strip * TWO, y, img);
            else if (row0 < R) tma_store_4d(&tmap_out, sbuf, nh * 256 + b * 64, strip * TWO, y, img);
            tma_store_commit();
          }
        }
        // this warp's lanes of half nh are in registers / staging: the next tile's MMAs may overwrite them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tm_empty_leader + 8u * nh);
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else {
    // ======================= stencil groups =======================
    reg_inc<88>();
    const int g = warp / FP_GW;
    const int t = tid - g * FP_GW * 32;
    const bool active = t < PG * 16;
    const int quad = t & 15;
    const int pg = active ? t >> 4 : PG - 1;
    const uint32_t in_off = (uint32_t)(pg * TW * S) * 128u + (uint32_t)quad * 8u;
    uint32_t a_off[TW];
#pragma unroll
    for (int c = 0; c < TW; ++c) {
      const uint32_t x = (uint32_t)(pg * TW + c);
      a_off[c] = x * 128u + ((((uint32_t)quad >> 1) ^ (x & 7u)) << 4) + ((uint32_t)quad & 1u) * 8u;
    }
    const uint32_t a_full_leader = mapa_shared(a_full, 0);
    uint32_t rstage = 0, rphase = 0;

    for (int u0 = 0; u0 < total_units; u0 += NG) {
      const int nu = total_units - u0 < NG ? total_units - u0 : NG;
      if (g >= nu) break;
      const int ul = u0 + g;
      const int kb = ul % NKB;
      f32x2 w[9][2], sh[2];                                 // the thread's 4 channels as two packed fp32 pairs
      {
        const uint32_t ch = (uint32_t)(kb * 64 + quad * 4) * 4u;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float4 a = lds128f(sTaps + (uint32_t)k * C * 4u + ch);
          w[k][0] = f2_pack(a.x, a.y); w[k][1] = f2_pack(a.z, a.w);
        }
        const float4 a = lds128f(sDsh + ch);
        sh[0] = f2_pack(a.x, a.y); sh[1] = f2_pack(a.z, a.w);
      }
      const uint32_t ast = (uint32_t)ul % NA, aph = ((uint32_t)ul / NA) & 1u;
      const uint32_t dstA = sA + ast * FP_A_BYTES;

      f32x2 acc[RING][TW][2];
      uint32_t rowbase = 0, cur_stage = 0, prev_stage = 0;
      uint2 nraw[NCOL];
      auto fetch_row = [&](int q) {                          // q is a compile-time constant at every call site
        if (q % RC == 0) {
          prev_stage = cur_stage;
          cur_stage = (uint32_t)(g * NIG) + rstage;
          mbar_wait_relaxed(in_full + 8u * cur_stage, rphase);
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
#pragma unroll
        for (int tr = 2; tr >= 0; --tr) {
          if ((q - tr) >= 0 && (q - tr) % S == 0 && (q - tr) / S < R) {
            const int o = (q - tr) / S, slot = o % RING;
#pragma unroll
            for (int c = 0; c < TW; ++c)
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const f32x2 init = tr == 0 ? sh[v] : acc[slot][c][v];
                acc[slot][c][v] = f2_fma(x[c * S + 2][v], w[3 * tr + 2][v],
                                       f2_fma(x[c * S + 1][v], w[3 * tr + 1][v], f2_fma(x[c * S][v], w[3 * tr][v], init)));
              }
            if (tr == 2) {                                  // output row o is complete
              if (o == 0) mbar_wait_relaxed(a_empty + 8u * ast, aph ^ 1u);   // the pair's MMAs that last read this A stage retired
              if (active) {
#pragma unroll
                for (int c = 0; c < TW; ++c)
                  sts64(dstA + a_off[c] + (uint32_t)(o * FP_TP) * 128u, pack2_f2<DW_RELU>(acc[slot][c][0], p.dw_cap2),
                        pack2_f2<DW_RELU>(acc[slot][c][1], p.dw_cap2));
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(a_full + 8u * ast);
        else mbar_arrive_cluster(a_full_leader + 8u * ast);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer may still arrive on this CTA's barriers / read its shared memory
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512u);
  }
}

cudaError_t fp_encode_b(CUtensorMap* map, const void* base, uint64_t rows, uint64_t k, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  cuuint64_t gdim[2] = {k, rows};
  cuuint64_t gstride[1] = {k * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { if (err) *err = "fused pair: filter tensor map encode failed"; return cudaErrorInvalidValue; }
  return cudaSuccess;
}

template <class Cfg>
cudaError_t launch_fp(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n, int pad_lo, int num_sms,
                      cudaStream_t st, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  constexpr int C = Cfg::C, COUT = Cfg::COUT, H = Cfg::H, Ho = Cfg::HO, Wo = Cfg::WO;
  CUtensorMap tin, tb, tout, tout2;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)H * C * 2, (cuuint64_t)H * H * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::RC, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "fused pair: input tensor map encode failed"; return cudaErrorInvalidValue; }
  }
  cudaError_t e = fp_encode_b(&tb, pw->w_bf16, (uint64_t)COUT, (uint64_t)C, err);
  if (e != cudaSuccess) return e;
  for (int rows2 = 1; rows2 <= 2; ++rows2) {   // output boxes of one and of two tile rows
    cuuint64_t gdim[4] = {(cuuint64_t)COUT, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)n};
    cuuint64_t gstr[3] = {(cuuint64_t)COUT * 2, (cuuint64_t)Wo * COUT * 2, (cuuint64_t)Ho * Wo * COUT * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Cfg::TWO, (cuuint32_t)rows2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(rows2 == 1 ? &tout : &tout2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "fused pair: output tensor map encode failed"; return cudaErrorInvalidValue; }
  }
  FpParams p{};
  p.dw_taps = dw->w_scaled; p.dw_shift = dw->shift;
  for (int i = 0; i < 512; ++i) {
    p.pw_scale[i] = i < (int)pw->h_scale.size() ? pw->h_scale[i] : 1.f;
    p.pw_shift[i] = i < (int)pw->h_shift.size() ? pw->h_shift[i] : 0.f;
  }
  p.dw_cap2 = dw->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.pw_cap2 = pw->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.pad_lo = pad_lo;
  p.images = n;
  const bool dr = dw->act != MNV1_ACT_NONE, pr = pw->act != MNV1_ACT_NONE;
  long clusters = num_sms / 2;
  if (clusters > n) clusters = n;
  {
    cudaError_t ea = cudaSuccess;
    auto set = [&](const void* f) { if (ea == cudaSuccess) ea = ensure_dyn_smem(f, (int)Cfg::SMEM); };
    set((const void*)fused_pair_kernel<Cfg, true, true>); set((const void*)fused_pair_kernel<Cfg, true, false>);
    set((const void*)fused_pair_kernel<Cfg, false, true>); set((const void*)fused_pair_kernel<Cfg, false, false>);
    if (ea != cudaSuccess) return ea;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (dr) return pr ? cudaLaunchKernelEx(&cfg, fused_pair_kernel<Cfg, true, true>, tin, tb, tout, tout2, p)
                    : cudaLaunchKernelEx(&cfg, fused_pair_kernel<Cfg, true, false>, tin, tb, tout, tout2, p);
  return pr ? cudaLaunchKernelEx(&cfg, fused_pair_kernel<Cfg, false, true>, tin, tb, tout, tout2, p)
            : cudaLaunchKernelEx(&cfg, fused_pair_kernel<Cfg, false, false>, tin, tb, tout, tout2, p);
}

//                   S NKB COUT  H TWO R TW RC NG GW NIG NA NB NSTG
#ifndef MNV1_FP_VARIANT
#define MNV1_FP_VARIANT 0
#endif
#if MNV1_FP_VARIANT == 1     // two groups of 8 warps, one column per thread, two A stages per group
using FpL14 = FpCfg<1, 8, 512, 14, 14, 7, 1, 3, 2, 8, 2, 4, 4, 2>;
#elif MNV1_FP_VARIANT == 2   // the same with three A stages per group
using FpL14 = FpCfg<1, 8, 512, 14, 14, 7, 1, 3, 2, 8, 3, 6, 3, 1>;
#else                        // four groups of 4 warps, one A stage per group
using FpL14 = FpCfg<1, 8, 512, 14, 14, 7, 2, 3, 4, 4, 2, 4, 4, 1>;
#endif   // 14x14x512 -> 14x14x512 (layers 14+15 ... 22+23)

bool fp_match(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride) {
  if (!dw->w_scaled || !pw->w_bf16 || pw->cin != dw->cout) return false;
  return stride == FpL14::S && dw->cout == FpL14::C && pw->cout == FpL14::COUT && rows == FpL14::H && cols == FpL14::H;
}

}  // namespace

bool fused_pair_supported(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride) {
  return fp_match(dw, pw, rows, cols, stride);
}

// cudaErrorNotSupported (nothing launched) when the block has no CTA-pair variant.
cudaError_t launch_fused_pair(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n, int rows, int cols,
                              int stride, int pad_lo, int num_sms, cudaStream_t st, std::string* err) {
  if (!fp_match(dw, pw, rows, cols, stride) || num_sms < 2) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
  return launch_fp<FpL14>(out, in, dw, pw, n, pad_lo, num_sms, st, err);
}

}  // namespace mnv1
