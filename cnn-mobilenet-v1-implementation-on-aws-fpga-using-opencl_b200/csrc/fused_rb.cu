// fused_rb.cu — depthwise 3x3 (stride 1 or 2) + pointwise 1x1 as ONE kernel, for the blocks whose
// pointwise filter fits in shared memory (layers 2-11 of the MobileNet.c schedule).
//
// Replaces a `depthwise` launch (kernel.cl:62-92) and the `pointwise` launch that follows it
// (kernel.cl:94-114) in mnv1_forward*: the depthwise map — 45 % of the activation traffic of those
// layers — is never written to HBM.  Arithmetic is the same as depthwise_tma.cu
// followed by pointwise_tc.cu (fp32 stencil in the same tap order, bf16 rounding of the depthwise
// value, bf16 x bf16 -> fp32 UMMA), so the results are bit-identical to the two-kernel path.
//
// A tile is R x TWO output pixels (<= 128, one UMMA M tile; TMEM lane = r*TP + x, TP = 16).  A unit
// is (tile, 64-channel k-block).  One persistent CTA per SM:
//   top warp      TMA producer: the whole pointwise filter once (it stays resident), then the
//                 depthwise input as halo row-chunks [RC][(TWO-1)*S+3][CK] of each unit through a 4-D
//                 tensor map (out-of-bounds = the layer's zero padding) into a ring of NI stages
//   next          TMEM allocator + single-thread tcgen05.mma issuer: D[128 x Cout] (+)= A_unit . B_kb^T;
//                 512 / Cout accumulator stages so the epilogue's latency never stalls the MMAs
//   next 4        epilogue, each warp on its own (no barrier between them): tcgen05.ld of its 32 lanes
//                 -> fma(scale, shift) -> ReLU/cap -> private swizzled staging -> one 4-D TMA store per tile row
//   warps 0..     NG stencil groups of 4 warps; group g owns units g, g+NG, ...: a thread keeps 4
//                 channels x TW columns, walks the halo rows once (ld.shared.v2, widen once), keeps a
//                 ring of 3 (S=1) / 2 (S=2) output-row accumulators, and writes each finished row as
//                 bf16 straight into the 128B-swizzled K-major A operand of its unit.
// Every group has a private ring of NIG chunk stages (consecutive uses of a stage are waited on by the
// same warps, which is what makes the parity wait safe: TMA completions are not ordered, so a group
// sharing stages with the others could run two phases ahead of a barrier).  The producer fills the
// rings in "wavefront" order (chunk k of the NG units of a round, then chunk k+1 ...).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

constexpr int RB_EPI_WARPS = 4;   // one per TMEM lane quarter
constexpr int RB_TP = 16;         // TMEM-lane pitch of a tile row
constexpr uint32_t RB_A_BYTES = 128 * 128;   // A tile: 128 rows x 64 bf16, 128B swizzle
constexpr uint32_t RB_O_BYTES = 4 * 4096;    // output staging per buffer: 4 warps x 2 tile rows x (16 lines x 128 B)

// S stride; CK channels per k-block; NKB k-blocks (C = CK*NKB); COUT; tile = R x TWO outputs;
// TW output columns per stencil thread; RC input rows per chunk; NG stencil groups of GW warps; NIG
// chunk stages per group; NA >= NG A-tile stages (unit u uses stage u % NA; the MMAs retire in order, so a
// group can never be two uses of a stage ahead of its a_empty barrier); NE
// epilogue sets of 4 warps (set e takes tiles e, e+NE, ...); NSTG output staging buffers per warp.
template <int S_, int CK_, int NKB_, int COUT_, int TWO_, int R_, int TW_, int RC_, int NG_, int GW_, int NIG_, int NA_, int NE_, int NSTG_>
struct RbCfg {
  static constexpr int S = S_, CK = CK_, NKB = NKB_, COUT = COUT_, TWO = TWO_, R = R_, TW = TW_, RC = RC_;
  static constexpr int NG = NG_, GW = GW_, NIG = NIG_, NI = NG_ * NIG_, NA = NA_, NE = NE_, NSTG = NSTG_;
  static constexpr int C = CK * NKB;
  static constexpr int HR = (R - 1) * S + 3;          // input rows per tile
  static constexpr int BW = (TWO - 1) * S + 3;        // input columns per tile
  static constexpr int NCHK = (HR + RC - 1) / RC;     // chunks per unit
  static constexpr int CQ = CK / 4;                   // channel quads per pixel
  static constexpr int PG = TWO / TW;                 // column groups
  static constexpr int NCOL = (TW - 1) * S + 3;       // input columns a thread reads per row
  static constexpr int RING = S == 1 ? 3 : 2;
  static constexpr uint32_t LINE = CK * 2;            // bytes per pixel in a chunk
  static constexpr uint32_t CHUNK_BYTES = (uint32_t)RC * BW * LINE;
  static constexpr uint32_t CHUNK_PITCH = (CHUNK_BYTES + 127u) & ~127u;
  static constexpr uint32_t B_BYTES = (uint32_t)COUT * 128;   // one k-block of the filter
  // More than 20 warps do not fit 96 registers each: the block is then launched with 80 per thread and setmaxnreg
  // moves registers from the single-thread roles and the epilogue to the stencil warps (REGSPLIT).  setmaxnreg is
  // warpgroup-wide, so the two single-thread roles get a warpgroup of their own (two idle warps).
  static constexpr bool REGSPLIT = 2 + NE * RB_EPI_WARPS + NG * GW > 20;
  static constexpr int WARPS = REGSPLIT ? NG * GW + NE * RB_EPI_WARPS + 4 : 2 + NE * RB_EPI_WARPS + NG * GW;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int WARPS_TMA_ROLE = NG * GW + NE * RB_EPI_WARPS + 1;   // the TMA producer's warp (W_TMA in the kernel)
  static_assert(!REGSPLIT || (GW % 4 == 0 && WARPS == 24), "register split is laid out for 16 stencil + 4 epilogue + 4 role warps");
  static constexpr int NACC = 512 / COUT > 8 ? 8 : 512 / COUT;   // TMEM accumulator stages of COUT columns
  static constexpr int NBAR = 2 * NI + 2 * NA + 2 * NACC + 1;
  static constexpr uint32_t OFF_B = 0;
  static constexpr uint32_t OFF_A = OFF_B + NKB * B_BYTES;
  static constexpr uint32_t OFF_O = OFF_A + NA * RB_A_BYTES;
  static constexpr uint32_t OFF_IN = OFF_O + NE * NSTG * RB_O_BYTES;
  static constexpr uint32_t OFF_TAPS = OFF_IN + NI * CHUNK_PITCH;
  static constexpr uint32_t OFF_DSH = OFF_TAPS + 9u * C * 4;
  static constexpr uint32_t OFF_BAR = OFF_DSH + (uint32_t)C * 4;
  static constexpr uint32_t OFF_END = OFF_BAR + 8u * NBAR + 16;
  static constexpr size_t SMEM = 1024 + OFF_END;
  static_assert(R * RB_TP <= 128 && TWO <= RB_TP, "tile does not fit one UMMA M tile");
  static_assert(TWO % TW == 0 && PG * CQ <= GW * 32, "tile does not fit a stencil group");
  static_assert(CK == 32 || CK == 64, "k-block is 32 or 64 channels");
  static_assert(COUT % 64 == 0 && COUT <= 256, "Cout: multiple of 64, at most 256 (two TMEM stages)");
  static_assert(NA >= NG && NIG >= 2, "an A stage per group at least; rings hold at least two chunks");
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
};

struct RbParams {
  const float* dw_taps;    // [9][C] taps x folded-BN scale
  const float* dw_shift;   // [C] or nullptr
  float pw_scale[256];     // folded-BN scale / shift of the pointwise layer, by value: they live in the
  float pw_shift[256];     // kernel-parameter constant bank and feed the epilogue FFMAs as immediates
  uint32_t dw_cap2, pw_cap2;
  int bands, strips, pad_lo;
  long tiles;              // n * bands * strips
  unsigned long long* trace;  // debug: per-event SM clock stamps of CTA 0 (MNV1_RB_TRACE), else nullptr
  int dbg;                 // experiment switches (MNV1_RB_DBG): 2 no TMA store, 4 no epilogue math
};

// trace[role][idx][slot]: role 0 producer, 1 MMA, 2 epilogue, 3+g stencil group g; 128 entries x 4 stamps
__device__ __forceinline__ void rb_stamp(unsigned long long* tr, int role, int idx, int slot) {
  if (tr && blockIdx.x == 0 && idx < 128) tr[(role * 128 + idx) * 4 + slot] = clock64();
}

template <class Cfg, bool DW_RELU, bool PW_RELU>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
fused_rb_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
                const __grid_constant__ RbParams p) {
  constexpr int S = Cfg::S, CK = Cfg::CK, NKB = Cfg::NKB, COUT = Cfg::COUT, TWO = Cfg::TWO, R = Cfg::R, TW = Cfg::TW;
  constexpr int RC = Cfg::RC, NG = Cfg::NG, GW = Cfg::GW, NIG = Cfg::NIG, NI = Cfg::NI, NA = Cfg::NA, NE = Cfg::NE, NSTG = Cfg::NSTG, C = Cfg::C;
  constexpr int HR = Cfg::HR, BW = Cfg::BW, NCHK = Cfg::NCHK, CQ = Cfg::CQ, PG = Cfg::PG, NCOL = Cfg::NCOL;
  constexpr int RING = Cfg::RING, NACC = Cfg::NACC;
  constexpr uint32_t LINE = Cfg::LINE;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_g = smem_raw + (smem - smem_u32(smem_raw));   // generic view of the same base
  const uint32_t sB = smem + Cfg::OFF_B, sA = smem + Cfg::OFF_A, sO = smem + Cfg::OFF_O, sIn = smem + Cfg::OFF_IN;
  const uint32_t sTaps = smem + Cfg::OFF_TAPS, sDsh = smem + Cfg::OFF_DSH;
  const uint32_t bars = smem + Cfg::OFF_BAR;
  const uint32_t in_full = bars, in_empty = in_full + 8u * NI, a_full = in_empty + 8u * NI, a_empty = a_full + 8u * NA;
  const uint32_t tm_full = a_empty + 8u * NA, tm_empty = tm_full + 8u * NACC, b_full = tm_empty + 8u * NACC, tmem_slot = b_full + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  // Warp roles.  The SM's issue arbiter favours the highest warp ids, so the latency-critical
  // single-thread roles and the epilogue sit above the bulk stencil workers.
  constexpr int W_EPI = NG * GW, W_MMA = W_EPI + NE * RB_EPI_WARPS, W_TMA = W_MMA + 1;

  // constants -> shared memory
  {
    constexpr int PER = (9 * C / 4 + Cfg::THREADS - 1) / Cfg::THREADS;
    stage_constants<PER>(reinterpret_cast<float*>(smem_g + Cfg::OFF_TAPS), p.dw_taps, 9 * C, tid, Cfg::THREADS);
    float* ds = reinterpret_cast<float*>(smem_g + Cfg::OFF_DSH);
    if (p.dw_shift) stage_constants<1>(ds, p.dw_shift, C, tid, Cfg::THREADS);
    else for (int i = tid; i < C; i += Cfg::THREADS) ds[i] = 0.f;
  }
  if (tid == 0) {
    prefetch_tmap(&tmap_in); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_out); prefetch_tmap(&tmap_out2);
    for (int s = 0; s < NI; ++s) { mbar_init(in_full + 8u * s, 1); mbar_init(in_empty + 8u * s, GW); }
    for (int s = 0; s < NA; ++s) { mbar_init(a_full + 8u * s, GW); mbar_init(a_empty + 8u * s, 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(tm_full + 8u * s, 1); mbar_init(tm_empty + 8u * s, RB_EPI_WARPS); }
    mbar_init(b_full, 1);
    mbar_init_fence();
  }
  if (MNV1_TRC(p.trace) && blockIdx.x == 0 && tid == 0) {           // debug: SM clock vs wall clock over the kernel
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    MNV1_TRC(p.trace)[(7 * 128 + 0) * 4 + 0] = clock64(); MNV1_TRC(p.trace)[(7 * 128 + 0) * 4 + 1] = gt;
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  // The pointwise filter does not depend on the previous layer: its load is started BEFORE the wait, so that it
  // travels while the last CTAs of the previous kernel are still running on other SMs.
  if (warp == Cfg::WARPS_TMA_ROLE && lane == 0) {
    mbar_expect_tx(b_full, NKB * Cfg::B_BYTES);
    for (int kb = 0; kb < NKB; ++kb) tma_load_2d(sB + kb * Cfg::B_BYTES, &tmap_b, b_full, kb * 64, 0);
  }
  pdl_wait();                                            // the previous layer's output is complete and visible

  // this CTA's tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long G = gridDim.x;
  const int nt = (int)((p.tiles - blockIdx.x + G - 1) / G);     // > 0: the grid never exceeds the tile count
  const int total_units = nt * NKB;
  const int per_img = p.bands * p.strips;

  // REGSPLIT: 768 threads x 80 registers = 61440 = 512 x 88 (stencil) + 128 x 72 (epilogue) + 128 x 56 (roles); every
  // setmaxnreg sits at the top of the branch whose code it governs (ptxas allocates per dominated region)
  if (warp >= W_MMA) {
   if constexpr (Cfg::REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == W_TMA) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage[NG]; uint32_t phase[NG];                   // per-group ring position
#pragma unroll
      for (int g = 0; g < NG; ++g) { stage[g] = 0; phase[g] = 0; }
      for (int u0 = 0; u0 < total_units; u0 += NG) {
        const int nu = total_units - u0 < NG ? total_units - u0 : NG;
        int cx[NG], cy[NG], cc[NG], ci[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int ul = u0 + (g < nu ? g : 0);
          const int lt = ul / NKB, kb = ul - lt * NKB;
          const uint32_t tile = blockIdx.x + (uint32_t)lt * (uint32_t)G;
          const uint32_t img = tile / (uint32_t)per_img, rem = tile - img * (uint32_t)per_img;
          const uint32_t band = rem / (uint32_t)p.strips, strip = rem - band * (uint32_t)p.strips;
          cx[g] = (int)strip * TWO * S - p.pad_lo; cy[g] = (int)band * R * S - p.pad_lo; cc[g] = kb * CK; ci[g] = (int)img;
        }
#pragma unroll 1
        for (int k = 0; k < NCHK; ++k) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (g < nu) {
              const uint32_t st = (uint32_t)(g * NIG + stage[g]);
              mbar_wait(in_empty + 8u * st, phase[g] ^ 1u);
              mbar_expect_tx(in_full + 8u * st, Cfg::CHUNK_BYTES);
              tma_load_4d(sIn + st * Cfg::CHUNK_PITCH, &tmap_in, in_full + 8u * st, cc[g], cx[g], cy[g] + k * RC, ci[g]);
              if (k == 0) rb_stamp(MNV1_TRC(p.trace), 0, u0 + g, 0); else if (k == NCHK - 1) rb_stamp(MNV1_TRC(p.trace), 0, u0 + g, 1);
              if (++stage[g] == NIG) { stage[g] = 0; phase[g] ^= 1u; }
            }
          }
        }
      }
    }
   } else if (warp == W_MMA) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(COUT);
      mbar_wait(b_full, 0);
      tc_fence_after();
      int ul = 0;
      for (int lt = 0; lt < nt; ++lt) {
        const uint32_t acc = (uint32_t)lt % NACC;
        mbar_wait(tm_empty + 8u * acc, (((uint32_t)lt / NACC) & 1u) ^ 1u);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)COUT;
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb, ++ul) {
          const uint32_t st = (uint32_t)ul % NA, ph = ((uint32_t)ul / NA) & 1u;
          rb_stamp(MNV1_TRC(p.trace), 1, ul, 0);
          mbar_wait(a_full + 8u * st, ph);
          tc_fence_after();
          rb_stamp(MNV1_TRC(p.trace), 1, ul, 1);
          const uint64_t da = umma_desc_sw128(sA + st * RB_A_BYTES);
          const uint64_t db = umma_desc_sw128(sB + kb * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < CK / 16; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(a_empty + 8u * st);
          rb_stamp(MNV1_TRC(p.trace), 1, ul, 2);
        }
        umma_commit(tm_full + 8u * acc);
      }
    }
   }
  } else if (warp >= W_EPI) {
    // ======================= epilogue warps =======================
    if constexpr (Cfg::REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    // A warp owns TMEM lanes 32*quarter .. +31 = tile rows 2*quarter, 2*quarter+1 and works on its own:
    // private staging (NSTG x 4 KB) and its own TMA stores, so the four warps never meet.
    const int quarter = warp & 3;                          // TMEM lane quarter = warp % 4
    const int rsel = lane >> 4, mx = lane & 15;            // lane = rsel*TP + x
    const bool valid = mx < TWO && 2 * quarter + rsel < R;
    const int eset = (warp - W_EPI) >> 2;                   // epilogue set: tiles eset, eset + NE, ...
    const uint32_t wbuf = sO + (uint32_t)(warp - W_EPI) * (NSTG * 4096u);
    const uint32_t line = (uint32_t)(rsel * TWO + mx);      // dense [2][TWO] lines of 128 B, as the 2-row TMA box reads them
    const uint32_t line_off = line * 128u, line_x = line & 7u;
    const bool tracer = warp == W_EPI && lane == 0;
    const int row0 = 2 * quarter;                          // first tile row of this warp
    uint32_t blk = 0;
    for (int lt = eset; lt < nt; lt += NE) {
      const uint32_t tile = blockIdx.x + (uint32_t)lt * (uint32_t)G;
      const uint32_t img = tile / (uint32_t)per_img, rem = tile - img * (uint32_t)per_img;
      const uint32_t band = rem / (uint32_t)p.strips, strip = rem - band * (uint32_t)p.strips;
      const uint32_t acc = (uint32_t)lt % NACC;
      if (tracer) rb_stamp(MNV1_TRC(p.trace), 2, lt, 0);
      mbar_wait(tm_full + 8u * acc, ((uint32_t)lt / NACC) & 1u);
      tc_fence_after();
      if (tracer) rb_stamp(MNV1_TRC(p.trace), 2, lt, 1);
#pragma unroll
      for (int b = 0; b < COUT / 64; ++b, ++blk) {
        const uint32_t sbuf = wbuf + (NSTG == 1 ? 0u : (blk % NSTG) * 4096u);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (uint32_t)COUT + (uint32_t)(b * 64);
        uint32_t v[32];
        if (!(MNV1_DBG(p.dbg) & 4)) tmem_ld32_nowait(taddr, v);
        if (lane == 0) tma_store_wait_read<NSTG - 1>();      // the store that last read this buffer is done with it
        __syncwarp();
        if (!(MNV1_DBG(p.dbg) & 4)) {
          // two 32-column halves through one set of 32 registers (the kernel is capped at 96 registers:
          // 18 warps are allocated as 20); scale / shift are compile-time offsets into the constant bank
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (tracer && b == 0 && half == 0) rb_stamp(MNV1_TRC(p.trace), 6, lt, 0);
            tmem_ld_wait();
            if (tracer && b == 0 && half == 0) rb_stamp(MNV1_TRC(p.trace), 6, lt, 1);
            uint32_t q[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int col = b * 64 + half * 32 + j;
              // one FFMA2 per output pair: scale / shift pairs straight from the constant bank
              const f32x2 acc2 = f2_pack(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              q[j / 2] = pack2_f2<PW_RELU>(f2_fma(acc2, f2_pack(p.pw_scale[col], p.pw_scale[col + 1]),
                                                  f2_pack(p.pw_shift[col], p.pw_shift[col + 1])), p.pw_cap2);
            }
            if (half == 0) tmem_ld32_nowait(taddr + 32u, v);   // v is free again: fetch the second half under the stores
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
              if (valid) sts128(sbuf + line_off + (((uint32_t)(half * 4 + c4) ^ line_x) << 4), q[4 * c4], q[4 * c4 + 1], q[4 * c4 + 2], q[4 * c4 + 3]);
          }
        }
        if (tracer && b == 0) rb_stamp(MNV1_TRC(p.trace), 6, lt, 2);
        fence_proxy_async();
        __syncwarp();
        if (tracer && b == 0) rb_stamp(MNV1_TRC(p.trace), 6, lt, 3);
        if (lane == 0 && !(MNV1_DBG(p.dbg) & 2)) {
          // the warp's two tile rows as one box (64 channels x TWO columns x 2 rows; smem = 2 x [16][64]
          // bf16, 128B swizzle), or the single-row box when the tile has an odd number of rows
          const int y = (int)band * R + row0;
          if (row0 + 1 < R) tma_store_4d(&tmap_out2, sbuf, b * 64, (int)strip * TWO, y, (int)img);
          else if (row0 < R) tma_store_4d(&tmap_out, sbuf, b * 64, (int)strip * TWO, y, (int)img);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tm_empty + 8u * acc);
      if (tracer) rb_stamp(MNV1_TRC(p.trace), 2, lt, 2);
    }
    if (lane == 0) tma_store_wait_all();
  } else {
    // ======================= stencil groups =======================
    if constexpr (Cfg::REGSPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int g = warp / GW;
    const int t = tid - g * GW * 32;                      // thread index inside the group
    const bool active = t < PG * CQ;
    const int quad = t % CQ;
    const int pg = active ? t / CQ : PG - 1;
    // byte offset of the thread's first input column / channel quad inside a chunk
    const uint32_t in_off = (uint32_t)(pg * TW * S) * LINE + (uint32_t)quad * 8u;
    // A operand: row = r*TP + x (128 B each), 16-byte chunk XOR (row & 7) -- TP % 8 == 0, so the
    // swizzle term depends on the column only
    uint32_t a_off[TW];
#pragma unroll
    for (int c = 0; c < TW; ++c) {
      const uint32_t x = (uint32_t)(pg * TW + c);
      a_off[c] = x * 128u + ((((uint32_t)quad >> 1) ^ (x & 7u)) << 4) + ((uint32_t)quad & 1u) * 8u;
    }
    f32x2 w[9][2], sh[2];                                   // the thread's 4 channels as two packed fp32 pairs
    auto load_taps = [&](int kb) {
      const uint32_t ch = (uint32_t)(kb * CK + quad * 4) * 4u;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 a = lds128f(sTaps + (uint32_t)k * C * 4u + ch);
        w[k][0] = f2_pack(a.x, a.y); w[k][1] = f2_pack(a.z, a.w);
      }
      const float4 a = lds128f(sDsh + ch);
      sh[0] = f2_pack(a.x, a.y); sh[1] = f2_pack(a.z, a.w);
    };
    if (NKB == 1) load_taps(0);
    uint32_t rstage = 0, rphase = 0;                      // position in the group's private chunk ring

    for (int u0 = 0; u0 < total_units; u0 += NG) {
      const int nu = total_units - u0 < NG ? total_units - u0 : NG;
      if (g >= nu) break;
      const int ul = u0 + g;
      if (NKB > 1) load_taps(ul % NKB);
      const uint32_t ast = (uint32_t)ul % NA, aph = ((uint32_t)ul / NA) & 1u;
      const uint32_t dstA = sA + ast * RB_A_BYTES;

      const bool tracer = t == 0;
      if (tracer) rb_stamp(MNV1_TRC(p.trace), 3 + g, ul / NG, 0);
      f32x2 acc[RING][TW][2];
      // The raw bf16 quads of input row q+1 are fetched while row q is being computed (one row of
      // software pipelining: without it every row exposes the ld.shared latency to its FFMAs).
      uint32_t rowbase = 0, cur_stage = 0, prev_stage = 0;
      uint2 nraw[NCOL];
      auto fetch_row = [&](int q) {                          // q is a compile-time constant at every call site
        if (q % RC == 0) {                                   // first row of the next chunk of the unit
          prev_stage = cur_stage;
          cur_stage = (uint32_t)(g * NIG) + rstage;
          mbar_wait_relaxed(in_full + 8u * cur_stage, rphase);
          if (q == 0 && tracer) rb_stamp(MNV1_TRC(p.trace), 3 + g, ul / NG, 1);
          rowbase = sIn + cur_stage * Cfg::CHUNK_PITCH + in_off;
          if (++rstage == NIG) { rstage = 0; rphase ^= 1u; }
        }
#pragma unroll
        for (int j = 0; j < NCOL; ++j) nraw[j] = lds64(rowbase + (uint32_t)(((q % RC) * BW + j) * LINE));
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
        // input row q feeds tap row tr of output row o = (q - tr) / S; tap rows accumulate in order
        // 0, 1, 2 (the order of depthwise_tma.cu), the shift seeds the accumulator
#pragma unroll
        for (int tr = 2; tr >= 0; --tr) {
          if ((q - tr) >= 0 && (q - tr) % S == 0 && (q - tr) / S < R ) {
            const int o = (q - tr) / S, slot = o % RING;
#pragma unroll
            for (int c = 0; c < TW; ++c)
#pragma unroll
              for (int v = 0; v < 2; ++v) {   // FFMA2: two channels per instruction
                const f32x2 init = tr == 0 ? sh[v] : acc[slot][c][v];
                acc[slot][c][v] = f2_fma(x[c * S + 2][v], w[3 * tr + 2][v],
                                       f2_fma(x[c * S + 1][v], w[3 * tr + 1][v], f2_fma(x[c * S][v], w[3 * tr][v], init)));
              }
            if (tr == 2) {                                  // output row o is complete
              if (o == 0) { mbar_wait_relaxed(a_empty + 8u * ast, aph ^ 1u); if (tracer) rb_stamp(MNV1_TRC(p.trace), 3 + g, ul / NG, 2); }   // the MMAs that last read this A stage retired
              if (active) {
#pragma unroll
                for (int c = 0; c < TW; ++c)
                  sts64(dstA + a_off[c] + (uint32_t)(o * RB_TP) * 128u, pack2_f2<DW_RELU>(acc[slot][c][0], p.dw_cap2),
                        pack2_f2<DW_RELU>(acc[slot][c][1], p.dw_cap2));
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8u * ast);
      if (tracer) rb_stamp(MNV1_TRC(p.trace), 3 + g, ul / NG, 3);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (MNV1_TRC(p.trace) && blockIdx.x == 0 && tid == 0) {
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    MNV1_TRC(p.trace)[(7 * 128 + 0) * 4 + 2] = clock64(); MNV1_TRC(p.trace)[(7 * 128 + 0) * 4 + 3] = gt;
  }
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

template <class Cfg>
cudaError_t launch_rb(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n, int H, int W,
                      int pad_lo, int num_sms, cudaStream_t st, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  constexpr int S = Cfg::S, C = Cfg::C, COUT = Cfg::COUT;
  const int Ho = H / S, Wo = W / S;
  if (Ho % Cfg::R || Wo % Cfg::TWO) return cudaErrorNotSupported;
  CUtensorMap tin, tout, tout2;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)Cfg::CK, (cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::RC, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "fused block: input tensor map encode failed"; return cudaErrorInvalidValue; }
  }
  for (int rows2 = 1; rows2 <= 2; ++rows2) {   // output boxes of one and of two tile rows
    cuuint64_t gdim[4] = {(cuuint64_t)COUT, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)n};
    cuuint64_t gstr[3] = {(cuuint64_t)COUT * 2, (cuuint64_t)Wo * COUT * 2, (cuuint64_t)Ho * Wo * COUT * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Cfg::TWO, (cuuint32_t)rows2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(rows2 == 1 ? &tout : &tout2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "fused block: output tensor map encode failed"; return cudaErrorInvalidValue; }
  }
  RbParams p{};
  p.dw_taps = dw->w_scaled; p.dw_shift = dw->shift;
  for (int i = 0; i < 256; ++i) {
    p.pw_scale[i] = i < (int)pw->h_scale.size() ? pw->h_scale[i] : 1.f;
    p.pw_shift[i] = i < (int)pw->h_shift.size() ? pw->h_shift[i] : 0.f;
  }
  p.dw_cap2 = dw->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.pw_cap2 = pw->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.bands = Ho / Cfg::R; p.strips = Wo / Cfg::TWO; p.pad_lo = pad_lo;
  p.tiles = (long)n * p.bands * p.strips;
#ifdef MNV1_TRACE   // libmnv1_trace.so only (make trace)
  p.dbg = getenv("MNV1_RB_DBG") ? atoi(getenv("MNV1_RB_DBG")) : 0;
  static unsigned long long* d_trace = nullptr;
  const bool tracing = !capturing(st) && getenv("MNV1_RB_TRACE") != nullptr;
  if (tracing) {
    if (!d_trace) cudaMalloc(&d_trace, 8 * 128 * 4 * 8);
    cudaMemsetAsync(d_trace, 0, 8 * 128 * 4 * 8, st);
    p.trace = d_trace;
  }
#endif
  const bool dr = dw->act != MNV1_ACT_NONE, pr = pw->act != MNV1_ACT_NONE;
  long grid = num_sms;
  if (grid > p.tiles) grid = p.tiles;
  {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f) { if (e == cudaSuccess) e = ensure_dyn_smem((const void*)f, (int)Cfg::SMEM); };
    set((const void*)fused_rb_kernel<Cfg, true, true>); set((const void*)fused_rb_kernel<Cfg, true, false>);
    set((const void*)fused_rb_kernel<Cfg, false, true>); set((const void*)fused_rb_kernel<Cfg, false, false>);
    if (e != cudaSuccess) return e;
  }
  cudaError_t le;
#define RB_LAUNCH(A, B) le = launch_pdl(fused_rb_kernel<Cfg, A, B>, dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::SMEM, st, tin, pw->tmap_b, tout, tout2, p)
  if (dr) { if (pr) RB_LAUNCH(true, true); else RB_LAUNCH(true, false); }
  else    { if (pr) RB_LAUNCH(false, true); else RB_LAUNCH(false, false); }
#undef RB_LAUNCH
  if (le != cudaSuccess) return le;
#ifdef MNV1_TRACE
  if (tracing) {   // debug only: dump the stamps of the last launch
    std::vector<unsigned long long> h(8 * 128 * 4);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("MNV1_RB_TRACE"), "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
  }
#endif
  return cudaGetLastError();
}

//                   S  CK NKB COUT TWO R TW  RC NG GW NIG NA NE NSTG
using CfgL02 = RbCfg<1, 32, 1,  64, 16, 8, 2, 10, 4, 2, 2, 4, 2, 1>;   // 112x112x32  -> 112x112x64
using CfgL04 = RbCfg<2, 64, 1, 128, 14, 8, 2,  2, 3, 4, 5, 3, 1, 2>;   // 112x112x64  -> 56x56x128
using CfgL06 = RbCfg<1, 64, 2, 128, 14, 8, 2,  5, 3, 4, 3, 3, 1, 2>;   // 56x56x128   -> 56x56x128
// (four stencil groups = 16 warps at 88 registers through the REGSPLIT path measured 121-134 us against 107 us for
//  this three-group configuration, eager timing: the stencil is not short of warps; experiments/README.md)
using CfgL08 = RbCfg<2, 64, 2, 256, 14, 7, 2,  2, 3, 4, 4, 3, 1, 1>;   // 56x56x128   -> 28x28x256
using CfgL10 = RbCfg<1, 64, 4, 256, 14, 7, 2,  3, 2, 4, 3, 2, 1, 1>;   // 28x28x256   -> 28x28x256

}  // namespace

namespace {
// index of the variant that handles this block, or -1
int rb_variant(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride) {
  if (!dw->w_scaled || !pw->has_tmap || pw->cin != dw->cout || pw->tmap_bn != pw->cout) return -1;
  const long mask = switches().rb_mask;   // MNV1_RB_MASK: bit i enables the i-th variant below; default all
  const int c = dw->cout, co = pw->cout;
#define RB_IS(CFG, HW, BIT) \
  if ((mask >> BIT & 1) && stride == CFG::S && c == CFG::C && co == CFG::COUT && rows == HW && cols == HW && \
      (HW / CFG::S) % CFG::R == 0 && (HW / CFG::S) % CFG::TWO == 0) return BIT
  RB_IS(CfgL02, 112, 0);
  RB_IS(CfgL04, 112, 1);
  RB_IS(CfgL06, 56, 2);
  RB_IS(CfgL08, 56, 3);
  RB_IS(CfgL10, 28, 4);
#undef RB_IS
  return -1;
}
}  // namespace

bool fused_dw_pw_supported(const mnv1_filter* dw, const mnv1_filter* pw, int rows, int cols, int stride) {
  return rb_variant(dw, pw, rows, cols, stride) >= 0;
}

// cudaErrorNotSupported (nothing launched) when the block has no fused variant (the filter must fit in
// shared memory: layers 2-11; the 512- and 1024-channel blocks run as two kernels).
cudaError_t launch_fused_dw_pw(bf16* out, const bf16* in, const mnv1_filter* dw, const mnv1_filter* pw, int n, int rows,
                               int cols, int stride, int pad_lo, int num_sms, cudaStream_t st, std::string* err) {
  const int v = rb_variant(dw, pw, rows, cols, stride);
  if (v < 0) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
  switch (v) {
    case 0: return launch_rb<CfgL02>(out, in, dw, pw, n, rows, cols, pad_lo, num_sms, st, err);
    case 1: return launch_rb<CfgL04>(out, in, dw, pw, n, rows, cols, pad_lo, num_sms, st, err);
    case 2: return launch_rb<CfgL06>(out, in, dw, pw, n, rows, cols, pad_lo, num_sms, st, err);
    case 3: return launch_rb<CfgL08>(out, in, dw, pw, n, rows, cols, pad_lo, num_sms, st, err);
    case 4: return launch_rb<CfgL10>(out, in, dw, pw, n, rows, cols, pad_lo, num_sms, st, err);
  }
  return cudaErrorNotSupported;
}

}  // namespace mnv1
