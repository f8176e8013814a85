// pointwise_pair.cu — the 1x1 convolution for the tensor-bound layers (Cout a multiple of 256, K >= 256:
// layers 13-27 of the MobileNet.c schedule) on CTA PAIRS: tcgen05.mma.cta_group::2, one 256 x 256 tile per pair.
//
// Same contract and arithmetic as pointwise_tc.cu (`pointwise`, kernel.cl:94-114).  Why a second kernel:
// on the 512x512 layers a 128x256x512 tile holds 4.1 k cycles of UMMA work but pointwise_tc needs 6-7.5 k,
// and experiments/mma_fill.cu says why.  (1) With every SM pulling from L2, shared memory fills at no more
// than ~84 B/cycle/SM, and only with >= 3 copies of 32 KB in flight; a 128x256 tile needs 48 KB per
// 512-cycle k-block = 94 B/cycle, and the 3 x 48 KB ring that fits keeps 2 copies in flight (~65 B/cycle).
// (2) One thread starts a TMA load every ~380 cycles, so two loads per k-block from one thread are
// issue-bound at ~760 cycles.  A CTA pair fixes both: each CTA stages its own 128 rows of A and HALF of
// the filter tile (the pair's UMMA reads the other half from the peer), 32 KB per k-block = 62 B/cycle
// with a 4-deep ring, and each operand has its own issuing thread.
//
//   warp 0      A producer: this CTA's 128 x 64 activation tile per k-block; the leader's thread also
//               announces the k-block's bytes of BOTH CTAs on the leader's full barrier
//   warp 10     filter producer: this CTA's half (128 output channels x 64) of the filter tile
//   warp 1      TMEM allocator; in the leader CTA the converged MMA issuer (one elected lane,
//               UMMA 256 x 256 x 16), tcgen05.commit multicast to the barriers of both CTAs
//   warps 2-9   epilogue: two warps per TMEM lane quarter, alternate 64-column blocks, private
//               staging, own TMA stores
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

constexpr int MC_BK = 64;
constexpr int MC_EPI_WARPS = 8;
constexpr int MC_THREADS = 64 + 32 * MC_EPI_WARPS + 32;
constexpr int MC_W_BPROD = 2 + MC_EPI_WARPS;
constexpr int MC_MAX_STAGES = 8;
constexpr uint32_t MC_A_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr uint32_t MC_BH_BYTES = 128 * 128;       // half of the filter tile: 128 output channels x 64 bf16
constexpr uint32_t MC_STAGE_BYTES = MC_A_BYTES + MC_BH_BYTES;   // per CTA: its A tile + its half of the filter tile
constexpr uint32_t MC_ACC_COLS = 256;

// debug: SM-clock stamps of CTA 0 (MNV1_PW_TRACE=<file>, tools/pw_trace.py): trace[role][idx][slot]
__device__ __forceinline__ void pp_stamp(unsigned long long* tr, int role, long idx, int slot) {
  if (tr && blockIdx.x == 0 && idx < 128) tr[(role * 128 + idx) * 4 + slot] = clock64();
}

struct McParams {
  const float* scale;
  const float* shift;
  uint32_t cap2;
  long M;
  int K, Cout, stages;
  // Tail balancing: units [0, full_units) are whole 256 x 256 tiles; when the last round would be less than half full, its
  // `tail_halves / 2` units are cut into 256 x 128 halves (N = 128 UMMAs) spread over twice as many clusters, so the
  // launch ends after half a tile time instead of a whole one (392 units on 74 pairs: 5.5 rounds instead of 6).
  long full_units;
  int tail_halves;
  unsigned long long* trace;
  bf16* out;
  int dbg;      // MNV1_PP_DBG (timing experiments only): 1 = st.global from registers instead of staging + TMA store, 2 = no operand loads, 4 = no epilogue output
};

// DIRECT: the epilogue stores its bf16 rows straight from registers with 256-bit st.global (32 B = one full sector
// per thread and instruction: no partial-sector writes) instead of staging them in shared memory for a TMA store —
// the staging costs 128 KB of shared-memory traffic per tile (64 KB written, 64 KB read back by the TMA unit) on a
// kernel whose shared-memory data pipe is ~93 % busy (profiles/r01_pp_l15_ncu.txt), and its 64 KB buy two more ring stages.
template <bool RELU, bool DIRECT>
__global__ void __launch_bounds__(MC_THREADS, 1)
pointwise_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_b64, const __grid_constant__ CUtensorMap tmap_out, const McParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_g = smem_raw + (smem - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t sRing = smem;
  const uint32_t sOut = sRing + (uint32_t)stages * MC_STAGE_BYTES;          // 8 warps x 2 x 4 KB (not DIRECT)
  const uint32_t sScale = sOut + (DIRECT ? 0u : MC_EPI_WARPS * 8192u);      // [Cout] fp32
  const uint32_t sShift = sScale + (uint32_t)p.Cout * 4u;
  const uint32_t bars = sShift + (uint32_t)p.Cout * 4u;
  const uint32_t full = bars, empty = full + 8u * MC_MAX_STAGES, tm_full = empty + 8u * MC_MAX_STAGES;
  const uint32_t tm_empty = tm_full + 16, tmem_slot = tm_empty + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const uint32_t rank = cluster_ctarank();
  const long cid = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_kb = p.K / MC_BK;
  const int n_tiles = p.Cout / 256;
  const long m_tiles = (p.M + 127) / 128, m_pairs = (m_tiles + 1) / 2;
  // a unit = (pair of m-tiles, n-tile); this CTA's m-tile is 2*pair + rank.  Work item `it` of this cluster:
  struct Work { long mp; int n_base, ncols; };
  auto get_work = [&](long it, Work& w) -> bool {
    const long u = cid + it * num_clusters;
    if (u < p.full_units) { w.mp = u / n_tiles; w.n_base = (int)(u % n_tiles) * 256; w.ncols = 256; return true; }
    const long hu = u - p.full_units;                      // the round after the last full one: half units
    if (hu >= p.tail_halves) return false;
    const long unit = p.full_units + (hu >> 1);
    w.mp = unit / n_tiles; w.n_base = (int)(unit % n_tiles) * 256 + (int)(hu & 1) * 128; w.ncols = 128;
    return true;
  };
  (void)m_pairs;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_b64); prefetch_tmap(&tmap_out);
    for (int s = 0; s < stages; ++s) { mbar_init(full + 8u * s, 1); mbar_init(empty + 8u * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tm_full + 8u * a, 1); mbar_init(tm_empty + 8u * a, 2 * MC_EPI_WARPS); }
    mbar_init_fence();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * MC_ACC_COLS);
  {  // folded-BN scale / shift -> shared memory, before the wait below: filter constants do not depend on the previous layer
    const int et = threadIdx.x - 64, n4 = p.Cout >> 2;
    if (et >= 0 && et < n4) {
      const float4 sv = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale) + et) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 tv = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift) + et) : make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(smem_g + (sScale - smem))[et] = sv;
      reinterpret_cast<float4*>(smem_g + (sShift - smem))[et] = tv;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  pdl_wait();                                     // the previous layer's output is complete and visible

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && !(MNV1_DBG(p.dbg) & 2)) {
      const uint32_t full_leader = mapa_shared(full, 0);
      int stage = 0; uint32_t phase = 0;
      Work w;
      for (long it = 0; get_work(it, w); ++it) {
        const int m_idx = (int)(w.mp * 2 + rank) * 128;
        const uint32_t stage_tx = 2 * (MC_A_BYTES + (uint32_t)(w.ncols / 2) * 128u);   // both CTAs' A tile + filter half
        for (int kb = 0; kb < num_kb; ++kb) {
          if (kb == 0) pp_stamp(MNV1_TRC(p.trace), 0, it, 0);
          mbar_wait(empty + 8u * stage, phase ^ 1u);       // the pair's MMAs that read this slot have retired
          const uint32_t sa = sRing + (uint32_t)stage * MC_STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(full + 8u * stage, stage_tx);   // both CTAs' bytes complete on the leader's barrier
          tma_load_2d_pair(sa, &tmap_a, full_leader + 8u * stage, kb * MC_BK, m_idx);
          if (kb == num_kb - 1) pp_stamp(MNV1_TRC(p.trace), 0, it, 2);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, one elected lane) =================
    const uint32_t elected = rank == 0 ? elect_one() : 0u;
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    Work w;
    for (long it = 0; rank == 0 && get_work(it, w); ++it) {   // the leader issues for the pair
      const long u = cid + it * num_clusters;
      const uint32_t idesc = w.ncols == 256 ? umma_idesc_bf16_m256(256) : umma_idesc_bf16_m256(128);
      if (elected) pp_stamp(MNV1_TRC(p.trace), 1, u / num_clusters, 0);
      mbar_wait(tm_empty + 8u * as, aphase ^ 1u);          // epilogue has drained this accumulator stage
      tc_fence_after();
      if (elected) pp_stamp(MNV1_TRC(p.trace), 1, u / num_clusters, 1);
      const uint32_t tmem_d = tmem_base + (uint32_t)as * MC_ACC_COLS;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!(MNV1_DBG(p.dbg) & 2)) mbar_wait(full + 8u * stage, phase);
        tc_fence_after();
        if (kb == 0 && elected) pp_stamp(MNV1_TRC(p.trace), 1, u / num_clusters, 2);
        const uint32_t sa = sRing + (uint32_t)stage * MC_STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + MC_A_BYTES);
#pragma unroll
        for (int k = 0; k < MC_BK / 16; ++k) umma_pair_if(elected, tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit_pair_if(elected, empty + 8u * stage);                  // frees the slot in both CTAs
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit_pair_if(elected, tm_full + 8u * as);                     // both CTAs' accumulators are complete
      if (elected) pp_stamp(MNV1_TRC(p.trace), 1, u / num_clusters, 3);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp == MC_W_BPROD) {
    // ================= filter producer =================
    // Its bytes are counted by the leader's expect_tx (a complete_tx that lands first only makes the
    // count negative for a while: the phase cannot complete before that arrival).
    if (lane == 0 && !(MNV1_DBG(p.dbg) & 2)) {
      const uint32_t full_leader = mapa_shared(full, 0);
      int stage = 0; uint32_t phase = 0;
      Work w;
      for (long it = 0; get_work(it, w); ++it) {
        const int n_idx = w.n_base + (int)rank * (w.ncols / 2);           // the half of the filter (sub)tile this CTA holds
        const CUtensorMap* tb = w.ncols == 256 ? &tmap_b : &tmap_b64;     // box of 128 or 64 filter rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty + 8u * stage, phase ^ 1u);
          tma_load_2d_pair(sRing + (uint32_t)stage * MC_STAGE_BYTES + MC_A_BYTES, tb, full_leader + 8u * stage, kb * MC_BK, n_idx);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ================= epilogue warps 2..9: independent, private staging =================
    const int quarter = warp & 3;                        // TMEM lanes 32*quarter .. +31
    const int h = (warp - 2) >> 2;                       // which 64-column blocks: b = h, h+2
    const uint32_t wbuf = sOut + (uint32_t)(warp - 2) * 8192u;
    const uint32_t row_off = (uint32_t)lane * 128u, row_x = (uint32_t)(lane & 7);
    const uint32_t tm_empty_leader = mapa_shared(tm_empty, 0);
    int as = 0; uint32_t aphase = 0;
    uint32_t blk = 0;
    Work w;
    for (long it = 0; get_work(it, w); ++it) {
      const long u = cid + it * num_clusters;
      const int m_idx = (int)(w.mp * 2 + rank) * 128 + quarter * 32;
      const int n_idx = w.n_base;
      const int nblk = w.ncols / 64;
      if (threadIdx.x == 64) pp_stamp(MNV1_TRC(p.trace), 2, u / num_clusters, 0);
      mbar_wait(tm_full + 8u * as, aphase);
      tc_fence_after();
      if (threadIdx.x == 64) pp_stamp(MNV1_TRC(p.trace), 2, u / num_clusters, 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)as * MC_ACC_COLS;
#pragma unroll 1
      for (int b = h; b < nblk; b += 2) {
        const uint32_t sbuf = wbuf + (blk & 1u) * 4096u;
        ++blk;
        uint32_t v[32];
        tmem_ld32_nowait(taddr + (uint32_t)(b * 64), v);
        if (!DIRECT) {
          if (lane == 0) tma_store_wait_read<1>();         // the store issued two blocks ago has consumed this buffer
          __syncwarp();
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          // this half's 32 scale / shift values first: their latency hides under the TMEM load
          const uint32_t col = (uint32_t)(n_idx + b * 64 + 32 * half);
          float4 s4[8], t4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { s4[j] = lds128f(sScale + (col + 4 * j) * 4u); t4[j] = lds128f(sShift + (col + 4 * j) * 4u); }
          tmem_ld_wait();
          uint32_t q[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            q[2 * j] = pack2<RELU>(fmaf(__uint_as_float(v[4 * j + 0]), s4[j].x, t4[j].x), fmaf(__uint_as_float(v[4 * j + 1]), s4[j].y, t4[j].y), p.cap2);
            q[2 * j + 1] = pack2<RELU>(fmaf(__uint_as_float(v[4 * j + 2]), s4[j].z, t4[j].z), fmaf(__uint_as_float(v[4 * j + 3]), s4[j].w, t4[j].w), p.cap2);
          }
          if (half == 0) tmem_ld32_nowait(taddr + (uint32_t)(b * 64 + 32), v);   // second half under the stores
          if (MNV1_DBG(p.dbg) & 4) {
          } else if (DIRECT) {
            const long row = (long)m_idx + lane;
            if (row < p.M) {
              bf16* gp = p.out + row * p.Cout + n_idx + b * 64 + half * 32;
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gp), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]),
                           "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gp + 16), "r"(q[8]), "r"(q[9]), "r"(q[10]), "r"(q[11]),
                           "r"(q[12]), "r"(q[13]), "r"(q[14]), "r"(q[15]) : "memory");
            }
          } else {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            sts128(sbuf + row_off + (((uint32_t)(4 * half + c4) ^ row_x) << 4), q[4 * c4], q[4 * c4 + 1], q[4 * c4 + 2], q[4 * c4 + 3]);
          }
        }
        if (DIRECT || (MNV1_DBG(p.dbg) & 4)) continue;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {   // box = 64 columns x 32 rows; rows past M are clipped by the TMA unit
          tma_store_2d(&tmap_out, sbuf, n_idx + b * 64, m_idx);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tm_empty_leader + 8u * as);
      if (threadIdx.x == 64) pp_stamp(MNV1_TRC(p.trace), 2, u / num_clusters, 2);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (!DIRECT && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer may still multicast into this CTA's ring / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * MC_ACC_COLS);
  }
}

cudaError_t encode_box(CUtensorMap* map, const void* base, uint64_t rows, uint64_t k, uint32_t box_rows, std::string* err) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) { if (err) *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  cuuint64_t gdim[2] = {k, rows};
  cuuint64_t gstride[1] = {k * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { if (err) *err = "pointwise_pair: tensor map encode failed"; return cudaErrorInvalidValue; }
  return cudaSuccess;
}

}  // namespace

// cudaErrorNotSupported (nothing launched) when the shape has no cluster variant.
cudaError_t launch_pointwise_pair(bf16* out, const bf16* in, const mnv1_filter* f, long m, int k, int cout, int num_sms,
                                cudaStream_t st, std::string* err) {
  // K = 256 (layer 13) stays on the single-CTA kernel: with four k-blocks per tile the pair's cluster hand-shakes cost more
  // than the halved filter traffic saves (in the graph 16.9 vs 18.7 us; MNV1_PAIR_MIN_K=256 brings the pair kernel back)
#ifndef MNV1_PAIR_MIN_K
#define MNV1_PAIR_MIN_K 512
#endif
  if (switches().no_pair || !f->w_bf16 || cout % 256 || k % MC_BK || k < MNV1_PAIR_MIN_K || cout > 1024 || num_sms < 2) return cudaErrorNotSupported;
  if (m <= 0) return cudaSuccess;
  CUtensorMap ta, tb, tb64, to;
  cudaError_t e = encode_box(&ta, in, (uint64_t)m, (uint64_t)k, 128, err);
  if (e == cudaSuccess) e = encode_box(&tb, f->w_bf16, (uint64_t)cout, (uint64_t)k, 128, err);
  if (e == cudaSuccess) e = encode_box(&tb64, f->w_bf16, (uint64_t)cout, (uint64_t)k, 64, err);
  if (e == cudaSuccess) e = encode_box(&to, out, (uint64_t)m, (uint64_t)cout, 32, err);
  if (e != cudaSuccess) return e;
  const bool direct = switches().pp_direct;
  const size_t fixed = 1024 + (direct ? 0 : MC_EPI_WARPS * 8192) + (size_t)cout * 8 + 16 * MC_MAX_STAGES + 64;
  int stages = (int)((227 * 1024 - fixed) / MC_STAGE_BYTES);
  if (stages > MC_MAX_STAGES) stages = MC_MAX_STAGES;
  if (stages < 2) return cudaErrorNotSupported;
  const size_t smem = fixed + (size_t)stages * MC_STAGE_BYTES;
  {
    e = ensure_dyn_smem((const void*)pointwise_pair_kernel<true, false>, 227 * 1024);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)pointwise_pair_kernel<false, false>, 227 * 1024);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)pointwise_pair_kernel<true, true>, 227 * 1024);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)pointwise_pair_kernel<false, true>, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  McParams p{};
  p.scale = f->scale; p.shift = f->shift;
  p.cap2 = f->act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  p.M = m; p.K = k; p.Cout = cout; p.stages = stages;
  p.out = out;
#ifdef MNV1_TRACE   // experiment switches / pipeline stamps exist in libmnv1_trace.so only (make trace)
  { static const char* de = getenv("MNV1_PP_DBG"); p.dbg = de ? atoi(de) : 0; }
  static unsigned long long* d_trace_buf = nullptr;
  const char* trace_path = capturing(st) ? nullptr : getenv("MNV1_PW_TRACE");
  if (trace_path) {
    if (!d_trace_buf) cudaMalloc(&d_trace_buf, 4 * 128 * 4 * 8);
    cudaMemsetAsync(d_trace_buf, 0, 4 * 128 * 4 * 8, st);
    p.trace = d_trace_buf;
  }
#endif
  const long m_tiles = (m + 127) / 128, units = ((m_tiles + 1) / 2) * (cout / 256);
  long clusters = num_sms / 2;
  if (clusters > units) clusters = units;
  {
    const long rem = units % clusters;
    const bool split = rem > 0 && 2 * rem <= clusters && !switches().no_pp_tail;
    p.full_units = split ? units - rem : units;
    p.tail_halves = split ? (int)(2 * rem) : 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(MC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (direct)
    e = f->act != MNV1_ACT_NONE ? cudaLaunchKernelEx(&cfg, pointwise_pair_kernel<true, true>, ta, tb, tb64, to, p)
                                : cudaLaunchKernelEx(&cfg, pointwise_pair_kernel<false, true>, ta, tb, tb64, to, p);
  else
    e = f->act != MNV1_ACT_NONE ? cudaLaunchKernelEx(&cfg, pointwise_pair_kernel<true, false>, ta, tb, tb64, to, p)
                                : cudaLaunchKernelEx(&cfg, pointwise_pair_kernel<false, false>, ta, tb, tb64, to, p);
#ifdef MNV1_TRACE
  if (trace_path && e == cudaSuccess) {   // dump the stamps of this launch
    std::vector<unsigned long long> hbuf(4 * 128 * 4);
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf.data(), d_trace_buf, hbuf.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* fp = fopen(trace_path, "wb")) { fwrite(hbuf.data(), 8, hbuf.size(), fp); fclose(fp); }
  }
#endif
  return e;
}

}  // namespace mnv1
