// pointwise_tc.cu — 1x1 convolution on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// Replaces `pointwise` (kernel.cl:94-114; 13 launches) for bf16 contexts:
//   out[M][Cout] = act( scale * (in[M][K] . w[Cout][K]^T) + shift ),  M = N*H*W, K = Cin.
// The NHWC feature map IS the K-major A operand and the reference's [Cout][Cin] filter order
// (kernel.cl:106) IS the K-major B operand, so neither is re-laid out.
//
// Persistent, warp-specialised CTA (one per SM), 320 threads:
//   warp 0      TMA producer: A tile 128 x 64 and B tile BN x 64 (bf16, 128B swizzle) per k-block
//               into a ring of `stages` shared-memory slots (full/empty mbarriers)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32
//               accumulators in TMEM, two accumulator stages so the epilogue of tile i overlaps
//               the MMAs of tile i+1)
//   warps 2-9   epilogue (two warps per TMEM lane quarter, each taking 32 of a 64-column block):
//               tcgen05.ld 32 lanes x 32 columns -> fma(scale, shift) -> ReLU folded into the bf16
//               convert, 6-cap as min.bf16x2 -> 128B-swizzled shared-memory staging (two 128 x 64
//               buffers) -> TMA store (cp.async.bulk.tensor).  With one warp per scheduler and
//               ~10 instructions per element the epilogue, not the MMAs, set the pace
//               (profiles/r01_pw_v1_ncu.txt); it is now ~3 instructions per element on 8 warps.
// K tails (K = 32 < 64) and M tails are zero-filled on load / clipped on store by the TMA unit.
#include "common.cuh"

#include <cstdio>
#include <cstdlib>

namespace mnv1 {

namespace {

constexpr int TC_BM = 128;       // UMMA M (cta_group::1)
constexpr int TC_BK = 64;        // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_COUT = 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (SM100 UMMA SmemDescriptor):
// start address >> 4 in bits [0,14), LBO (unused for swizzled K-major, = 1) in [16,30),
// SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The issuing warp runs its loop CONVERGED (all 32 lanes wait, compute the same descriptors) and only
// the tcgen05 instructions are predicated on one elected lane.  Inside an `if (lane == 0)` branch every
// operand lives in per-thread registers and each UTCHMMA is preceded by an election loop of
// R2UR.BROADCASTs (~24 instructions, ~150 cycles): the MMAs were ISSUE-bound, 200 cycles apiece.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_bf16_if(uint32_t elected, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t elected, uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// fp32 pair -> bf16x2; ReLU rides on the convert, the upper cap is a packed min (exact: rounding is
// monotonic and the cap is representable)
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi, uint32_t cap2) {
  uint32_t d;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  asm("min.bf16x2 %0, %0, %1;" : "+r"(d) : "r"(cap2));
  return d;
}

// debug: SM-clock stamps of CTA 0 (MNV1_PW_TRACE=<file>): trace[role][idx][slot], 128 entries per role
__device__ __forceinline__ void pw_stamp(unsigned long long* tr, int role, long idx, int slot) {
  if (tr && blockIdx.x == 0 && idx < 128) tr[(role * 128 + idx) * 4 + slot] = clock64();
}

struct __align__(8) TcBarriers {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t resb_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// BN: UMMA N; RELU: activation folded into the convert; RESB: the whole filter (all k-blocks of the
// single n-tile) stays resident in shared memory for the life of the CTA and the ring carries A only.
template <int BN, bool RELU, bool RESB>
__global__ void __launch_bounds__(TC_THREADS, 1)
pointwise_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const float* __restrict__ scale,
                    const float* __restrict__ shift, uint32_t cap2, long M, int K, int Cout, int stages, int group, int dbg,
                    unsigned long long* trace) {
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
  constexpr uint32_t B_BYTES = BN * TC_BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + (RESB ? 0u : B_BYTES);
  constexpr uint32_t OUT_BYTES = TC_BM * 128;       // one 128 x 64 bf16 output block
  // two accumulator stages of 256 fp32 columns each; a stage holds `group` = 256/BN consecutive
  // m-tiles when the layer has a single n-tile, so one full/empty handshake covers 256 columns
  constexpr uint32_t ACC_COLS = 256;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] 1024-aligned, then scale/shift, then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  uint8_t* s_resb = smem + (size_t)stages * STAGE_BYTES;           // resident filter (RESB): num_kb x B tile
  uint8_t* s_out = s_resb + (RESB ? (size_t)num_kb * B_BYTES : 0);  // 2 x (128 rows x 128 B), 1024-aligned
  float* s_scale = reinterpret_cast<float*>(s_out + 2 * OUT_BYTES);
  float* s_shift = s_scale + TC_MAX_COUT;
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(s_shift + TC_MAX_COUT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const int n_tiles = Cout / BN;
  const long m_tiles = (M + TC_BM - 1) / TC_BM;
  // work unit = `group` consecutive m-tiles of one n-tile (group > 1 only when n_tiles == 1)
  const long m_groups = (m_tiles + group - 1) / group;
  const long num_units = m_groups * n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    for (int s = 0; s < stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], TC_EPI_WARPS); }
    mbar_init(&bars->resb_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();   // the previous layer's output is complete and visible

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && !(MNV1_DBG(dbg) & 2)) {
      if (RESB) {  // the filter is loaded once: every k-block of the single n-tile
        mbar_expect_tx(&bars->resb_full, (uint32_t)num_kb * B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(s_resb + (size_t)kb * B_BYTES, &tmap_b, &bars->resb_full, kb * TC_BK, 0);
      }
      int stage = 0; uint32_t phase = 0;
      for (long u = blockIdx.x; u < num_units; u += gridDim.x) {
        const long mt0 = (u / n_tiles) * group;
        const int n_idx = (int)(u % n_tiles) * BN;
        for (int g = 0; g < group && mt0 + g < m_tiles; ++g) {
          const int m_idx = (int)(mt0 + g) * TC_BM;
          for (int kb = 0; kb < num_kb; ++kb) {
            if (kb == 0 && g == 0) pw_stamp(MNV1_TRC(trace), 0, u / gridDim.x, 0);
            mbar_wait(&bars->empty[stage], phase ^ 1u);
            uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
            mbar_expect_tx(&bars->full[stage], STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &bars->full[stage], kb * TC_BK, m_idx);
            if (!RESB) tma_load_2d(sa + A_BYTES, &tmap_b, &bars->full[stage], kb * TC_BK, n_idx);
            if (kb == num_kb - 1) pw_stamp(MNV1_TRC(trace), 0, u / gridDim.x, 2);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) =================
    {
      constexpr uint32_t idesc = make_idesc(BN);
      const uint32_t elected = elect_one();
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      if (RESB && !(MNV1_DBG(dbg) & 2)) { mbar_wait(&bars->resb_full, 0); tc_fence_after(); }
      const uint32_t resb_u = smem_u32(s_resb);
      const uint32_t ring_u = smem_u32(smem);
      for (long u = blockIdx.x; u < num_units; u += gridDim.x) {
        const long mt0 = (u / n_tiles) * group;
        if (elected) pw_stamp(MNV1_TRC(trace), 1, u / gridDim.x, 0);
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1u);   // epilogue has drained this accumulator stage
        tc_fence_after();
        if (elected) pw_stamp(MNV1_TRC(trace), 1, u / gridDim.x, 1);
        for (int g = 0; g < group && mt0 + g < m_tiles; ++g) {
          const uint32_t tmem_d = tmem_base + (uint32_t)as * ACC_COLS + (uint32_t)(g * BN);
          for (int kb = 0; kb < num_kb; ++kb) {
            if (!(MNV1_DBG(dbg) & 2)) mbar_wait(&bars->full[stage], phase);          // TMA bytes have landed
            tc_fence_after();
            if (kb == 0 && g == 0 && elected) pw_stamp(MNV1_TRC(trace), 1, u / gridDim.x, 2);
            const uint32_t sa = ring_u + (uint32_t)stage * STAGE_BYTES;
            const uint64_t da = make_smem_desc(sa);
            const uint64_t db = make_smem_desc(RESB ? resb_u + (uint32_t)kb * B_BYTES : sa + A_BYTES);
            const int krem = K - kb * TC_BK;                // K tail: skip the zero-filled 16-wide slices
            const int nk = krem >= TC_BK ? TC_BK / TC_UMMA_K : (krem + TC_UMMA_K - 1) / TC_UMMA_K;
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
              // advance 16 elements = 32 bytes inside the 128B swizzle row: +2 in the (>>4) address field
              if (k < nk) umma_bf16_if(elected, tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit_if(elected, &bars->empty[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit_if(elected, &bars->tmem_full[as]);     // accumulator stage complete -> epilogue
        if (elected) pw_stamp(MNV1_TRC(trace), 1, u / gridDim.x, 3);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ================= epilogue warps 2..9 =================
    // The epilogue warps stage scale / shift themselves (every 128-bit load issued before the first
    // store: one memory latency) while the producer and the MMA issuer are already at work; a serial
    // copy loop in front of the CTA-wide barrier cost 2-4 us of these 25-40 us kernels.
    {
      const int et = threadIdx.x - 64, n4 = Cout >> 2;   // Cout % 64 == 0, <= 1024: at most one float4 per thread and array
      float4 sv = make_float4(1.f, 1.f, 1.f, 1.f), tv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (et < n4) {
        if (scale) sv = __ldg(reinterpret_cast<const float4*>(scale) + et);
        if (shift) tv = __ldg(reinterpret_cast<const float4*>(shift) + et);
        reinterpret_cast<float4*>(s_scale)[et] = sv;
        reinterpret_cast<float4*>(s_shift)[et] = tv;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
    }
    const int quarter = warp & 3;                        // TMEM lanes 32*quarter .. +31 belong to this warp
    const int half = (warp - 2) >> 2;                    // which 32 columns of each 64-column block
    const int row = quarter * 32 + lane;                 // row of the tile this thread owns
    const bool leader = threadIdx.x == 64;               // issues the TMA stores
    const uint32_t s_out_u = smem_u32(s_out);
    const uint32_t s_scale_u = smem_u32(s_scale), s_shift_u = smem_u32(s_shift);
    const uint32_t row_off = (uint32_t)row * 128u, row_x = (uint32_t)(row & 7);
    int as = 0; uint32_t aphase = 0;
    uint32_t blk = 0;                                    // running 64-column block counter -> staging buffer
    for (long u = blockIdx.x; u < num_units; u += gridDim.x) {
      const long mt0 = (u / n_tiles) * group;
      const int n_idx = (int)(u % n_tiles) * BN;
      if (threadIdx.x == 64) pw_stamp(MNV1_TRC(trace), 2, u / gridDim.x, 0);
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      if (threadIdx.x == 64) pw_stamp(MNV1_TRC(trace), 2, u / gridDim.x, 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)as * ACC_COLS + (uint32_t)(32 * half);
      const int nblk = group * (BN / 64);                // 64-column blocks in this accumulator stage
#pragma unroll 1
      for (int b = 0; b < nblk; ++b) {
        const int g = b / (BN / 64), c0 = (b % (BN / 64)) * 64;
        if (mt0 + g >= m_tiles) break;                   // uniform: the last group may be short
        const int m_idx = (int)(mt0 + g) * TC_BM;
        const uint32_t sbuf = s_out_u + (blk & 1u) * OUT_BYTES;
        ++blk;
        // the store issued two blocks ago read this buffer: wait until it has been consumed
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(b * 64), v);
        const uint32_t col = (uint32_t)(n_idx + c0 + 32 * half);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 s0 = lds128f(s_scale_u + (col + j) * 4u), s1 = lds128f(s_scale_u + (col + j + 4) * 4u);
          const float4 t0 = lds128f(s_shift_u + (col + j) * 4u), t1 = lds128f(s_shift_u + (col + j + 4) * 4u);
          const uint32_t p0 = pack2<RELU>(fmaf(__uint_as_float(v[j + 0]), s0.x, t0.x), fmaf(__uint_as_float(v[j + 1]), s0.y, t0.y), cap2);
          const uint32_t p1 = pack2<RELU>(fmaf(__uint_as_float(v[j + 2]), s0.z, t0.z), fmaf(__uint_as_float(v[j + 3]), s0.w, t0.w), cap2);
          const uint32_t p2 = pack2<RELU>(fmaf(__uint_as_float(v[j + 4]), s1.x, t1.x), fmaf(__uint_as_float(v[j + 5]), s1.y, t1.y), cap2);
          const uint32_t p3 = pack2<RELU>(fmaf(__uint_as_float(v[j + 6]), s1.z, t1.z), fmaf(__uint_as_float(v[j + 7]), s1.w, t1.w), cap2);
          // 16-byte chunk index inside the 128-byte row, XOR-swizzled with the row (SWIZZLE_128B)
          const uint32_t chunk = (uint32_t)(4 * half + j / 8);
          sts128(sbuf + row_off + ((chunk ^ row_x) << 4), p0, p1, p2, p3);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
#ifndef PW_EXP_NOSTORE
        if (leader) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_out),
                       "r"(sbuf), "r"(n_idx + c0), "r"(m_idx)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#endif
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
      if (threadIdx.x == 64) pw_stamp(MNV1_TRC(trace), 2, u / gridDim.x, 2);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = "cuTensorMapEncodeTiled entry point not available";
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D bf16 tensor [rows][k] (k contiguous) with a (64 x box_rows) box and 128B swizzle
// (used for A = activations, B = filter and the output tile alike).
cudaError_t encode_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t k, uint32_t box_rows,
                      std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t gdim[2] = {k, rows};
  cuuint64_t gstride[1] = {k * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char b[160];
      snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu k=%llu box_rows=%u", (int)r,
               (unsigned long long)rows, (unsigned long long)k, box_rows);
      *err = b;
    }
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

int pick_bn(int cout) {
  if (cout % 256 == 0) return 256;
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  return 0;
}

template <int BN>
cudaError_t launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const float* scale,
                      const float* shift, int act, long m, int k, int cout, int num_sms, cudaStream_t st) {
  const int num_kb = (k + TC_BK - 1) / TC_BK;
  const int n_tiles = cout / BN;
  const size_t b_bytes = (size_t)BN * TC_BK * 2;
  // resident filter: single n-tile and the whole [BN][K] filter fits in 64 KB
  const bool resb = n_tiles == 1 && (size_t)num_kb * b_bytes <= 64 * 1024;
  const int group = n_tiles == 1 ? 256 / BN : 1;
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (resb ? 0 : b_bytes);
  const size_t fixed = 1024 /*align slack*/ + (resb ? (size_t)num_kb * b_bytes : 0) + 2 * TC_BM * 128 /*output staging*/ +
                       2 * TC_MAX_COUT * sizeof(float) + sizeof(TcBarriers);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  // a ring deeper than ~2 tiles' worth of k-blocks only costs smem; keep at least 4
  int want = 2 * num_kb < 4 ? 4 : 2 * num_kb;
  if (resb && want < 8) want = 8;
  if (stages > want) stages = want;
  if (stages < 2) return cudaErrorInvalidValue;
  const size_t smem = fixed + stages * stage_bytes;
  {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* fn) {
      if (e == cudaSuccess) e = ensure_dyn_smem((const void*)fn, 227 * 1024);
    };
    set((const void*)pointwise_tc_kernel<BN, true, true>);
    set((const void*)pointwise_tc_kernel<BN, true, false>);
    set((const void*)pointwise_tc_kernel<BN, false, true>);
    set((const void*)pointwise_tc_kernel<BN, false, false>);
    if (e != cudaSuccess) return e;
  }
  const long m_tiles = (m + TC_BM - 1) / TC_BM;
  const long units = ((m_tiles + group - 1) / group) * n_tiles;
  const unsigned grid = (unsigned)(units < num_sms ? units : num_sms);
  const uint32_t cap2 = act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;  // bf16x2 (6, 6) or (+inf, +inf)
  const bool relu = act != MNV1_ACT_NONE;
  int dbg = 0;
  unsigned long long* d_trace = nullptr;
#ifdef MNV1_TRACE   // libmnv1_trace.so only: MNV1_PW_DBG=2 runs without loads, MNV1_PW_TRACE=<file> dumps CTA 0's stamps
  { static const int dbg_env = getenv("MNV1_PW_DBG") ? atoi(getenv("MNV1_PW_DBG")) : 0; dbg = dbg_env; }
  static unsigned long long* d_trace_buf = nullptr;
  const char* trace_path = capturing(st) ? nullptr : getenv("MNV1_PW_TRACE");
  if (trace_path) {
    if (!d_trace_buf) cudaMalloc(&d_trace_buf, 4 * 128 * 4 * 8);
    cudaMemsetAsync(d_trace_buf, 0, 4 * 128 * 4 * 8, st);
    d_trace = d_trace_buf;
  }
#endif
  cudaError_t le;
#define PW_LAUNCH(R, B) \
  le = launch_pdl(pointwise_tc_kernel<BN, R, B>, dim3(grid), dim3(TC_THREADS), smem, st, ta, tb, to, scale, shift, cap2, m, k, cout, stages, group, dbg, d_trace)
  if (relu) { if (resb) PW_LAUNCH(true, true); else PW_LAUNCH(true, false); }
  else      { if (resb) PW_LAUNCH(false, true); else PW_LAUNCH(false, false); }
#undef PW_LAUNCH
#ifdef MNV1_TRACE
  if (trace_path && le == cudaSuccess) {
    std::vector<unsigned long long> hbuf(4 * 128 * 4);
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf.data(), d_trace_buf, hbuf.size() * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "wb")) { fwrite(hbuf.data(), 8, hbuf.size(), f); fclose(f); }
  }
#endif
  return le;
}

}  // namespace

cudaError_t make_weight_tmap(mnv1_filter* f, std::string* err) {
  const int bn = pick_bn(f->cout);
  if (!bn || f->cin % 8 || f->cout > TC_MAX_COUT) {
    if (err) *err = "pointwise tcgen05 path needs Cout % 64 == 0, Cout <= 1024 and Cin % 8 == 0";
    return cudaErrorInvalidValue;
  }
  cudaError_t e = encode_2d(&f->tmap_b, f->w_bf16, (uint64_t)f->cout, (uint64_t)f->cin, (uint32_t)bn, err);
  if (e == cudaSuccess) { f->has_tmap = true; f->tmap_bn = bn; }
  return e;
}

cudaError_t launch_pointwise_tc(bf16* out, const bf16* in, const mnv1_filter* f, long m, int k, int cout,
                                int num_sms, cudaStream_t st, std::string* err) {
  if (!f->has_tmap || k != f->cin || cout != f->cout) {
    if (err) *err = "pointwise_tc: filter has no TMA descriptor or shape mismatch";
    return cudaErrorInvalidValue;
  }
  if (m <= 0) return cudaSuccess;
  CUtensorMap ta, to;
  cudaError_t e = encode_2d(&ta, in, (uint64_t)m, (uint64_t)k, TC_BM, err);
  if (e != cudaSuccess) return e;
  e = encode_2d(&to, out, (uint64_t)m, (uint64_t)cout, TC_BM, err);
  if (e != cudaSuccess) return e;
  switch (f->tmap_bn) {
    case 256: return launch_bn<256>(ta, f->tmap_b, to, f->scale, f->shift, f->act, m, k, cout, num_sms, st);
    case 128: return launch_bn<128>(ta, f->tmap_b, to, f->scale, f->shift, f->act, m, k, cout, num_sms, st);
    case 64:  return launch_bn<64>(ta, f->tmap_b, to, f->scale, f->shift, f->act, m, k, cout, num_sms, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace mnv1
