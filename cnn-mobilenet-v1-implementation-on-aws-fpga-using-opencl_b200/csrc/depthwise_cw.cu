// depthwise_cw.cu — depthwise 3x3 for the 7x7x1024 map (layer 26): one thread owns ONE channel
// pair of one image and walks down the map.
//
// Contract: `depthwise` (kernel.cl:62-92, intended semantics: per-channel 3x3, stride 1 or 2, zero
// padding, folded scale/shift, ReLU / ReLU6), NHWC bf16 in and out, fp32 accumulation.
//
// Why not tiles + halos here: a 7x7 map read through 9x9 TMA boxes moves 1.65x the bytes, the box
// rows carry 14 bytes of payload per channel chunk, and the shared-memory hand-shakes cost more
// than the arithmetic (depthwise_tma.cu: 22.7 us for 51 MB).  With a thread per channel pair every
// input element is read from global memory exactly once, adjacent threads read adjacent channel
// pairs (128 B per warp per pixel, full sectors), no thread waits for another, and the kernel is a
// stream of independent 4-byte copies kept D output rows ahead of the arithmetic in a
// thread-private shared-memory ring of 3 + S*D input rows (cp.async groups, one per output row).
//   per input element: 1 async copy, 1-3 shared loads, 2 ALU ops to widen the bf16 pair to an
//   f32x2, up to 3 FFMA2 per filter row; per output: convert + clamp + 1 store.
//   Algorithmic bytes = in + out, no halo.  15.5 us at batch 256 (3.3 TB/s).
// The template also covers bands of rows, column splits and stride 2 (14x14 maps); those
// configurations were measured and lost to depthwise_ring.cu, see the dispatch at the bottom.
#include "common.cuh"
#include "sm100.cuh"

namespace mnv1 {
namespace {

using namespace ptx;

// 4-byte asynchronous copy global -> shared: the prefetch that the register
// allocator cannot undo.  (A first version kept the row ring in registers; ptxas moved every load
// back next to its first use to save registers and the kernel ran at the latency of one row.)
__device__ __forceinline__ void cp_async4(uint32_t dst, const uint32_t* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// n is a constant after unrolling; the switch folds to one instruction
__device__ __forceinline__ void cp_async_wait_n(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

template <int S_, int HI_, int WI_, int PAD_, int C_, int BANDS_, int CS_, int D_, int THREADS_, int MINB_>
struct CwCfg {
  static constexpr int S = S_, HI = HI_, WI = WI_, PAD = PAD_, C = C_, BANDS = BANDS_, CS = CS_, D = D_;
  static constexpr int THREADS = THREADS_, MINB = MINB_, CP = C / 2;
  static constexpr int HO = HI / S, WO = WI / S, BR = HO / BANDS, NR = 3 + S * D;
  static constexpr int WOC = WO / CS;               // output columns per thread
  static constexpr int WIC = (WOC - 1) * S + 3;     // input columns per thread (window incl. halo / padding)
  static constexpr size_t SMEM = (size_t)NR * WIC * THREADS * 4;
  static_assert(HO % BANDS == 0 && WO % CS == 0, "bands / column splits must divide the output map");
  static_assert(CP % THREADS == 0 && THREADS % 32 == 0, "a block covers whole warps of one pixel's channel pairs");
};

// Work item = (image, band of output rows, column split, channel pair), channel pair fastest.
// Everything but the item decode is compile-time: every load, ring slot and store is
// [thread base + immediate].
template <typename Cfg, bool RELU>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
depthwise_cw_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, const float* __restrict__ w9xC,
                    const float* __restrict__ shift, uint32_t cap2, int items) {
  constexpr int S = Cfg::S, HI = Cfg::HI, WI = Cfg::WI, PAD = Cfg::PAD, BANDS = Cfg::BANDS, CS = Cfg::CS, D = Cfg::D;
  constexpr int WO = Cfg::WO, HO = Cfg::HO, BR = Cfg::BR, NR = Cfg::NR, WOC = Cfg::WOC, WIC = Cfg::WIC;
  constexpr int CP = Cfg::CP, THREADS = Cfg::THREADS;
  pdl_trigger();
  const int item = (int)blockIdx.x * THREADS + (int)threadIdx.x;
  if (item >= items) return;
  const int cp = item % CP;
  const int rest = item / CP;
  const int cs = rest % CS, band = (rest / CS) % BANDS, img = rest / (CS * BANDS);

  f32x2 w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(w9xC) + t * CP + cp);
    w[t] = f2_pack(v.x, v.y);
  }
  f32x2 sh = f2_pack(0.f, 0.f);
  if (shift) { const float2 v = __ldg(reinterpret_cast<const float2*>(shift) + cp); sh = f2_pack(v.x, v.y); }
  pdl_wait();   // the previous layer's output is complete

  const int iy_first = band * BR * S - PAD;          // input row of window row 0
  const int ix_first = cs * WOC * S - PAD;           // input column of window column 0
  // window origin; may point outside the image (only dereferenced under the row / column predicates)
  const uint32_t* src = in + ((long)img * HI * WI + (long)iy_first * WI + ix_first) * CP + cp;
  uint32_t* dst = out + (((long)img * HO + band * BR) * WO + cs * WOC) * CP + cp;

  // thread-private ring of NR input rows in shared memory, [row][column][thread]
  extern __shared__ uint32_t cw_ring[];
  const uint32_t ring = smem_u32(cw_ring) + threadIdx.x * 4;
  auto slot = [&](int jr, int lx) { return ring + (uint32_t)((jr * WIC + lx) * THREADS * 4); };
  // rows / columns outside the image: static where the split leaves no choice, a predicate otherwise
  auto issue_row = [&](int j) {
    bool row_ok = true;
    if (BANDS == 1) { if (j - PAD < 0 || j - PAD >= HI) row_ok = false; }
    else row_ok = iy_first + j >= 0 && iy_first + j < HI;
#pragma unroll
    for (int lx = 0; lx < WIC; ++lx) {
      bool ok = row_ok;
      if (CS == 1) { if (lx - PAD < 0 || lx - PAD >= WI) continue; }   // padding column: never read, never multiplied
      else ok = ok && ix_first + lx >= 0 && ix_first + lx < WI;
      if (ok) cp_async4(slot(j % NR, lx), src + (j * WI + lx) * CP);
      else asm volatile("st.shared.u32 [%0], %1;" ::"r"(slot(j % NR, lx)), "r"(0u) : "memory");
    }
  };
  // one copy group per output row: the rows first needed by output row r (r = 0: rows 0..2, then S more)
  auto issue_for = [&](int r) {
    if (r == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) issue_row(j);
    } else {
#pragma unroll
      for (int j = r * S + 3 - S; j <= r * S + 2; ++j) issue_row(j);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int r = 0; r < D && r < BR; ++r) issue_for(r);
#pragma unroll
  for (int r = 0; r < BR; ++r) {
    if (r + D < BR) issue_for(r + D);
    // groups 0..r must have landed; at most min(D, BR-1-r) younger ones stay in flight
    cp_async_wait_n(BR - 1 - r < D ? BR - 1 - r : D);
    f32x2 acc[WOC];
#pragma unroll
    for (int ox = 0; ox < WOC; ++ox) acc[ox] = sh;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int lx = 0; lx < WIC; ++lx) {
        if (CS == 1 && (lx - PAD < 0 || lx - PAD >= WI)) continue;
        // window column lx feeds output ox through tap dx = lx - S*ox when 0 <= dx <= 2
        bool used = false;
#pragma unroll
        for (int ox = 0; ox < WOC; ++ox) { const int dx = lx - S * ox; used |= dx >= 0 && dx <= 2; }
        if (!used) continue;
        const f32x2 v = f2_from_bf16x2(lds32(slot((r * S + dy) % NR, lx)));
#pragma unroll
        for (int ox = 0; ox < WOC; ++ox) {
          const int dx = lx - S * ox;
          if (dx >= 0 && dx <= 2) acc[ox] = f2_fma(v, w[dy * 3 + dx], acc[ox]);
        }
      }
    }
#pragma unroll
    for (int ox = 0; ox < WOC; ++ox) dst[(r * WO + ox) * CP] = pack2_f2<RELU>(acc[ox], cap2);
  }
}

template <typename Cfg>
cudaError_t launch_cw(bf16* out, const bf16* in, const float* w, const float* shift, int act, int n, cudaStream_t st) {
  const long items = (long)n * Cfg::BANDS * Cfg::CS * Cfg::CP;
  if (items >= (1L << 31) || (long)n * Cfg::HI * Cfg::WI * Cfg::CP >= (1L << 31)) return cudaErrorNotSupported;
  const long grid = (items + Cfg::THREADS - 1) / Cfg::THREADS;
  const uint32_t cap2 = act == MNV1_ACT_RELU6 ? 0x40c040c0u : 0x7f807f80u;
  {
    cudaError_t e = ensure_dyn_smem((const void*)depthwise_cw_kernel<Cfg, true>, (int)Cfg::SMEM);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)depthwise_cw_kernel<Cfg, false>, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
  }
  if (act != MNV1_ACT_NONE)
    return launch_pdl(depthwise_cw_kernel<Cfg, true>, dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::SMEM, st,
                      reinterpret_cast<uint32_t*>(out), reinterpret_cast<const uint32_t*>(in), w, shift, cap2, (int)items);
  return launch_pdl(depthwise_cw_kernel<Cfg, false>, dim3((unsigned)grid), dim3(Cfg::THREADS), Cfg::SMEM, st,
                    reinterpret_cast<uint32_t*>(out), reinterpret_cast<const uint32_t*>(in), w, shift, cap2, (int)items);
}

}  // namespace

// Thread-per-channel-pair depthwise for the 7x7x1024 map (bf16).  cudaErrorNotSupported = no variant.
cudaError_t launch_depthwise_cw(bf16* out, const bf16* in, const float* w9xC_scaled, const float* shift, int act, int n,
                                int rows, int cols, int stride, int c, int pad_lo, cudaStream_t st) {
  if (rows != cols) return cudaErrorNotSupported;
  if (n <= 0) return cudaSuccess;
#define CW_GO(...) return launch_cw<CwCfg<__VA_ARGS__>>(out, in, w9xC_scaled, shift, act, n, st)
  //                                                                  S HI WI PAD   C  BANDS CS D THREADS MINB
  if (stride == 1 && rows == 7 && pad_lo == 1 && c == 1024) CW_GO(1, 7, 7, 1, 1024, 1, 1, 2, 128, 8);
  // 14x14x512 (stride 1: two bands, D = 2; stride 2: one band) measured 27.6 / 16.8 us against the
  // 24.5 / 14.5 us of depthwise_ring.cu, so those shapes stay there (experiments/README.md)
#undef CW_GO
  return cudaErrorNotSupported;
}

}  // namespace mnv1
