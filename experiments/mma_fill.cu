// microbenchmark: does filling shared memory (bulk copies from L2, as a TMA operand ring does) slow the
// tcgen05.mma stream that reads its operands from shared memory?  One thread issues 128x256x16 bf16 UMMAs back
// to back on a fixed 48 KB operand stage; a second warp keeps Q bulk copies of `chunk` bytes in flight into a
// separate region of the same shared memory.  Reports cycles per UMMA and the achieved fill rate in B/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
  asm volatile("{.reg .pred p; W2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D2; bra W2; D2:}" ::"r"(bar), "r"(par) : "memory");
}
constexpr int QMAX = 8, FMAX = 3;
__global__ void __launch_bounds__(128, 1) k(int n, int iters, int q, int chunk, int nmma, int nf, const uint8_t* src, size_t src_bytes,
                                            unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t smem = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, fbar_all[FMAX][QMAX]; __shared__ uint32_t slot; __shared__ volatile int done;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int f = 0; f < FMAX; ++f) for (int i = 0; i < QMAX; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fbar_all[f][i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    done = 0;
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    unsigned long long g0, g1; long long c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const uint64_t da = desc(smem), db = desc(smem + 16384u);
    if (nmma)
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tm + (uint32_t)((i & 1) * 256)), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(idesc), "r"(1u) : "memory");
      if ((i & 15) == 15) {   // keep the issue queue bounded like a real pipeline: wait for the batch
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait(smem_u32(&bar), (uint32_t)(i >> 4) & 1u);
      }
    }
    else { while (clock64() - c0 < 400000) {} }
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    done = 1;
    if (blockIdx.x == 0) { out[0] = (unsigned long long)(c1 - c0); out[1] = g1 - g0; }
  } else if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) <= nf && q > 0) {
    const int f = (threadIdx.x >> 5) - 1;
    uint64_t* fbar = fbar_all[f];
    // filler: q copies in flight, each `chunk` bytes, round-robin over q slots behind the operand stage
    const uint32_t fill0 = smem + 49152u + (uint32_t)f * (uint32_t)(q * chunk);
    unsigned long long bytes = 0; long long c0 = clock64();
    size_t off = (((size_t)blockIdx.x * 3 + f) * 7919u * 16384u) % (src_bytes - (size_t)chunk);
    off &= ~(size_t)1023;
    uint32_t ph[QMAX] = {0};
    for (int s = 0; s < q; ++s) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fbar[s])), "r"((uint32_t)chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fill0 + s * chunk), "l"(src + off), "r"((uint32_t)chunk), "r"(smem_u32(&fbar[s])) : "memory");
      off += chunk; if (off + chunk > src_bytes) off = 0;
    }
    int s = 0;
    while (!done) {
      wait(smem_u32(&fbar[s]), ph[s]); ph[s] ^= 1; bytes += chunk;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fbar[s])), "r"((uint32_t)chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fill0 + s * chunk), "l"(src + off), "r"((uint32_t)chunk), "r"(smem_u32(&fbar[s])) : "memory");
      off += chunk; if (off + chunk > src_bytes) off = 0;
      s = (s + 1 == q) ? 0 : s + 1;
    }
    long long c1 = clock64();
    for (int t = 0; t < q; ++t) { wait(smem_u32(&fbar[s]), ph[s]); s = (s + 1 == q) ? 0 : s + 1; }   // drain
    if (blockIdx.x == 0 && f == 0) { out[2] = bytes; out[3] = (unsigned long long)(c1 - c0); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 32); unsigned long long h[4];
  const size_t src_bytes = 48u << 20;   // L2-resident source
  uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  for (int grid : {148}) for (int nf : {1, 2, 3}) for (int q : {1, 2, 3}) for (int chunk : {8192, 16384, 32768, 49152}) {
    const int n = 256, nmma = 1;
    if (nf * q * chunk > 147456) continue;
    cudaMemset(d, 0, 32);
    k<<<grid, 128, 200 * 1024>>>(n, iters, q, chunk, nmma, nf, src, src_bytes, d); cudaDeviceSynchronize();
    k<<<grid, 128, 200 * 1024>>>(n, iters, q, chunk, nmma, nf, src, src_bytes, d); cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("grid %3d fillers %d q=%d x %5d B: %6.1f cycles / UMMA, fill %6.1f B/clk/SM per filler (%6.1f total), %5.0f cycles per copy [%s]\n", grid, nf, q, chunk,
           (double)h[0] / (iters * 4), h[3] ? (double)h[2] / h[3] : 0.0, h[3] ? nf * (double)h[2] / h[3] : 0.0, h[2] ? (double)h[3] * chunk / h[2] : 0.0, cudaGetErrorString(e));
  }
  return 0;
}
