// microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128, N in {64,128,256}) issued back to back by one
// elected thread, operands = garbage in smem (128B swizzle descriptors), one CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__global__ void __launch_bounds__(128, 1) k(int n, int iters, int same, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t smem = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
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
    for (int i = 0; i < iters; ++i) {
      const uint32_t st = same ? 0 : (uint32_t)(i & 3);           // rotate over 4 stages of 48 KB like the real ring
      const uint64_t da = desc(smem + st * 49152u), db = desc(smem + st * 49152u + 16384u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tm + (uint32_t)((i & 1) * 256)), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D:}" ::"r"(smem_u32(&bar)) : "memory");
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (blockIdx.x == 0) { out[0] = (unsigned long long)(c1 - c0); out[1] = g1 - g0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 16); unsigned long long h[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148}) for (int same : {1, 0}) for (int n : {64, 128, 256}) {
    const int iters = 2000;
    k<<<grid, 128, 200 * 1024>>>(n, iters, same, d); cudaDeviceSynchronize();
    k<<<grid, 128, 200 * 1024>>>(n, iters, same, d); cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("grid %3d same_stage %d N=%3d: %.1f cycles / MMA (128xNx16), %.2f GHz, %.0f TFLOP/s per chip-equivalent  [%s]\n", grid, same, n,
           (double)h[0] / (iters * 4), (double)h[0] / h[1], 148.0 * 2.0 * 128 * n * 16 * iters * 4 / (h[1] * 1e-9) / 1e12, cudaGetErrorString(e));
  }
  return 0;
}
