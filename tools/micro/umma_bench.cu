// Microbenchmark: issue rate of tcgen05.mma (kind::f16, cta_group::1, M=128) from one thread, as a function of
// N, swizzle mode (K block 64/32/16 -> 128/64/32-byte rows), number of independent accumulators and A start offset.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo, uint32_t layout) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | ((layout & 7u) << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t a) { return ((a >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
__global__ void bench(int N, int bk, int nacc, int iters, int a_off_bytes, long long* out, int uniform) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (uniform ? (warp == 1) : (threadIdx.x == 32)) {
    uint32_t leader = 1;
    if (uniform) asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(leader));
    const uint32_t layout = bk == 64 ? 2u : bk == 32 ? 4u : 6u;
    const uint32_t hi = desc_hi(8u * bk * 2u, layout);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_lo = desc_lo(smem_u32(smem) + a_off_bytes), b_lo = desc_lo(smem_u32(smem + 48 * 1024));
    const int mma_per_kb = bk / 16;
    long long t0 = clock64();
    if (uniform == 2) {
      if (leader) {
        const uint32_t d0 = tmem, d1 = tmem + N, d2 = tmem + 2 * N, d3 = tmem + 3 * N, d4 = tmem + 4 * N, d5 = tmem + 5 * N, d6 = tmem + 6 * N, d7 = tmem + 7 * N;
        for (int it = 0; it < iters * mma_per_kb; ++it) {
          umma(d0, a_lo, b_lo, hi, idesc, 1); umma(d1, a_lo, b_lo, hi, idesc, 1); umma(d2, a_lo, b_lo, hi, idesc, 1); umma(d3, a_lo, b_lo, hi, idesc, 1);
          if (nacc == 8) { umma(d4, a_lo, b_lo, hi, idesc, 1); umma(d5, a_lo, b_lo, hi, idesc, 1); umma(d6, a_lo, b_lo, hi, idesc, 1); umma(d7, a_lo, b_lo, hi, idesc, 1); }
        }
      }
    } else
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < mma_per_kb; ++k)
        for (int a = 0; a < nacc; ++a) { if (leader) umma(tmem + a * N, a_lo + 2 * k, b_lo + 2 * k, hi, idesc, 1); }
      if (uniform) __syncwarp();
    }
    long long t1 = clock64();   // issue time only
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t2 = clock64();   // all MMAs retired
    if (leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 256;
  printf("%4s %3s %4s %6s | %10s %10s (cycles per MMA: issue-only, issue+retire)\n", "N", "bk", "nacc", "a_off", "issue", "total");
  for (int uniform : {2})
  for (int bk : {64})
    for (int N : {32, 64, 128})
      for (int nacc : {4, 8})
        for (int off : {0}) {
          if (nacc * N > 512) continue;
          if (off && bk == 64 && 0) continue;
          bench<<<1, 64, 100 * 1024>>>(N, bk, nacc, iters, off * (bk == 16 ? 1 : 2), d, uniform);   // 64 / 128 byte shifts
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          cudaError_t e = cudaGetLastError();
          const double n = (double)iters * (bk / 16) * nacc;
          printf("u%d %4d %3d %4d %6d | %10.1f %10.1f %s\n", uniform, N, bk, nacc, off * (bk == 16 ? 1 : 2), h[0] / n, h[1] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
