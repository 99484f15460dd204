// Microbenchmark: issue rate of tcgen05.mma (kind::f16, SS operands) on one SM, to read the classifier's kernels against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/micro/mma_rate.cu && /tmp/mma_rate
// One CTA per SM; one thread issues `iters` MMAs (M = 128 or 256 with cta_group::2, N, K = 16) over operand tiles that already sit in
// shared memory (contents irrelevant), into `nacc` accumulators round-robin, then commits and waits; clock64 around it.
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (long it = 0; !done && it < (1L << 26); ++it)
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (!done) __trap();
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout_type & 7u) << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred != 0;
}

template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}

struct Args { int n, iters, nacc, layout_type, row_bytes, kslices, m64, elect, commit_every; };

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(Args a, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 16384;         // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0;
  const int M = CG == 2 ? 256 : (a.m64 ? 64 : 128);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(a.n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t hi = desc_hi(8u * a.row_bytes, a.layout_type);
  const uint32_t a_lo = desc_lo(smem_u32(sA)), b_lo = desc_lo(smem_u32(sB));
  uint64_t* scratch = bar + 1;     // second barrier: the per-k-block commits of the `commit_every` variant land here
  if (threadIdx.x == 0) mbar_init(scratch, (1 << 20) - 1);
  __syncthreads();
  if (a.elect == 2) {
    // burst: ONE elected lane issues all the MMAs back to back (no per-instruction warp synchronisation): the hardware's own interval
    if (threadIdx.x < 32 && rank == 0) {
      t0 = clock64();
      if (elect_one()) {
        int acc_i = 0, ks = 0;
        for (int i = 0; i < a.iters; ++i) {
          umma<CG>(tmem + (uint32_t)(acc_i * a.n), a_lo + 2 * ks, b_lo + 2 * ks, hi, idesc, 1u);
          if (++acc_i == a.nacc) acc_i = 0;
          if (++ks == a.kslices) {
            ks = 0;
            if (a.commit_every) {
              if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(scratch)) : "memory");
              else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(scratch)) : "memory");
            }
          }
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(bar, 0);
      t1 = clock64();
      if (blockIdx.x < 2 && threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
  } else if (a.elect) {
    // the form the classifier uses now: the whole warp walks the loop on uniform values, one elected lane issues
    if (threadIdx.x < 32 && rank == 0) {
      t0 = clock64();
      int acc_i = 0, ks = 0;
      for (int i = 0; i < a.iters; ++i) {
        if (elect_one()) {
          umma<CG>(tmem + (uint32_t)(acc_i * a.n), a_lo + 2 * ks, b_lo + 2 * ks, hi, idesc, 1u);
          if (a.commit_every && ks == a.kslices - 1) {
            if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(scratch)) : "memory");
            else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(scratch)) : "memory");
          }
        }
        __syncwarp();
        if (++acc_i == a.nacc) acc_i = 0;
        if (++ks == a.kslices) ks = 0;
      }
      if (elect_one()) {
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(bar, 0);
      t1 = clock64();
      if (blockIdx.x < 2 && threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
  } else if (threadIdx.x == 0 && rank == 0) {
    // the old form: one thread inside `if (lane == 0)` - ptxas wraps every instruction in an ELECT / R2UR / BRA.U.ANY loop
    t0 = clock64();
    int acc_i = 0, ks = 0;
    for (int i = 0; i < a.iters; ++i) {
      umma<CG>(tmem + (uint32_t)(acc_i * a.n), a_lo + 2 * ks, b_lo + 2 * ks, hi, idesc, 1u);
      if (++acc_i == a.nacc) acc_i = 0;
      if (++ks == a.kslices) ks = 0;
    }
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    mbar_wait(bar, 0);
    t1 = clock64();
    if (blockIdx.x < 2) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (threadIdx.x < 32) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  const int smem = 1024 + 16384 + 32768 + 64;
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("issue cg  M   N  nacc swz kslices commit/kblock  clk/MMA  floor(N/2*M/128/cg)  MAC/clk/SM\n");
  const int iters = 4096;
  for (int el = 2; el >= 0; --el)
  for (int ce = 0; ce <= (el ? 1 : 0); ++ce)
  for (int cg = 1; cg <= 2; ++cg)
    for (int m64 = 0; m64 <= (cg == 1 && el == 2 && ce == 0 ? 1 : 0); ++m64)
      for (int n : {32, 64, 96, 128, 192, 256})
        for (int nacc : {1, 2, 4})
          for (int lt : {2, 4}) {       // 2 = 128B swizzle (64-wide K block), 4 = 64B swizzle (32-wide)
            if (nacc * n > 512) continue;
            if (m64 && (n % 8)) continue;
            if (lt == 4 && (el != 2 || ce == 1 || nacc == 4)) continue;
            if (el != 2 && nacc != 1) continue;
            Args a{n, iters, nacc, lt, lt == 2 ? 128 : 64, lt == 2 ? 4 : 2, m64, el, ce};
            long long h[2] = {0, 0};
            cudaMemset(d_out, 0, 16);
            if (cg == 1) rate_kernel<1><<<148, 128, smem>>>(a, d_out);
            else {
              cudaLaunchConfig_t cfg = {};
              cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
              cudaLaunchAttribute at[1];
              at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
              cfg.attrs = at; cfg.numAttrs = 1;
              cudaLaunchKernelEx(&cfg, rate_kernel<2>, a, d_out);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cg %d n %d: %s\n", cg, n, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
            const int M = cg == 2 ? 256 : (m64 ? 64 : 128);
            const double clk = (double)h[0] / iters;
            printf("%s %d  %3d %3d  %d    %s  %d   %d     %7.1f   %6.1f   %8.0f\n", el == 2 ? "burst" : el ? "elect" : "lane0", cg, M, n, nacc, lt == 2 ? "128B" : " 64B", a.kslices, ce, clk,
                   n / 2.0 * (M / 128.0) / cg * (m64 ? 2 : 1), (double)M * n * 16 / clk / cg);
          }
  return 0;
}
