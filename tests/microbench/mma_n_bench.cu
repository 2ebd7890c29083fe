// tcgen05.mma issue-rate micro-benchmark (sm_100a): clocks per MMA (M = 128, cta_group::1, SS operands in 128B-swizzled
// shared memory) as a function of N, for kind::f16 (bf16, K = 16) and kind::i8 (K = 32).  Operand contents are
// irrelevant (zeros).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_n_bench mma_n_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool I8>
__global__ void __launch_bounds__(128) bench(int n, int iters, long long* clocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = I8 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24))
                              : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24));
    const uint64_t adesc = make_desc(smem_u32(smem)), bdesc = make_desc(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (I8)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + (uint32_t)((it & 1) * 256)), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + (uint32_t)((it & 1) * 256)), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u), "r"(0x989680u) : "memory");
    const long long t1 = clock64();
    clocks[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool I8>
void run(int n, long long* dclk) {
  const int iters = 2000, grid = 148;
  const size_t smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(bench<I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<I8><<<grid, 128, smem>>>(n, iters, dclk);
  bench<I8><<<grid, 128, smem>>>(n, iters, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, dclk, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double per = mean / (iters * 4.0);
  const double macs = 128.0 * n * (I8 ? 32 : 16);
  printf("kind::%s M=128 N=%3d K=%2d : %7.1f clk per MMA, %6.0f MAC/clk/SM (ideal %d: %5.1f%%)\n", I8 ? "i8 " : "f16", n,
         I8 ? 32 : 16, per, macs / per, I8 ? 8192 : 4096, 100.0 * macs / per / (I8 ? 8192 : 4096));
}

int main() {
  long long* dclk;
  cudaMalloc(&dclk, 148 * sizeof(long long));
  for (int n : {64, 96, 128, 160, 192, 208, 224, 240, 256}) run<false>(n, dclk);
  for (int n : {64, 96, 128, 160, 192, 208, 224, 240, 256}) run<true>(n, dclk);
  return 0;
}
