// TMEM read-out micro-benchmark (sm_100a): how many bytes per clock does one SM move TMEM -> registers with
// tcgen05.ld.32x32b.xN, as a function of the number of reading warps, the load width and the loads in flight?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu ; run: ./tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(taddr));
}

template <int X> __device__ __forceinline__ void tmem_ld(uint32_t t, uint32_t (&r)[X]);
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t t, uint32_t (&r)[16]) { tmem_ld_x16(t, r); }
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t t, uint32_t (&r)[32]) { tmem_ld_x32(t, r); }
template <> __device__ __forceinline__ void tmem_ld<64>(uint32_t t, uint32_t (&r)[64]) { tmem_ld_x64(t, r); }

template <int X, int INFLIGHT>
__global__ void __launch_bounds__(512) bench(int iters, long long* clocks, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t col = ((warp >> 2) * X * INFLIGHT) & 511;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[INFLIGHT][X];
#pragma unroll
    for (int f = 0; f < INFLIGHT; ++f) tmem_ld<X>(base + ((col + f * X) & 511), r[f]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int f = 0; f < INFLIGHT; ++f)
#pragma unroll
      for (int j = 0; j < X; ++j) acc ^= r[f][j];
    col = (col + X * INFLIGHT) & 511;
    if (col + X * INFLIGHT > 512) col = 0;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
  }
}

template <int X, int INFLIGHT>
void run(int warps, long long* dclk, uint32_t* dsink) {
  const int iters = 2000, grid = 148;
  bench<X, INFLIGHT><<<grid, warps * 32>>>(iters, dclk, dsink);
  bench<X, INFLIGHT><<<grid, warps * 32>>>(iters, dclk, dsink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("x%d inflight %d warps %d: %s\n", X, INFLIGHT, warps, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, dclk, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double bytes = (double)iters * warps * INFLIGHT * 32.0 * X * 4.0;
  printf("tcgen05.ld.32x32b.x%-2d  loads in flight %d  warps %2d : %8.0f clk for %.1f MB per SM = %6.1f B/clk/SM  (%5.1f clk per load)\n",
         X, INFLIGHT, warps, mean, bytes / 1e6, bytes / mean, mean / (iters * INFLIGHT));
}

int main() {
  long long* dclk; uint32_t* dsink;
  cudaMalloc(&dclk, 148 * sizeof(long long));
  cudaMalloc(&dsink, 148 * 512 * sizeof(uint32_t));
  for (int warps : {4, 8, 16}) {
    run<16, 1>(warps, dclk, dsink); run<16, 2>(warps, dclk, dsink); run<16, 4>(warps, dclk, dsink);
    run<32, 1>(warps, dclk, dsink); run<32, 2>(warps, dclk, dsink); run<32, 4>(warps, dclk, dsink);
    run<64, 1>(warps, dclk, dsink); run<64, 2>(warps, dclk, dsink);
  }
  return 0;
}
