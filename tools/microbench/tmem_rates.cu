// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as the attention softmax warps use them
// (32x32b.x32: one TMEM lane per thread, 32 consecutive 32-bit columns; 64 columns = one thread's share of a 128 x 128 S block).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rates tmem_rates.cu && ./tmem_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD32(taddr, r)                                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                       \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                       \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                     \
               : "r"(taddr) : "memory")
#define ST32(taddr, r)                                                                                                       \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                      \
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                              \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),      \
               "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),     \
               "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

// MODE 0: two loads (64 columns) + wait per iteration; 1: one store (32 columns) + wait; 2: loads + store
template <int MODE>
__global__ void __launch_bounds__(256, 2) tmem_bench(uint32_t* out, int iters, long long* cyc) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tptr;
  const uint32_t taddr = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t acc = 0, r0[32], r1[32];
  for (int i = 0; i < 32; ++i) { r0[i] = threadIdx.x + i; r1[i] = i; }
  ST32(taddr, r0);
  ST32(taddr + 32, r1);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      LD32(taddr, r0);
      LD32(taddr + 32, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r0[0] ^ r0[31] ^ r1[5] ^ r1[17];
    }
    if (MODE == 1 || MODE == 2) {
      r0[0] = acc + it;
      ST32(tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 128 + (warp >> 2) * 32, r0);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 256 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

template <int MODE>
void run(const char* name, int bytes_per_thread_it) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 296 * 256 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4000;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    tmem_bench<MODE><<<148 * ctas, 256>>>(out, iters, cyc);
    tmem_bench<MODE><<<148 * ctas, 256>>>(out, iters, cyc);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_it = (double)c / iters;
    printf("%-34s CTAs/SM=%d (%2d warps)  %7.1f clk per iteration -> %6.1f B/clk/SM   (%s)\n", name, ctas, 8 * ctas, per_it,
           ctas * 256.0 * bytes_per_thread_it / per_it, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tcgen05.ld 2 x (32x32b.x32) + wait", 256);
  run<1>("tcgen05.st 32x32b.x32 + wait", 128);
  run<2>("ld 64 columns + st 32 columns", 384);
  return 0;
}
