// Micro-benchmark: issue rate of the instructions the epilogue / softmax code leans on (per SM sub-partition).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

constexpr int kIters = 2048, kChains = 8;

template <int OP>
__global__ void bench(float* out, long long* cycles, float seed) {
  uint64_t a2[kChains];
  float a[kChains * 2];
  for (int i = 0; i < kChains; ++i) { a[2 * i] = seed + i + threadIdx.x; a[2 * i + 1] = seed * 2 + i; a2[i] = pack2(a[2 * i], a[2 * i + 1]); }
  const uint64_t b2 = pack2(seed * 0.999f, seed * 1.001f);
  const float b = seed * 0.999f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a2[i]) : "l"(b2));                       // FFMA2 reg
      if (OP == 1) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[2 * i + 1]) : "f"(b)); }  // 2 x FFMA reg
      if (OP == 2) { asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(a[2 * i + 1]) : "f"(b)); }  // 2 x FFMA imm
      if (OP == 3) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(b2));                               // FMUL2
      if (OP == 4) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(b2));                               // FADD2
      if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i + 1])); }   // 2 x MUFU.EX2
      if (OP == 6) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i + 1])); }   // 2 x MUFU.RCP
      if (OP == 7) { asm volatile("max.f32 %0, %0, %1;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[2 * i + 1]) : "f"(b)); }  // 2 x FMNMX
      if (OP == 8) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[2 * i]), "f"(a[2 * i + 1])); a[2 * i] = __uint_as_float(r); }  // F2FP
      if (OP == 10) { asm volatile("add.f32 %0, %0, %1;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[2 * i + 1]) : "f"(b)); }   // 2 x FADD reg
      if (OP == 11) { asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("mul.f32 %0, %0, 0f3F7FBE77;" : "+f"(a[2 * i + 1])); }   // FMUL reg + FMUL imm
      if (OP == 12) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[2 * i]) : "f"(b), "f"(a[2 * i + 1])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[2 * i + 1]) : "f"(b), "f"(a[2 * i])); }  // 2 x FFMA 3 distinct regs
      if (OP == 13) { uint32_t u = __float_as_uint(a[2 * i]), w = __float_as_uint(a[2 * i + 1]); asm volatile("mad.lo.u32 %0, %1, 8388608, %0;" : "+r"(u) : "r"(w)); a[2 * i] = __uint_as_float(u); }  // IMAD shift-add
      if (OP == 14) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[2 * i]) : "f"(b), "f"(a[2 * i + 1])); }   // FMNMX3
      if (OP == 15) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(b2), "l"(a2[(i + 1) % kChains])); }   // FFMA2 3 distinct regs
      if (OP == 16) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i + 1])); }  // FFMA + MUFU mix
      if (OP == 17) { asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(a[2 * i]) : "f"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[2 * i + 1]) : "f"(b)); }  // FFMA imm + FMNMX mix
      if (OP == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a2[i]) : "l"(b2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[2 * i])); }  // FFMA2 + MUFU mix
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < kChains; ++i) { float x, y; unpack2(a2[i], x, y); s += x + y + a[2 * i] + a[2 * i + 1]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_instr) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    bench<OP><<<148, 128 * warps_per_smsp>>>(out, cyc, 1.0f);
    bench<OP><<<148, 128 * warps_per_smsp>>>(out, cyc, 1.0f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_instr = (double)c / (kIters * kChains * per_iter_instr) / 1.0;
    printf("%-22s warps/SMSP=%d  %.2f clk per warp-instruction per warp  -> %.2f clk per instr per SMSP\n", name, warps_per_smsp, per_instr, per_instr / warps_per_smsp);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA2 (reg)", 1);
  run<1>("FFMA (reg) x2", 2);
  run<2>("FFMA (imm) x2", 2);
  run<3>("FMUL2", 1);
  run<4>("FADD2", 1);
  run<5>("MUFU.EX2 x2", 2);
  run<6>("MUFU.RCP x2", 2);
  run<7>("FMNMX x2", 2);
  run<8>("F2FP.BF16 pack", 1);
  run<9>("FFMA2 + MUFU.EX2", 2);
  run<10>("FADD (reg) x2", 2);
  run<11>("FMUL reg + FMUL imm", 2);
  run<12>("FFMA 3 regs x2", 2);
  run<13>("IMAD shl-add", 1);
  run<14>("FMNMX3", 1);
  run<15>("FFMA2 3 regs", 1);
  run<16>("FFMA + MUFU.EX2", 2);
  run<17>("FFMA imm + FMNMX", 2);
  return 0;
}
