// Micro-benchmark: the per-block arithmetic of the attention softmax warps in isolation (no TMEM, no MMAs, no mbarriers):
// 16 warps per SM (2 CTAs x 8 warps), each thread 64 scores per iteration from shared memory (stand-in for tcgen05.ld),
// block maximum, 64 exponentials, row sum, bf16 packing, 32 words back to shared memory (stand-in for tcgen05.st).
// Prints cycles per iteration per SM against the MUFU bound, for several instruction mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_math softmax_math.cu && ./softmax_math
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t magic2 = pack2(12582912.f, 12582912.f);
  const uint64_t t2 = add2(x2, magic2);
  const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), x2);
  uint64_t q2 = fma2(f2, pack2(0.05508868396282196f, 0.05508868396282196f), pack2(0.24260404706001282f, 0.24260404706001282f));
  q2 = fma2(q2, f2, pack2(0.6932762265205383f, 0.6932762265205383f));
  q2 = fma2(q2, f2, pack2(0.9999289512634277f, 0.9999289512634277f));
  float t0, t1, q0, q1;
  unpack2(t2, t0, t1);
  unpack2(q2, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// VARIANT bits: 0..7 poly mask | 0x100 scalar FFMA/FADD instead of packed | 0x200 no row sum | 0x400 no pair exchange
//               0x800 no bf16 packing (xor the raw bits) | 0x1000 exponentials only (no max, no scale)
template <int VARIANT>
__global__ void __launch_bounds__(256, 2) softmax_math(float* out, int iters, long long* cyc, float scale_log2) {
  extern __shared__ uint4 sm4[];
  uint4* s_in = sm4;                  // [16][256] float4
  uint4* s_out = sm4 + 16 * 256;      // [8][256]
  float* s_max = reinterpret_cast<float*>(sm4 + 24 * 256);   // [2][256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;
  for (int i = 0; i < 16; ++i) {
    uint32_t h = (tid * 16 + i) * 2654435761u + blockIdx.x;
    float4 v;
    v.x = ((h >> 8) & 1023) * (8.f / 1024.f); v.y = ((h >> 10) & 1023) * (8.f / 1024.f);
    v.z = ((h >> 12) & 1023) * (8.f / 1024.f); v.w = ((h >> 14) & 1023) * (8.f / 1024.f);
    s_in[i * 256 + tid] = *reinterpret_cast<uint4*>(&v);
  }
  __syncthreads();
  constexpr uint32_t kPoly = VARIANT & 0xff;
  float m_ref = -INFINITY, l_run = 0.f;
  uint32_t acc = 0;
  const uint64_t scale2 = pack2(scale_log2, scale_log2);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t sr[2][32];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_in + i * 256 + tid))));
      sr[i >> 3][(i & 7) * 4 + 0] = u.x; sr[i >> 3][(i & 7) * 4 + 1] = u.y; sr[i >> 3][(i & 7) * 4 + 2] = u.z; sr[i >> 3][(i & 7) * 4 + 3] = u.w;
    }
    float alpha = 1.f;
    if (!(VARIANT & 0x1000)) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = max3(mx0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
        mx1 = max3(mx1, __uint_as_float(sr[0][i + 2]), __uint_as_float(sr[0][i + 3]));
        mx2 = max3(mx2, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
        mx3 = max3(mx3, __uint_as_float(sr[1][i + 2]), __uint_as_float(sr[1][i + 3]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (!(VARIANT & 0x400)) {
        float* xmax = s_max + (it & 1) * 256;
        xmax[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
        mx = fmaxf(mx, xmax[(half ^ 1) * 128 + row]);
      }
      const float m_new = fmaxf(m_ref, mx * scale_log2);
      const bool moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);
      if (moved) {
        alpha = ex2(m_ref - m_new);
        m_ref = m_new;
      }
    }
    const uint64_t negm2 = pack2(-m_ref, -m_ref);
    uint64_t lsum2 = 0ull;
    float lsum_a = 0.f, lsum_b = 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0, p1;
        if (VARIANT & 0x1000) {
          p0 = ex2(__uint_as_float(sr[c][i]));
          p1 = ex2(__uint_as_float(sr[c][i + 1]));
        } else if (VARIANT & 0x100) {
          const float x0 = fmaf(__uint_as_float(sr[c][i]), scale_log2, -m_ref), x1 = fmaf(__uint_as_float(sr[c][i + 1]), scale_log2, -m_ref);
          if ((kPoly >> ((i >> 1) & 7)) & 1u) exp2_poly2(pack2(x0, x1), p0, p1);
          else { p0 = ex2(x0); p1 = ex2(x1); }
        } else {
          const uint64_t x2 = fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), scale2, negm2);
          if ((kPoly >> ((i >> 1) & 7)) & 1u) {
            exp2_poly2(x2, p0, p1);
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
        }
        if (!(VARIANT & 0x200)) {
          if (VARIANT & 0x100) { lsum_a += p0; lsum_b += p1; }
          else lsum2 = add2(lsum2, pack2(p0, p1));
        }
        if (VARIANT & 0x800) pk[c * 16 + (i >> 1)] = __float_as_uint(p0) ^ __float_as_uint(p1);
        else pk[c * 16 + (i >> 1)] = cvt_bf16x2(p0, p1);
      }
    }
    float l0, l1;
    unpack2(lsum2, l0, l1);
    l_run = l_run * alpha + (l0 + l1) + (lsum_a + lsum_b);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_out + i * 256 + tid))),
                   "r"(pk[4 * i]), "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
    acc ^= pk[it & 31];
  }
  const long long t1 = clock64();
  out[blockIdx.x * 256 + tid] = l_run + __uint_as_float(acc & 0x3fffffff);
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int VARIANT>
void run(const char* name, int mufu_per_thread) {
  float* out; long long* cyc;
  cudaMalloc(&out, 296 * 256 * 4); cudaMalloc(&cyc, 8);
  const int smem = (24 * 256) * 16 + 2 * 256 * 4, iters = 4000;
  cudaFuncSetAttribute(softmax_math<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
    softmax_math<VARIANT><<<148 * ctas_per_sm, 256, smem>>>(out, iters, cyc, 0.18f);
    softmax_math<VARIANT><<<148 * ctas_per_sm, 256, smem>>>(out, iters, cyc, 0.18f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    const double per_it = (double)c / iters;                       // cycles per iteration of one CTA = per 128x128 block of each resident CTA
    const double bound = ctas_per_sm * 8.0 * mufu_per_thread * 8.0 / 4.0;   // MUFU: one warp instruction per 8 cycles per scheduler
    printf("%-44s CTAs/SM=%d  %7.0f clk per block round  MUFU bound %5.0f  -> %.2f of MUFU peak   (%s)\n", name, ctas_per_sm, per_it, bound, bound / per_it, cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0x02>("kernel mix (packed, poly 1/8)", 57);
  run<0x00>("packed, no poly", 65);
  run<0x06>("packed, poly 2/8", 49);
  run<0x102>("scalar FFMA/FADD, poly 1/8", 57);
  run<0x100>("scalar FFMA/FADD, no poly", 65);
  run<0x202>("packed, poly 1/8, no row sum", 57);
  run<0x402>("packed, poly 1/8, no pair exchange", 57);
  run<0x802>("packed, poly 1/8, no bf16 packing", 57);
  run<0x602>("packed, poly 1/8, no sum, no exchange", 57);
  run<0x1000>("exponentials + pack only", 64);
  run<0x1a00>("exponentials only", 64);
  return 0;
}
