// LayerNorm forward/backward and CLS-concat + LayerNorm: one warp per row, 16-byte vector loads,
// row cached in registers, fp32 statistics via warp shuffles.  HBM-bound: 2*rows*d*elt bytes.
#include "common.cuh"

namespace vdr {

constexpr int kLnWarps = 8;

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void load8_f32(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// Row statistics + normalise for a row held as MAXC chunks of 8 per lane.
template <int MAXC>
__device__ __forceinline__ void ln_row(float (&x)[MAXC][8], int chunks, int lane, int d, float eps,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float& mean_out, float& rstd_out) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (lane + c * 32 < chunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s += x[c][i];
    }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (lane + c * 32 < chunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = x[c][i] - mean;
        q += t * t;
      }
    }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (lane + c * 32 < chunks) {
      const int col = (lane + c * 32) * 8;
      float g[8], b[8];
      load8_f32(gamma + col, g);
      load8_f32(beta + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[c][i] = (x[c][i] - mean) * rstd * g[i] + b[i];
    }
  mean_out = mean;
  rstd_out = rstd;
}

// Forward: gamma / beta are staged in shared memory once per block (re-reading them from L1 per row cost twice the
// payload in load traffic), each warp keeps the NEXT row's raw 16-byte vectors in flight while it reduces the current
// one (the kernel is HBM-bound: what matters is bytes in flight per SM), stores bypass L1.
template <int MAXC, bool OUT_F32>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, void* __restrict__ y, int64_t ldy,
                     float* __restrict__ mean, float* __restrict__ rstd, int rows, int d, float eps) {
  extern __shared__ __align__(16) float s_gb[];   // gamma[d] | beta[d]
  for (int i = threadIdx.x; i < d; i += kLnWarps * 32) {
    s_gb[i] = gamma[i];
    s_gb[d + i] = beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int chunks = d >> 3;
  const int64_t stride = (int64_t)gridDim.x * kLnWarps;
  int64_t row = blockIdx.x * (int64_t)kLnWarps + (threadIdx.x >> 5);
  uint4 nxt[MAXC];
  auto fetch = [&](int64_t r) {
    const __nv_bfloat16* xr = x + r * ldx;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < chunks) nxt[c] = __ldcs(reinterpret_cast<const uint4*>(xr + (lane + c * 32) * 8));
  };
  if (row < rows) fetch(row);
  for (; row < rows; row += stride) {
    float v[MAXC][8];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const float2 p0 = unpack_bf16x2(nxt[c].x), p1 = unpack_bf16x2(nxt[c].y), p2 = unpack_bf16x2(nxt[c].z), p3 = unpack_bf16x2(nxt[c].w);
      v[c][0] = p0.x; v[c][1] = p0.y; v[c][2] = p1.x; v[c][3] = p1.y; v[c][4] = p2.x; v[c][5] = p2.y; v[c][6] = p3.x; v[c][7] = p3.y;
    }
    if (row + stride < rows) fetch(row + stride);
    float mu, rs;
    ln_row<MAXC>(v, chunks, lane, d, eps, s_gb, s_gb + d, mu, rs);
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < chunks) {
        const int col = (lane + c * 32) * 8;
        if (OUT_F32) {
          float* yp = static_cast<float*>(y) + row * ldy + col;
          __stcs(reinterpret_cast<float4*>(yp), make_float4(v[c][0], v[c][1], v[c][2], v[c][3]));
          __stcs(reinterpret_cast<float4*>(yp + 4), make_float4(v[c][4], v[c][5], v[c][6], v[c][7]));
        } else {
          uint4 o;
          o.x = pack_bf16x2(v[c][0], v[c][1]); o.y = pack_bf16x2(v[c][2], v[c][3]);
          o.z = pack_bf16x2(v[c][4], v[c][5]); o.w = pack_bf16x2(v[c][6], v[c][7]);
          __stcs(reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(y) + row * ldy + col), o);
        }
      }
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
  }
}

// Y[0] = LN(cls), Y[1+i] = LN(X[i]); X, cls f32; Y bf16.
template <int MAXC>
__global__ void __launch_bounds__(kLnWarps * 32)
cls_concat_layernorm_kernel(const float* __restrict__ X, const float* __restrict__ cls,
                            const float* __restrict__ gamma, const float* __restrict__ beta,
                            __nv_bfloat16* __restrict__ Y, float* __restrict__ mean, float* __restrict__ rstd,
                            int n, int d, float eps) {
  const int lane = threadIdx.x & 31;
  const int chunks = d >> 3;
  for (int64_t row = blockIdx.x * (int64_t)kLnWarps + (threadIdx.x >> 5); row < (int64_t)n + 1;
       row += (int64_t)gridDim.x * kLnWarps) {
    const float* xr = (row == 0) ? cls : X + (row - 1) * d;
    float v[MAXC][8];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < chunks) load8_f32(xr + (lane + c * 32) * 8, v[c]);
    float mu, rs;
    ln_row<MAXC>(v, chunks, lane, d, eps, gamma, beta, mu, rs);
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < chunks) store8_bf16(Y + row * d + (lane + c * 32) * 8, v[c]);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += sum dy*xhat;  dbeta += sum dy.
template <int MAXC>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ x,
                     int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx, int64_t lddx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int d) {
  extern __shared__ float red[];  // [kLnWarps][2][d]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int chunks = d >> 3;
  float dg[MAXC][8], db[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) dg[c][i] = db[c][i] = 0.f;
  for (int64_t row = blockIdx.x * (int64_t)kLnWarps + warp; row < rows; row += (int64_t)gridDim.x * kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float xh[MAXC][8], gdy[MAXC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < chunks) {
        const int col = (lane + c * 32) * 8;
        float xv[8], dv[8], g[8];
        load8_bf16(x + row * ldx + col, xv);
        load8_bf16(dy + row * lddy + col, dv);
        load8_f32(gamma + col, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[c][i] = (xv[i] - mu) * rs;
          gdy[c][i] = dv[i] * g[i];
          s1 += gdy[c][i];
          s2 += gdy[c][i] * xh[c][i];
          dg[c][i] += dv[i] * xh[c][i];
          db[c][i] += dv[i];
        }
      }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < chunks) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (gdy[c][i] - s1 - xh[c][i] * s2);
        store8_bf16(dx + row * lddx + (lane + c * 32) * 8, o);
      }
  }
  // block reduce of dgamma / dbeta partials, then one atomicAdd per column per block
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (lane + c * 32 < chunks) {
      const int col = (lane + c * 32) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[(warp * 2 + 0) * d + col + i] = dg[c][i];
        red[(warp * 2 + 1) * d + col + i] = db[c][i];
      }
    }
  __syncthreads();
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) {
      a += red[(w * 2 + 0) * d + col];
      b += red[(w * 2 + 1) * d + col];
    }
    atomicAdd(dgamma + col, a);
    atomicAdd(dbeta + col, b);
  }
}

static int ln_grid(int64_t rows) {
  int64_t blocks = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// (sum, sum of squares) of each row: one warp per row, 16-byte loads, grid-stride over rows.  Read-only pass: rows*d*2 bytes.
__global__ void __launch_bounds__(kLnWarps * 32)
row_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int rows, int d, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * kLnWarps;
  for (int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5); row < rows; row += warps) {
    const __nv_bfloat16* xr = x + static_cast<int64_t>(row) * ldx;
    float s = 0.f, q = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
      float f[8];
      load8_bf16(xr + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s += f[i];
        q = fmaf(f[i], f[i], q);
      }
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) stats[row] = make_float2(s, q);
  }
}

// LayerNorm folded into the Linear that consumes it: one warp per output feature n.
__global__ void __launch_bounds__(kLnWarps * 32)
fold_layernorm_kernel(const __nv_bfloat16* __restrict__ W, int64_t ldw, const float* __restrict__ bias, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int N, int K, __nv_bfloat16* __restrict__ Wf, int64_t ldwf,
                      float* __restrict__ bias_f, float* __restrict__ colsum) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (n >= N) return;
  float cs = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = __bfloat162float(W[n * ldw + k]);
    const __nv_bfloat16 wf = __float2bfloat16_rn(w * gamma[k]);
    Wf[n * ldwf + k] = wf;
    cs += __bfloat162float(wf);
    bb = fmaf(beta[k], w, bb);
  }
  cs = warp_sum(cs);
  bb = warp_sum(bb);
  if (lane == 0) {
    colsum[n] = cs;
    bias_f[n] = (bias ? bias[n] : 0.f) + bb;
  }
}

}  // namespace vdr

extern "C" int vdr_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y,
                                 int64_t ldy, int y_dtype, float* mean, float* rstd, int rows, int d, float eps,
                                 vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(x && gamma && beta && y, VDR_EINVAL, "vdr_layernorm_fwd: null pointer");
  VDR_CHECK_ARG(rows > 0 && d > 0, VDR_EINVAL, "vdr_layernorm_fwd: bad shape rows=%d d=%d", rows, d);
  VDR_CHECK_ARG(d % 8 == 0 && d <= 4096, VDR_EINVAL, "vdr_layernorm_fwd: d (%d) must be a multiple of 8 and <= 4096", d);
  VDR_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= d && ldy >= d, VDR_EALIGN, "vdr_layernorm_fwd: ldx/ldy must be multiples of 8 and >= d");
  VDR_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), VDR_EALIGN, "vdr_layernorm_fwd: pointers must be 16-byte aligned");
  VDR_CHECK_ARG(y_dtype == VDR_DTYPE_BF16 || y_dtype == VDR_DTYPE_F32, VDR_EINVAL, "vdr_layernorm_fwd: bad y_dtype");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ln_grid(rows);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const size_t smem = (size_t)2 * d * sizeof(float);
#define VDR_LN_LAUNCH(MAXC)                                                                                   \
  if (y_dtype == VDR_DTYPE_F32)                                                                               \
    layernorm_fwd_kernel<MAXC, true><<<grid, kLnWarps * 32, smem, s>>>(xb, ldx, gamma, beta, y, ldy, mean, rstd, rows, d, eps); \
  else                                                                                                        \
    layernorm_fwd_kernel<MAXC, false><<<grid, kLnWarps * 32, smem, s>>>(xb, ldx, gamma, beta, y, ldy, mean, rstd, rows, d, eps);
  // chunks of 8 columns per lane: exact for the backbone widths (384 -> 2, 768 -> 3, 1024 -> 4), generic otherwise
  if (d <= 256) { VDR_LN_LAUNCH(1) }
  else if (d <= 512) { VDR_LN_LAUNCH(2) }
  else if (d <= 768) { VDR_LN_LAUNCH(3) }
  else if (d <= 1024) { VDR_LN_LAUNCH(4) }
  else { VDR_LN_LAUNCH(16) }
#undef VDR_LN_LAUNCH
  count_launch();
  VDR_CHECK_LAUNCH("layernorm_fwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_cls_concat_layernorm_fwd(const float* X, const float* cls, const float* gamma, const float* beta,
                                            void* Y_bf16, float* mean, float* rstd, int n, int d, float eps,
                                            vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(cls && gamma && beta && Y_bf16, VDR_EINVAL, "vdr_cls_concat_layernorm_fwd: null pointer");
  VDR_CHECK_ARG(n >= 0 && (n == 0 || X != nullptr), VDR_EINVAL, "vdr_cls_concat_layernorm_fwd: bad n / null X");
  VDR_CHECK_ARG(d % 8 == 0 && d > 0 && d <= 1024, VDR_EINVAL, "vdr_cls_concat_layernorm_fwd: d (%d) must be a multiple of 8 and <= 1024", d);
  VDR_CHECK_ARG(aligned16(X) && aligned16(cls) && aligned16(gamma) && aligned16(beta) && aligned16(Y_bf16), VDR_EALIGN, "vdr_cls_concat_layernorm_fwd: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ln_grid((int64_t)n + 1);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(Y_bf16);
  if (d <= 256) cls_concat_layernorm_kernel<1><<<grid, kLnWarps * 32, 0, s>>>(X, cls, gamma, beta, yb, mean, rstd, n, d, eps);
  else cls_concat_layernorm_kernel<4><<<grid, kLnWarps * 32, 0, s>>>(X, cls, gamma, beta, yb, mean, rstd, n, d, eps);
  count_launch();
  VDR_CHECK_LAUNCH("cls_concat_layernorm_kernel");
  return VDR_OK;
}

extern "C" int vdr_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* gamma,
                                 const float* mean, const float* rstd, void* dx, int64_t lddx, float* dgamma,
                                 float* dbeta, int rows, int d, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, VDR_EINVAL, "vdr_layernorm_bwd: null pointer");
  VDR_CHECK_ARG(rows > 0 && d > 0 && d % 8 == 0 && d <= 1024, VDR_EINVAL, "vdr_layernorm_bwd: d (%d) must be a multiple of 8 and <= 1024", d);
  VDR_CHECK_ARG(lddy % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0, VDR_EALIGN, "vdr_layernorm_bwd: leading dims must be multiples of 8");
  VDR_CHECK_ARG(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma), VDR_EALIGN, "vdr_layernorm_bwd: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int grid = ln_grid(rows);
  if (grid > num_sms() * 2) grid = num_sms() * 2;  // fewer blocks -> fewer atomics
  const size_t smem = (size_t)kLnWarps * 2 * d * sizeof(float);
  const __nv_bfloat16* dyb = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dxb = static_cast<__nv_bfloat16*>(dx);
  if (d <= 256) {
    layernorm_bwd_kernel<1><<<grid, kLnWarps * 32, smem, s>>>(dyb, lddy, xb, ldx, gamma, mean, rstd, dxb, lddx, dgamma, dbeta, rows, d);
  } else {
    static DeviceFlags configured;
    if (!configured.current()) {
      cudaError_t e = cudaFuncSetAttribute(layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(layernorm_bwd)");
      configured.current() = true;
    }
    layernorm_bwd_kernel<4><<<grid, kLnWarps * 32, smem, s>>>(dyb, lddy, xb, ldx, gamma, mean, rstd, dxb, lddx, dgamma, dbeta, rows, d);
  }
  count_launch();
  VDR_CHECK_LAUNCH("layernorm_bwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_row_stats(const void* x, int64_t ldx, int rows, int d, float* stats, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(x && stats, VDR_EINVAL, "vdr_row_stats: null pointer");
  VDR_CHECK_ARG(rows > 0 && d > 0 && d % 8 == 0, VDR_EINVAL, "vdr_row_stats: bad shape rows=%d d=%d (d must be a multiple of 8)", rows, d);
  VDR_CHECK_ARG(ldx % 8 == 0 && ldx >= d && aligned16(x) && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, VDR_EALIGN,
                "vdr_row_stats: x must be 16-byte aligned with ldx %% 8 == 0, stats 8-byte aligned");
  row_stats_kernel<<<ln_grid(rows), kLnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, rows, d, reinterpret_cast<float2*>(stats));
  count_launch();
  VDR_CHECK_LAUNCH("row_stats_kernel");
  return VDR_OK;
}

extern "C" int vdr_fold_layernorm(const void* W, int64_t ldw, const float* bias, const float* gamma, const float* beta, int N, int K,
                                  void* Wf, int64_t ldwf, float* bias_f, float* colsum, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(W && gamma && beta && Wf && bias_f && colsum, VDR_EINVAL, "vdr_fold_layernorm: null pointer");
  VDR_CHECK_ARG(N > 0 && K > 0 && ldw >= K && ldwf >= K, VDR_EINVAL, "vdr_fold_layernorm: bad shape N=%d K=%d", N, K);
  fold_layernorm_kernel<<<(N + kLnWarps - 1) / kLnWarps, kLnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(W), ldw, bias, gamma, beta, N, K, static_cast<__nv_bfloat16*>(Wf), ldwf, bias_f, colsum);
  count_launch();
  VDR_CHECK_LAUNCH("fold_layernorm_kernel");
  return VDR_OK;
}
