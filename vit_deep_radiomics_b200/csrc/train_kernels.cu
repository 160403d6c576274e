// Kernels that only the classifier's training step needs (SURVEY.md rows K10-K12 backward):
// GELU fwd/bwd, bf16 transpose (operands of the wgrad / dgrad GEMMs), column sums (bias grads),
// softmax recompute / dS for the attention backward, CLS-concat LayerNorm backward, and the
// fused classification head (dense1 -> GELU -> dense2) forward + backward.
#include "common.cuh"

namespace vdr {

__device__ __forceinline__ float gelu_grad(float x) {
  // d/dx [0.5 x (1 + erf(x/sqrt2))] = 0.5 (1 + erf(x/sqrt2)) + x * exp(-x^2/2) / sqrt(2 pi)
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ h, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(z)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      o[k] = pack_bf16x2(gelu_erf(f.x), gelu_erf(f.y));
    }
    reinterpret_cast<uint4*>(h)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ z,
                                                       __nv_bfloat16* __restrict__ dz, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 a = reinterpret_cast<const uint4*>(dh)[i], b = reinterpret_cast<const uint4*>(z)[i];
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = unpack_bf16x2(wa[k]), x = unpack_bf16x2(wb[k]);
      o[k] = pack_bf16x2(g.x * gelu_grad(x.x), g.y * gelu_grad(x.y));
    }
    reinterpret_cast<uint4*>(dz)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[c][r] = in[r][c]; in (rows, cols) with pitch ld_in, out (cols, rows) with pitch ld_out (bf16).
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in,
                                                             __nv_bfloat16* __restrict__ out, int64_t ld_out, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<int64_t>(r) * ld_in + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[static_cast<int64_t>(c) * ld_out + r] = tile[threadIdx.x][j];
  }
}

// out[c] += sum_r in[r][c]   (bf16 in, f32 accumulate; one atomicAdd per column per block)
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld, int rows, int cols,
                                                     float* __restrict__ out, int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  int r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(in[static_cast<int64_t>(r) * ld + c]);
  atomicAdd(out + c, acc);
}

// delta[i] = sum_c dO[i][c] * O[i][c] per head  (rows N, head h covers columns h*64..h*64+63): one warp per (row, head)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                                         int64_t ld, int N, int heads, float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<int64_t>(N) * heads) return;
  const int h = static_cast<int>(wid % heads);
  const int64_t i = wid / heads;
  const int64_t off = i * ld + h * 64 + lane * 2;
  const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dO + off));
  const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(O + off));
  const float s = warp_sum(a.x * b.x + a.y * b.y);
  if (lane == 0) delta[static_cast<int64_t>(h) * N + i] = s;
}

// P = exp(S*scale - lse[i]) (bf16, zero for key columns >= N);  in place option: dS = P * (dP - delta[i]) * scale.
// S, dP: (N, ldp) f32;  P, dS: (N, ldp) bf16.  mode 0: write P;  mode 1: read S and dP, write dS (and P is recomputed).
__global__ void __launch_bounds__(256) attn_p_ds_kernel(const float* __restrict__ S, const float* __restrict__ dP,
                                                        const float* __restrict__ lse, const float* __restrict__ delta,
                                                        __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ dS,
                                                        int N, int64_t ldp, float scale) {
  const int64_t total = static_cast<int64_t>(N) * ldp;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / ldp;
    const int j = static_cast<int>(e - i * ldp);
    float p = 0.f, ds = 0.f;
    if (j < N) {
      p = __expf(S[e] * scale - lse[i]);
      ds = p * (dP[e] - delta[i]) * scale;
    }
    P[e] = __float2bfloat16_rn(p);
    dS[e] = __float2bfloat16_rn(ds);
  }
}

// Backward of Y = LN(cat(cls, X)) restricted to what training needs: dgamma, dbeta (accumulated) and
// dcls (accumulated) -- X is data, it takes no gradient.
__global__ void __launch_bounds__(256)
cls_concat_layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dY, const float* __restrict__ X, const float* __restrict__ cls,
                                const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcls, int n, int d,
                                int rows_per_block) {
  // thread c handles column c (d <= 1024 -> up to 4 columns per thread), loops over this block's rows
  const int r0 = blockIdx.x * rows_per_block;
  int r1 = r0 + rows_per_block;
  if (r1 > n + 1) r1 = n + 1;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float dg = 0.f, db = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float xv = (r == 0) ? cls[c] : X[static_cast<int64_t>(r - 1) * d + c];
      const float g = __bfloat162float(dY[static_cast<int64_t>(r) * d + c]);
      dg += g * (xv - mean[r]) * rstd[r];
      db += g;
    }
    atomicAdd(dgamma + c, dg);
    atomicAdd(dbeta + c, db);
  }
  if (blockIdx.x == 0) {  // dx of row 0 (the CLS token): needs two row reductions
    __shared__ float s1s, s2s;
    __shared__ float red[2][8];
    float a = 0.f, b = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float gdy = __bfloat162float(dY[c]) * gamma[c];
      const float xh = (cls[c] - mean[0]) * rstd[0];
      a += gdy;
      b += gdy * xh;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float x = 0.f, y = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { x += red[0][w]; y += red[1][w]; }
      s1s = x / d; s2s = y / d;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float gdy = __bfloat162float(dY[c]) * gamma[c];
      const float xh = (cls[c] - mean[0]) * rstd[0];
      atomicAdd(dcls + c, rstd[0] * (gdy - s1s - xh * s2s));
    }
  }
}

// ---- classification head (MLPLayer, models_archs.py:186-200): zc = W1 cls + b1; hc = gelu(zc); logits = W2 hc + b2, fp32 weights.
// One warp per hidden unit, eight per CTA, H1 / 8 CTAs: W1 (d x H1 floats, 4.7 MB for d = 768) is streamed by the whole GPU
// instead of by one CTA (0.7 ms -> a few microseconds).
constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32)
cls_head_hidden_kernel(const __nv_bfloat16* __restrict__ cls, const float* __restrict__ W1, const float* __restrict__ b1,
                       float* __restrict__ zc, int d, int H1) {
  const int lane = threadIdx.x & 31, j = blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
  if (j >= H1) return;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(__ldg(W1 + static_cast<int64_t>(j) * d + c), __bfloat162float(cls[c]), acc);
  acc = warp_sum(acc);
  if (lane == 0) zc[j] = acc + b1[j];
}

__global__ void __launch_bounds__(256)
cls_head_out_kernel(const float* __restrict__ zc, const float* __restrict__ W2, const float* __restrict__ b2,
                    float* __restrict__ logits, int H1, int C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = warp; k < C; k += nw) {
    float acc = 0.f;
    for (int j = lane; j < H1; j += 32) acc = fmaf(W2[static_cast<int64_t>(k) * H1 + j], gelu_erf(zc[j]), acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[k] = acc + b2[k];
  }
}

// Backward.  Gradients are ACCUMULATED into dW1, db1, dW2, db2; dcls (d) = W1^T dzc + dcls_in.
__global__ void __launch_bounds__(256)
cls_head_bwd_init_kernel(const float* __restrict__ dlogits, const float* __restrict__ dcls_in, float* __restrict__ db2,
                         float* __restrict__ dcls, int d, int C) {
  for (int c = threadIdx.x; c < d; c += blockDim.x) dcls[c] = dcls_in ? dcls_in[c] : 0.f;
  if (static_cast<int>(threadIdx.x) < C) db2[threadIdx.x] += dlogits[threadIdx.x];
}

__global__ void __launch_bounds__(kHeadWarps * 32)
cls_head_bwd_rows_kernel(const __nv_bfloat16* __restrict__ cls, const float* __restrict__ W1, const float* __restrict__ W2,
                         const float* __restrict__ zc, const float* __restrict__ dlogits, float* __restrict__ dW1,
                         float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ dcls, int d, int H1, int C) {
  extern __shared__ float s_dcls[];   // [d] this CTA's partial W1^T dzc
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_dcls[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, j = blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
  if (j < H1) {
    const float z = zc[j], h = gelu_erf(z);
    float dh = 0.f;
    for (int k = 0; k < C; ++k) {
      const float dl = dlogits[k];
      dh = fmaf(W2[static_cast<int64_t>(k) * H1 + j], dl, dh);
      if (lane == 0) dW2[static_cast<int64_t>(k) * H1 + j] += dl * h;
    }
    const float dz = dh * gelu_grad(z);
    if (lane == 0) db1[j] += dz;
    for (int c = lane; c < d; c += 32) {
      const int64_t e = static_cast<int64_t>(j) * d + c;
      dW1[e] += dz * __bfloat162float(cls[c]);
      atomicAdd(&s_dcls[c], __ldg(W1 + e) * dz);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) atomicAdd(dcls + c, s_dcls[c]);
}

static int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace vdr

using namespace vdr;
#define S_(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" int vdr_gelu_fwd(const void* z, void* h, int64_t n, vdr_stream_t stream) {
  VDR_CHECK_ARG(z && h && n > 0 && n % 8 == 0, VDR_EINVAL, "vdr_gelu_fwd: null pointer or n %% 8 != 0");
  VDR_CHECK_ARG(aligned16(z) && aligned16(h), VDR_EALIGN, "vdr_gelu_fwd: pointers must be 16-byte aligned");
  gelu_fwd_kernel<<<ew_grid(n / 8), 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(z), static_cast<__nv_bfloat16*>(h), n / 8);
  count_launch();
  VDR_CHECK_LAUNCH("gelu_fwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_gelu_bwd(const void* dh, const void* z, void* dz, int64_t n, vdr_stream_t stream) {
  VDR_CHECK_ARG(dh && z && dz && n > 0 && n % 8 == 0, VDR_EINVAL, "vdr_gelu_bwd: null pointer or n %% 8 != 0");
  VDR_CHECK_ARG(aligned16(dh) && aligned16(z) && aligned16(dz), VDR_EALIGN, "vdr_gelu_bwd: pointers must be 16-byte aligned");
  gelu_bwd_kernel<<<ew_grid(n / 8), 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(dh), static_cast<const __nv_bfloat16*>(z),
                                                          static_cast<__nv_bfloat16*>(dz), n / 8);
  count_launch();
  VDR_CHECK_LAUNCH("gelu_bwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols, vdr_stream_t stream) {
  VDR_CHECK_ARG(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, VDR_EINVAL, "vdr_transpose_bf16: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  VDR_CHECK_ARG(grid.y <= 65535, VDR_EINVAL, "vdr_transpose_bf16: too many rows");
  transpose_bf16_kernel<<<grid, block, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  count_launch();
  VDR_CHECK_LAUNCH("transpose_bf16_kernel");
  return VDR_OK;
}

extern "C" int vdr_colsum_bf16(const void* in, int64_t ld, int rows, int cols, float* out_accum, vdr_stream_t stream) {
  VDR_CHECK_ARG(in && out_accum && rows > 0 && cols > 0 && ld >= cols, VDR_EINVAL, "vdr_colsum_bf16: bad arguments");
  const int rpb = 64;
  dim3 grid((cols + 255) / 256, (rows + rpb - 1) / rpb);
  VDR_CHECK_ARG(grid.y <= 65535, VDR_EINVAL, "vdr_colsum_bf16: too many rows");
  colsum_kernel<<<grid, 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(in), ld, rows, cols, out_accum, rpb);
  count_launch();
  VDR_CHECK_LAUNCH("colsum_kernel");
  return VDR_OK;
}

extern "C" int vdr_attn_delta(const void* dO, const void* O, int64_t ld, int N, int heads, float* delta, vdr_stream_t stream) {
  VDR_CHECK_ARG(dO && O && delta && N > 0 && heads > 0 && ld >= heads * 64, VDR_EINVAL, "vdr_attn_delta: bad arguments");
  const int64_t warps = (int64_t)N * heads;
  attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(dO),
                                                                                static_cast<const __nv_bfloat16*>(O), ld, N, heads, delta);
  count_launch();
  VDR_CHECK_LAUNCH("attn_delta_kernel");
  return VDR_OK;
}

extern "C" int vdr_attn_p_ds(const float* S, const float* dP, const float* lse, const float* delta, void* P, void* dS, int N,
                             int64_t ldp, float scale, vdr_stream_t stream) {
  VDR_CHECK_ARG(S && dP && lse && delta && P && dS && N > 0 && ldp >= N, VDR_EINVAL, "vdr_attn_p_ds: bad arguments");
  attn_p_ds_kernel<<<ew_grid((int64_t)N * ldp), 256, 0, S_(stream)>>>(S, dP, lse, delta, static_cast<__nv_bfloat16*>(P),
                                                                     static_cast<__nv_bfloat16*>(dS), N, ldp, scale);
  count_launch();
  VDR_CHECK_LAUNCH("attn_p_ds_kernel");
  return VDR_OK;
}

extern "C" int vdr_cls_concat_layernorm_bwd(const void* dY, const float* X, const float* cls, const float* gamma,
                                            const float* mean, const float* rstd, float* dgamma, float* dbeta, float* dcls,
                                            int n, int d, vdr_stream_t stream) {
  VDR_CHECK_ARG(dY && cls && gamma && mean && rstd && dgamma && dbeta && dcls && (n == 0 || X), VDR_EINVAL, "vdr_cls_concat_layernorm_bwd: null pointer");
  VDR_CHECK_ARG(n >= 0 && d > 0, VDR_EINVAL, "vdr_cls_concat_layernorm_bwd: bad shape");
  const int rpb = 32;
  cls_concat_layernorm_bwd_kernel<<<(n + 1 + rpb - 1) / rpb, 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(dY), X, cls, gamma, mean,
                                                                                 rstd, dgamma, dbeta, dcls, n, d, rpb);
  count_launch();
  VDR_CHECK_LAUNCH("cls_concat_layernorm_bwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_cls_head_fwd(const void* cls_bf16, const float* W1, const float* b1, const float* W2, const float* b2,
                                float* zc, float* logits, int d, int H1, int C, vdr_stream_t stream) {
  VDR_CHECK_ARG(cls_bf16 && W1 && b1 && W2 && b2 && zc && logits, VDR_EINVAL, "vdr_cls_head_fwd: null pointer");
  VDR_CHECK_ARG(d > 0 && H1 > 0 && C > 0, VDR_EINVAL, "vdr_cls_head_fwd: bad shape");
  cls_head_hidden_kernel<<<(H1 + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(cls_bf16), W1, b1, zc, d, H1);
  cls_head_out_kernel<<<1, 256, 0, S_(stream)>>>(zc, W2, b2, logits, H1, C);
  count_launch(2);
  VDR_CHECK_LAUNCH("cls_head_fwd kernels");
  return VDR_OK;
}

extern "C" int vdr_cls_head_bwd(const void* cls_bf16, const float* W1, const float* W2, const float* zc, const float* dlogits,
                                const float* dcls_in, float* dW1, float* db1, float* dW2, float* db2, float* dcls, int d,
                                int H1, int C, vdr_stream_t stream) {
  VDR_CHECK_ARG(cls_bf16 && W1 && W2 && zc && dlogits && dW1 && db1 && dW2 && db2 && dcls, VDR_EINVAL, "vdr_cls_head_bwd: null pointer");
  VDR_CHECK_ARG(d > 0 && H1 > 0 && C > 0 && C <= 256 && (size_t)d * 4 <= 48 * 1024, VDR_EINVAL, "vdr_cls_head_bwd: bad shape");
  cls_head_bwd_init_kernel<<<1, 256, 0, S_(stream)>>>(dlogits, dcls_in, db2, dcls, d, C);
  cls_head_bwd_rows_kernel<<<(H1 + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, d * sizeof(float), S_(stream)>>>(
      static_cast<const __nv_bfloat16*>(cls_bf16), W1, W2, zc, dlogits, dW1, db1, dW2, dcls, d, H1, C);
  count_launch(2);
  VDR_CHECK_LAUNCH("cls_head_bwd kernels");
  return VDR_OK;
}
