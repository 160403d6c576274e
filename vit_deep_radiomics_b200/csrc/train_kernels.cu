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

// eight consecutive columns (one 16-byte vector) of dropout site `d`: multipliers 0 / scale (all 1 when thr16 == 0).
// i = index of the 8-column vector in a contiguous (rows, cols) matrix with cols % 8 == 0
__device__ __forceinline__ void drop8(const DropSpec& d, int64_t i, int cols8, float (&f)[8]) {
  if (d.thr16 == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 1.f;
    return;
  }
  const uint4 bits = drop_bits8(d, static_cast<uint64_t>(i / cols8), static_cast<uint32_t>(i % cols8));
  const float sc = drop_scale(d);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = drop_lane16(bits, k) >= d.thr16 ? sc : 0.f;
}

// h = dropout(gelu(z)): the feed-forward block's inner dropout (nn.TransformerEncoderLayer._ff_block) fused into the activation
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ h, int64_t n8,
                                                       int cols8, DropSpec drop) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(z)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float m[8];
    drop8(drop, i, cols8, m);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      o[k] = pack_bf16x2(gelu_erf(f.x) * m[2 * k], gelu_erf(f.y) * m[2 * k + 1]);
    }
    reinterpret_cast<uint4*>(h)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out = x * mask / (1 - p) (bf16 matrix, contiguous columns, row pitches ld): the backward of a dropout that the forward applied
// inside a GEMM epilogue (sub-layer outputs before the residual add)
__global__ void __launch_bounds__(256) dropout_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                                            int64_t ldo, int64_t rows, int cols8, DropSpec drop) {
  const int64_t n8 = rows * cols8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols8;
    const int c8 = static_cast<int>(i % cols8);
    const uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + c8 * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float m[8];
    drop8(drop, i, cols8, m);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      o[k] = pack_bf16x2(f.x * m[2 * k], f.y * m[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(out + r * ldo + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// the mask itself (1 = kept), for tests and for reference computations that must use the device's masks
__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ out, int64_t rows, int64_t cols, DropSpec drop) {
  const int64_t n = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = drop_factor(drop, static_cast<uint64_t>(i / cols), static_cast<uint32_t>(i % cols)) != 0.f;
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ z,
                                                       __nv_bfloat16* __restrict__ dz, int64_t n8, int cols8, DropSpec drop) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 a = reinterpret_cast<const uint4*>(dh)[i], b = reinterpret_cast<const uint4*>(z)[i];
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
    float m[8];
    drop8(drop, i, cols8, m);                       // the mask of the forward, regenerated
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = unpack_bf16x2(wa[k]), x = unpack_bf16x2(wb[k]);
      o[k] = pack_bf16x2(g.x * m[2 * k] * gelu_grad(x.x), g.y * m[2 * k + 1] * gelu_grad(x.y));
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
                    float* __restrict__ logits, int H1, int C, DropSpec drop) {
  // MLPLayer (:193-199): dense1 -> GELU -> dropout -> dense2 -> dropout (the second one on the outputs themselves);
  // hidden mask = row 0 of the site, output mask = row 1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = warp; k < C; k += nw) {
    float acc = 0.f;
    for (int j = lane; j < H1; j += 32) acc = fmaf(W2[static_cast<int64_t>(k) * H1 + j], gelu_erf(zc[j]) * drop_factor(drop, 0, j), acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[k] = (acc + b2[k]) * drop_factor(drop, 1, k);
  }
}

// Backward.  Gradients are ACCUMULATED into dW1, db1, dW2, db2; dcls (d) = W1^T dzc + dcls_in.
__global__ void __launch_bounds__(256)
cls_head_bwd_init_kernel(const float* __restrict__ dlogits, const float* __restrict__ dcls_in, float* __restrict__ db2,
                         float* __restrict__ dcls, int d, int C, DropSpec drop) {
  for (int c = threadIdx.x; c < d; c += blockDim.x) dcls[c] = dcls_in ? dcls_in[c] : 0.f;
  if (static_cast<int>(threadIdx.x) < C) db2[threadIdx.x] += dlogits[threadIdx.x] * drop_factor(drop, 1, threadIdx.x);
}

__global__ void __launch_bounds__(kHeadWarps * 32)
cls_head_bwd_rows_kernel(const __nv_bfloat16* __restrict__ cls, const float* __restrict__ W1, const float* __restrict__ W2,
                         const float* __restrict__ zc, const float* __restrict__ dlogits, float* __restrict__ dW1,
                         float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ dcls, int d, int H1, int C, DropSpec drop) {
  extern __shared__ float s_dcls[];   // [d] this CTA's partial W1^T dzc
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_dcls[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, j = blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
  if (j < H1) {
    const float mh = drop_factor(drop, 0, j);                 // hidden-unit mask of the forward, regenerated
    const float z = zc[j], h = gelu_erf(z) * mh;
    float dh = 0.f;
    for (int k = 0; k < C; ++k) {
      const float dl = dlogits[k] * drop_factor(drop, 1, k);  // gradient behind the output dropout
      dh = fmaf(W2[static_cast<int64_t>(k) * H1 + j], dl, dh);
      if (lane == 0) dW2[static_cast<int64_t>(k) * H1 + j] += dl * h;
    }
    const float dz = dh * mh * gelu_grad(z);
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

// ---- bimodal classifier pieces (models_archs.py:76-124): plain Linear on one vector, and the cross attention of
// CrossAttentionLayer (:174-183, nn.MultiheadAttention) evaluated for the ONE query row the model keeps (`x_attn[:, 0, :]`).
__global__ void __launch_bounds__(kHeadWarps * 32)
linear_vec_fwd_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ x, float* __restrict__ y,
                      int rows, int cols) {
  const int lane = threadIdx.x & 31, j = blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
  if (j >= rows) return;
  float acc = 0.f;
  for (int c = lane; c < cols; c += 32) acc = fmaf(__ldg(W + static_cast<int64_t>(j) * cols + c), x[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[j] = acc + (b ? b[j] : 0.f);
}

// dW += dy x^T, db += dy, dx += W^T dy   (dx is ACCUMULATED: zero it or seed it with another path's gradient first)
__global__ void __launch_bounds__(kHeadWarps * 32)
linear_vec_bwd_kernel(const float* __restrict__ W, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dW,
                      float* __restrict__ db, float* __restrict__ dx, int rows, int cols) {
  extern __shared__ float s_dx[];   // [cols]
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s_dx[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, j = blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
  if (j < rows) {
    const float g = dy[j];
    if (lane == 0 && db) db[j] += g;
    for (int c = lane; c < cols; c += 32) {
      const int64_t e = static_cast<int64_t>(j) * cols + c;
      dW[e] += g * x[c];
      atomicAdd(&s_dx[c], __ldg(W + e) * g);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) atomicAdd(dx + c, s_dx[c]);
}

__device__ __forceinline__ float block_reduce(float v, float* scratch, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, scratch[i]) : r + scratch[i];
  return r;
}

// One CTA per head: p = softmax(q0_h . K_h^T * scale) over the n keys, o_h = p V_h.  kv: (n, 2d) bf16 = [K | V] as the
// in-projection GEMM writes it.  p (heads, n) f32 is kept for the backward.
__global__ void __launch_bounds__(256)
cross_cls_attn_fwd_kernel(const float* __restrict__ q0, const __nv_bfloat16* __restrict__ kv, int64_t ld, int n, int d, float scale,
                          float* __restrict__ p_out, float* __restrict__ o) {
  __shared__ float s_q[64], s_red[8], s_o[4][64];
  const int h = blockIdx.x, tid = threadIdx.x;
  if (tid < 64) s_q[tid] = q0[h * 64 + tid] * scale;
  __syncthreads();
  float* p = p_out + static_cast<int64_t>(h) * n;
  float mx = -INFINITY;
  for (int k = tid; k < n; k += blockDim.x) {
    const uint4* kr = reinterpret_cast<const uint4*>(kv + static_cast<int64_t>(k) * ld + h * 64);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = __ldg(kr + c);
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
      const float* q = s_q + c * 8;
      acc += q[0] * a0.x + q[1] * a0.y + q[2] * a1.x + q[3] * a1.y + q[4] * a2.x + q[5] * a2.y + q[6] * a3.x + q[7] * a3.y;
    }
    p[k] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = block_reduce(mx, s_red, true);
  float sum = 0.f;
  for (int k = tid; k < n; k += blockDim.x) {
    const float e = __expf(p[k] - mx);
    p[k] = e;
    sum += e;
  }
  sum = block_reduce(sum, s_red, false);
  const float inv = 1.f / sum;
  for (int k = tid; k < n; k += blockDim.x) p[k] *= inv;
  __syncthreads();
  // o_h[dd] = sum_k p_k V[k][dd]: 64 dims x 4 key partitions
  const int dd = tid & 63, part = tid >> 6;
  float acc = 0.f;
  for (int k = part; k < n; k += 4) acc = fmaf(p[k], __bfloat162float(kv[static_cast<int64_t>(k) * ld + d + h * 64 + dd]), acc);
  s_o[part][dd] = acc;
  __syncthreads();
  if (tid < 64) o[h * 64 + tid] = s_o[0][tid] + s_o[1][tid] + s_o[2][tid] + s_o[3][tid];
}

// Backward of the above: given do (d): dV_k = p_k do_h, dp_k = V_k . do_h, ds_k = p_k (dp_k - sum_j p_j dp_j) scale,
// dq0_h = sum_k ds_k K_k, dK_k = ds_k q0_h.  dkv (n, 2d) bf16 is written, dq0 (d) f32 is written.
__global__ void __launch_bounds__(256)
cross_cls_attn_bwd_kernel(const float* __restrict__ q0, const __nv_bfloat16* __restrict__ kv, int64_t ld, const float* __restrict__ p_in,
                          const float* __restrict__ d_o, int n, int d, float scale, float* __restrict__ dq0,
                          __nv_bfloat16* __restrict__ dkv, int64_t ld_dkv, float* __restrict__ dp_scratch) {
  __shared__ float s_q[64], s_do[64], s_red[8], s_dq[4][64];
  const int h = blockIdx.x, tid = threadIdx.x;
  if (tid < 64) { s_q[tid] = q0[h * 64 + tid]; s_do[tid] = d_o[h * 64 + tid]; }
  __syncthreads();
  const float* p = p_in + static_cast<int64_t>(h) * n;
  float* dp = dp_scratch + static_cast<int64_t>(h) * n;
  float c_acc = 0.f;
  for (int k = tid; k < n; k += blockDim.x) {
    const uint4* vr = reinterpret_cast<const uint4*>(kv + static_cast<int64_t>(k) * ld + d + h * 64);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = __ldg(vr + c);
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
      const float* g = s_do + c * 8;
      acc += g[0] * a0.x + g[1] * a0.y + g[2] * a1.x + g[3] * a1.y + g[4] * a2.x + g[5] * a2.y + g[6] * a3.x + g[7] * a3.y;
    }
    dp[k] = acc;
    c_acc = fmaf(p[k], acc, c_acc);
  }
  c_acc = block_reduce(c_acc, s_red, false);
  // per key: ds, dK row, dV row (each thread writes the two 128-byte rows of its keys)
  for (int k = tid; k < n; k += blockDim.x) {
    const float pk = p[k], ds = pk * (dp[k] - c_acc) * scale;
    dp[k] = ds;
    __nv_bfloat16* dk = dkv + static_cast<int64_t>(k) * ld_dkv + h * 64;
    __nv_bfloat16* dv = dkv + static_cast<int64_t>(k) * ld_dkv + d + h * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      uint4 a, b;
      a.x = pack_bf16x2(ds * s_q[c], ds * s_q[c + 1]); a.y = pack_bf16x2(ds * s_q[c + 2], ds * s_q[c + 3]);
      a.z = pack_bf16x2(ds * s_q[c + 4], ds * s_q[c + 5]); a.w = pack_bf16x2(ds * s_q[c + 6], ds * s_q[c + 7]);
      b.x = pack_bf16x2(pk * s_do[c], pk * s_do[c + 1]); b.y = pack_bf16x2(pk * s_do[c + 2], pk * s_do[c + 3]);
      b.z = pack_bf16x2(pk * s_do[c + 4], pk * s_do[c + 5]); b.w = pack_bf16x2(pk * s_do[c + 6], pk * s_do[c + 7]);
      *reinterpret_cast<uint4*>(dk + c) = a;
      *reinterpret_cast<uint4*>(dv + c) = b;
    }
  }
  __syncthreads();
  const int dd = tid & 63, part = tid >> 6;
  float acc = 0.f;
  for (int k = part; k < n; k += 4) acc = fmaf(dp[k], __bfloat162float(kv[static_cast<int64_t>(k) * ld + h * 64 + dd]), acc);
  s_dq[part][dd] = acc;
  __syncthreads();
  if (tid < 64) dq0[h * 64 + tid] = s_dq[0][tid] + s_dq[1][tid] + s_dq[2][tid] + s_dq[3][tid];
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

static int check_drop(const char* who, const vdr_dropout* d, int64_t n, int cols) {
  VDR_CHECK_ARG(d == nullptr || d->thr16 < 65536u, VDR_EINVAL, "%s: dropout threshold must be < 65536 (p < 1)", who);
  VDR_CHECK_ARG(d == nullptr || d->thr16 == 0 || (cols > 0 && cols % 8 == 0 && n % cols == 0), VDR_EINVAL,
                "%s: dropout needs the matrix width (cols %% 8 == 0, n %% cols == 0), got cols = %d", who, cols);
  return VDR_OK;
}
static vdr::DropSpec drop_spec(const vdr_dropout* d) {
  vdr::DropSpec s{0ull, 0u, 0u, nullptr};
  if (d != nullptr) { s.seed = d->seed; s.site = d->site; s.thr16 = d->thr16; s.seed_offset = reinterpret_cast<const unsigned long long*>(d->seed_offset); }
  return s;
}

extern "C" int vdr_gelu_fwd(const void* z, void* h, int64_t n, int cols, const vdr_dropout* drop, vdr_stream_t stream) {
  VDR_CHECK_ARG(z && h && n > 0 && n % 8 == 0, VDR_EINVAL, "vdr_gelu_fwd: null pointer or n %% 8 != 0");
  VDR_CHECK_ARG(aligned16(z) && aligned16(h), VDR_EALIGN, "vdr_gelu_fwd: pointers must be 16-byte aligned");
  if (int rc = check_drop("vdr_gelu_fwd", drop, n, cols)) return rc;
  gelu_fwd_kernel<<<ew_grid(n / 8), 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(z), static_cast<__nv_bfloat16*>(h), n / 8,
                                                          cols > 0 ? cols / 8 : 1, drop_spec(drop));
  count_launch();
  VDR_CHECK_LAUNCH("gelu_fwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_gelu_bwd(const void* dh, const void* z, void* dz, int64_t n, int cols, const vdr_dropout* drop, vdr_stream_t stream) {
  VDR_CHECK_ARG(dh && z && dz && n > 0 && n % 8 == 0, VDR_EINVAL, "vdr_gelu_bwd: null pointer or n %% 8 != 0");
  VDR_CHECK_ARG(aligned16(dh) && aligned16(z) && aligned16(dz), VDR_EALIGN, "vdr_gelu_bwd: pointers must be 16-byte aligned");
  if (int rc = check_drop("vdr_gelu_bwd", drop, n, cols)) return rc;
  gelu_bwd_kernel<<<ew_grid(n / 8), 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(dh), static_cast<const __nv_bfloat16*>(z),
                                                          static_cast<__nv_bfloat16*>(dz), n / 8, cols > 0 ? cols / 8 : 1, drop_spec(drop));
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

extern "C" int vdr_dropout_apply(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int cols, const vdr_dropout* drop,
                                 vdr_stream_t stream) {
  VDR_CHECK_ARG(x && out && drop && rows > 0 && cols > 0 && cols % 8 == 0, VDR_EINVAL, "vdr_dropout_apply: null pointer or cols %% 8 != 0");
  VDR_CHECK_ARG(aligned16(x) && aligned16(out) && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= cols && ldo >= cols, VDR_EALIGN,
                "vdr_dropout_apply: pointers must be 16-byte aligned, pitches multiples of 8 and >= cols");
  VDR_CHECK_ARG(drop->thr16 < 65536u, VDR_EINVAL, "vdr_dropout_apply: dropout threshold must be < 65536");
  dropout_apply_kernel<<<ew_grid(rows * (cols / 8)), 256, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out),
                                                                            ldo, rows, cols / 8, drop_spec(drop));
  count_launch();
  VDR_CHECK_LAUNCH("dropout_apply_kernel");
  return VDR_OK;
}

extern "C" int vdr_dropout_mask(uint8_t* out, int64_t rows, int64_t cols, const vdr_dropout* drop, vdr_stream_t stream) {
  VDR_CHECK_ARG(out && drop && rows > 0 && cols > 0 && cols < (1ll << 35), VDR_EINVAL, "vdr_dropout_mask: bad arguments");
  VDR_CHECK_ARG(drop->thr16 < 65536u, VDR_EINVAL, "vdr_dropout_mask: dropout threshold must be < 65536");
  dropout_mask_kernel<<<ew_grid(rows * cols), 256, 0, S_(stream)>>>(out, rows, cols, drop_spec(drop));
  count_launch();
  VDR_CHECK_LAUNCH("dropout_mask_kernel");
  return VDR_OK;
}

extern "C" int vdr_cls_head_fwd(const void* cls_bf16, const float* W1, const float* b1, const float* W2, const float* b2,
                                float* zc, float* logits, int d, int H1, int C, const vdr_dropout* drop, vdr_stream_t stream) {
  VDR_CHECK_ARG(cls_bf16 && W1 && b1 && W2 && b2 && zc && logits, VDR_EINVAL, "vdr_cls_head_fwd: null pointer");
  VDR_CHECK_ARG(d > 0 && H1 > 0 && C > 0, VDR_EINVAL, "vdr_cls_head_fwd: bad shape");
  VDR_CHECK_ARG(drop == nullptr || drop->thr16 < 65536u, VDR_EINVAL, "vdr_cls_head_fwd: dropout threshold must be < 65536");
  cls_head_hidden_kernel<<<(H1 + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, 0, S_(stream)>>>(static_cast<const __nv_bfloat16*>(cls_bf16), W1, b1, zc, d, H1);
  cls_head_out_kernel<<<1, 256, 0, S_(stream)>>>(zc, W2, b2, logits, H1, C, drop_spec(drop));
  count_launch(2);
  VDR_CHECK_LAUNCH("cls_head_fwd kernels");
  return VDR_OK;
}

extern "C" int vdr_cls_head_bwd(const void* cls_bf16, const float* W1, const float* W2, const float* zc, const float* dlogits,
                                const float* dcls_in, float* dW1, float* db1, float* dW2, float* db2, float* dcls, int d,
                                int H1, int C, const vdr_dropout* drop, vdr_stream_t stream) {
  VDR_CHECK_ARG(cls_bf16 && W1 && W2 && zc && dlogits && dW1 && db1 && dW2 && db2 && dcls, VDR_EINVAL, "vdr_cls_head_bwd: null pointer");
  VDR_CHECK_ARG(d > 0 && H1 > 0 && C > 0 && C <= 256 && (size_t)d * 4 <= 48 * 1024, VDR_EINVAL, "vdr_cls_head_bwd: bad shape");
  VDR_CHECK_ARG(drop == nullptr || drop->thr16 < 65536u, VDR_EINVAL, "vdr_cls_head_bwd: dropout threshold must be < 65536");
  cls_head_bwd_init_kernel<<<1, 256, 0, S_(stream)>>>(dlogits, dcls_in, db2, dcls, d, C, drop_spec(drop));
  cls_head_bwd_rows_kernel<<<(H1 + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, d * sizeof(float), S_(stream)>>>(
      static_cast<const __nv_bfloat16*>(cls_bf16), W1, W2, zc, dlogits, dW1, db1, dW2, dcls, d, H1, C, drop_spec(drop));
  count_launch(2);
  VDR_CHECK_LAUNCH("cls_head_bwd kernels");
  return VDR_OK;
}

extern "C" int vdr_linear_vec_fwd(const float* W, const float* b, const float* x, float* y, int rows, int cols, vdr_stream_t stream) {
  VDR_CHECK_ARG(W && x && y && rows > 0 && cols > 0, VDR_EINVAL, "vdr_linear_vec_fwd: bad arguments");
  linear_vec_fwd_kernel<<<(rows + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, 0, S_(stream)>>>(W, b, x, y, rows, cols);
  count_launch();
  VDR_CHECK_LAUNCH("linear_vec_fwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_linear_vec_bwd(const float* W, const float* x, const float* dy, float* dW, float* db, float* dx, int rows, int cols,
                                  vdr_stream_t stream) {
  VDR_CHECK_ARG(W && x && dy && dW && dx && rows > 0 && cols > 0 && (size_t)cols * 4 <= 48 * 1024, VDR_EINVAL, "vdr_linear_vec_bwd: bad arguments");
  linear_vec_bwd_kernel<<<(rows + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, cols * sizeof(float), S_(stream)>>>(W, x, dy, dW, db, dx, rows, cols);
  count_launch();
  VDR_CHECK_LAUNCH("linear_vec_bwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_cross_cls_attn_fwd(const float* q0, const void* kv, int64_t ld_kv, int n, int heads, float scale, float* p, float* o,
                                      vdr_stream_t stream) {
  VDR_CHECK_ARG(q0 && kv && p && o && n > 0 && heads > 0 && ld_kv >= 2 * heads * 64 && ld_kv % 8 == 0 && aligned16(kv), VDR_EINVAL,
                "vdr_cross_cls_attn_fwd: bad arguments");
  cross_cls_attn_fwd_kernel<<<heads, 256, 0, S_(stream)>>>(q0, static_cast<const __nv_bfloat16*>(kv), ld_kv, n, heads * 64, scale, p, o);
  count_launch();
  VDR_CHECK_LAUNCH("cross_cls_attn_fwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_cross_cls_attn_bwd(const float* q0, const void* kv, int64_t ld_kv, const float* p, const float* d_o, int n, int heads,
                                      float scale, float* dq0, void* dkv, int64_t ld_dkv, float* scratch, vdr_stream_t stream) {
  VDR_CHECK_ARG(q0 && kv && p && d_o && dq0 && dkv && scratch && n > 0 && heads > 0, VDR_EINVAL, "vdr_cross_cls_attn_bwd: bad arguments");
  VDR_CHECK_ARG(ld_kv >= 2 * heads * 64 && ld_kv % 8 == 0 && ld_dkv >= 2 * heads * 64 && ld_dkv % 8 == 0 && aligned16(kv) && aligned16(dkv), VDR_EALIGN,
                "vdr_cross_cls_attn_bwd: kv / dkv must be 16-byte aligned with leading dimensions that are multiples of 8");
  cross_cls_attn_bwd_kernel<<<heads, 256, 0, S_(stream)>>>(q0, static_cast<const __nv_bfloat16*>(kv), ld_kv, p, d_o, n, heads * 64, scale, dq0,
                                                           static_cast<__nv_bfloat16*>(dkv), ld_dkv, scratch);
  count_launch();
  VDR_CHECK_LAUNCH("cross_cls_attn_bwd_kernel");
  return VDR_OK;
}
