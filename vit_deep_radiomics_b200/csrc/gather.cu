// Tumour-mask gathers.
//   G1  vdr_mask_gather : stable stream compaction of in-mask ViT tokens into a point cloud
//       (warp ballot + popc ranks inside a tile, tile counts -> exclusive scan -> scatter),
//       rows copied warp-wide with 16-byte accesses, optional fp64 3-D sinusoidal PE epilogue.
//   G2  vdr_voxel_bbox / vdr_voxel_gather : bounding box of the mask, then every voxel inside it.
// Index contracts are in SURVEY.md Appendix A1/A2/A6 and restated in oracle/gather_np.py.
#include "common.cuh"

namespace vdr {

constexpr int kTile = 2048;      // candidates per block
constexpr int kGThreads = 256;   // 8 warps
constexpr int kIters = kTile / kGThreads;

struct G1Geom {
  const uint8_t* mask;
  const int32_t* row_map;
  const int32_t* col_map;
  int S, h, w;
  int64_t mask_slice_stride, mask_row_stride, mask_col_stride;   // bytes between slices / rows / columns of the pixel mask
  int64_t feat_slice_rows, feat_row_pitch, feat_row0;  // token row = k*slice_rows + row0 + a*pitch + b
  int64_t total;
};

// candidate n = a*(w*S) + b*S + k  (slice fastest)  ->  resized-mask value
__device__ __forceinline__ bool g1_pred(const G1Geom& g, int64_t n) {
  const int k = static_cast<int>(n % g.S);
  const int64_t q = n / g.S;
  const int b = static_cast<int>(q % g.w);
  const int a = static_cast<int>(q / g.w);
  const int64_t off = static_cast<int64_t>(k) * g.mask_slice_stride + __ldg(g.row_map + a) * g.mask_row_stride + __ldg(g.col_map + b) * g.mask_col_stride;
  return __ldg(g.mask + off) != 0;
}

__global__ void __launch_bounds__(kGThreads) g1_count_kernel(G1Geom g, int32_t* __restrict__ tile_counts) {
  __shared__ int s_warp[kGThreads / 32];
  const int64_t tile_base = static_cast<int64_t>(blockIdx.x) * kTile;
  int cnt = 0;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int64_t n = tile_base + it * kGThreads + threadIdx.x;
    const bool p = (n < g.total) && g1_pred(g, n);
    cnt += __popc(__ballot_sync(0xffffffffu, p));
  }
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;  // every lane holds the warp total
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kGThreads / 32; ++i) t += s_warp[i];
    tile_counts[blockIdx.x] = t;
  }
}

// Single-block exclusive scan of the tile counts; writes the grand total to out_count.
__global__ void __launch_bounds__(1024) tile_scan_kernel(const int32_t* __restrict__ counts,
                                                         int32_t* __restrict__ offsets, int num_tiles,
                                                         int32_t* __restrict__ out_count) {
  __shared__ int s_warp[32];
  const int per = (num_tiles + 1023) / 1024;
  const int begin = threadIdx.x * per;
  int end = begin + per;
  if (end > num_tiles) end = num_tiles;
  int sum = 0;
  for (int i = begin; i < end; ++i) sum += counts[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - w;  // exclusive warp offsets
  }
  __syncthreads();
  int run = s_warp[warp] + incl - sum;
  for (int i = begin; i < end; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
  if (threadIdx.x == 1023) *out_count = run;
}

struct G1Pe {
  double scale;          // 0 = no positional encoding
  const double* div;     // [D/6] divisors 10000^(6 i / D), host-computed in f64
  double* table;         // [(h + w + S)][2*(D/6)] f64: scale * sin|cos(coord / div) per distinct coordinate (workspace)
  double w_orig, h_orig, res0, res1, res2, noise0, noise1, noise2, mean_x, mean_y, mean_z;
};

// The encoding of a token depends on its three grid indices separately (xi in [0,h), yi in [0,w), zi in [0,S) under
// the reference's meshgrid quirk), so the h + w + S distinct coordinate rows are evaluated once in f64 (same operation
// order as train_models.py:166-176 and :34-44) and the emit kernel only adds table entries: (h+w+S) * D/3 sin/cos
// instead of n_sel * D.
__global__ void __launch_bounds__(256) g1_pe_table_kernel(G1Pe pe, int S, int h, int w, int npair2) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(h + w + S) * npair2;
  if (idx >= total) return;
  const int e = static_cast<int>(idx / npair2), jj = static_cast<int>(idx % npair2);
  double v;
  if (e < h) v = __dadd_rn(__dsub_rn(__dmul_rn(__dmul_rn(static_cast<double>(e) / static_cast<double>(w), pe.w_orig), pe.res0), pe.mean_x), pe.noise0);
  else if (e < h + w) v = __dadd_rn(__dsub_rn(__dmul_rn(__dmul_rn(static_cast<double>(e - h) / static_cast<double>(h), pe.h_orig), pe.res1), pe.mean_y), pe.noise1);
  else v = __dadd_rn(__dsub_rn(__dmul_rn(static_cast<double>(e - h - w), pe.res2), pe.mean_z), pe.noise2);
  const double arg = v / pe.div[jj >> 1];
  const double enc = (jj & 1) ? cos(arg) : sin(arg);
  pe.table[idx] = __dmul_rn(enc, pe.scale);
}

__global__ void __launch_bounds__(kGThreads)
g1_scatter_kernel(G1Geom g, const int32_t* __restrict__ tile_offsets, int32_t* __restrict__ out_src, int cap) {
  __shared__ int s_cnt[kIters * 8];
  __shared__ int s_off[kIters * 8 + 1];
  __shared__ int s_sel[kTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile_base = static_cast<int64_t>(blockIdx.x) * kTile;
  uint32_t ballots[kIters];
  uint32_t mine = 0;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int64_t n = tile_base + it * kGThreads + threadIdx.x;
    const bool p = (n < g.total) && g1_pred(g, n);
    ballots[it] = __ballot_sync(0xffffffffu, p);
    mine |= (p ? 1u : 0u) << it;
    if (lane == 0) s_cnt[it * 8 + warp] = __popc(ballots[it]);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 64 (iteration, warp) counts, two per lane
    const int v0 = s_cnt[lane * 2], v1 = s_cnt[lane * 2 + 1];
    const int sum = v0 + v1;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    s_off[lane * 2] = incl - sum;
    s_off[lane * 2 + 1] = incl - sum + v0;
    if (lane == 31) s_off[kIters * 8] = incl;
  }
  __syncthreads();
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kIters; ++it)
    if ((mine >> it) & 1u) {
      const int pos = s_off[it * 8 + warp] + __popc(ballots[it] & lt_mask);
      s_sel[pos] = it * kGThreads + threadIdx.x;
    }
  __syncthreads();
  const int tile_n = s_off[kIters * 8];
  const int64_t tile_off = tile_offsets[blockIdx.x];
  // write the (slice, row, col) of every selected candidate of this tile, in order
  for (int j = threadIdx.x; j < tile_n; j += kGThreads) {
    const int64_t row_out = tile_off + j;
    if (row_out >= cap) break;
    const int64_t n = tile_base + s_sel[j];
    const int k = static_cast<int>(n % g.S);
    const int64_t q = n / g.S;
    out_src[row_out * 3 + 0] = k;
    out_src[row_out * 3 + 1] = static_cast<int>(q / g.w);
    out_src[row_out * 3 + 2] = static_cast<int>(q % g.w);
  }
}

// One warp per OUTPUT row (grid-stride over the device-side count): copies the descriptor row and adds the
// positional encoding from the per-coordinate f64 table (pure memory traffic: row in, row out, 3 L2-resident table rows).
template <bool FEAT_BF16>
__global__ void __launch_bounds__(kGThreads)
g1_emit_kernel(G1Geom g, const void* __restrict__ feat, int64_t ld_feat, int D, const int32_t* __restrict__ out_src,
               const int32_t* __restrict__ out_count, float* __restrict__ out_tok, int cap, G1Pe pe, bool vec_ok) {
  const int lane = threadIdx.x & 31;
  int64_t count = *out_count;
  if (count > cap) count = cap;
  const int third = D / 3, two_third = (2 * D) / 3, npair2 = 2 * (D / 6);
  for (int64_t row_out = blockIdx.x * (int64_t)(kGThreads / 32) + (threadIdx.x >> 5); row_out < count;
       row_out += (int64_t)gridDim.x * (kGThreads / 32)) {
    const int k = out_src[row_out * 3 + 0], a = out_src[row_out * 3 + 1], b = out_src[row_out * 3 + 2];
    const int64_t n = (static_cast<int64_t>(a) * g.w + b) * g.S + k;
    // reference meshgrid(indexing='xy') quirk: xi = (n / S) % h, yi = n / (h*S), zi = n % S
    const double* tx = pe.table + ((n / g.S) % g.h) * npair2;
    const double* ty = pe.table + (g.h + n / (static_cast<int64_t>(g.h) * g.S)) * npair2;
    const double* tz = pe.table + (g.h + g.w + k) * npair2;
    const int64_t src_row = static_cast<int64_t>(k) * g.feat_slice_rows + g.feat_row0 + a * g.feat_row_pitch + b;
    auto pe_add = [&](float f, int col) -> float {
      int jj;
      const double* t;
      if (col < third) { jj = col; t = tx; }
      else if (col < two_third) { jj = col - third; t = ty; }
      else { jj = col - two_third; t = tz; }
      if (jj < npair2) return static_cast<float>(__dadd_rn(static_cast<double>(f), __ldg(t + jj)));
      return f;
    };
    if (vec_ok) {
      // D % 24 == 0 (every 8-column chunk lies inside one axis block and all of its columns are encoded): the chunk's
      // eight table entries are four 16-byte loads issued together with the descriptor loads -- one memory round trip.
      const bool pe_vec = pe.scale != 0. && D % 24 == 0;
      for (int c0 = lane * 8; c0 < D; c0 += 256) {
        float f[8];
        double2 t4[4];
        if (pe_vec) {
          const double* t = (c0 < third) ? tx + c0 : (c0 < two_third) ? ty + (c0 - third) : tz + (c0 - two_third);
#pragma unroll
          for (int i = 0; i < 4; ++i) t4[i] = __ldg(reinterpret_cast<const double2*>(t) + i);
        }
        if (FEAT_BF16) {
          const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(feat) + src_row * ld_feat + c0);
          const float2 p0 = unpack_bf16x2(v.x), p1 = unpack_bf16x2(v.y), p2 = unpack_bf16x2(v.z), p3 = unpack_bf16x2(v.w);
          f[0] = p0.x; f[1] = p0.y; f[2] = p1.x; f[3] = p1.y; f[4] = p2.x; f[5] = p2.y; f[6] = p3.x; f[7] = p3.y;
        } else {
          const float* fp = static_cast<const float*>(feat) + src_row * ld_feat + c0;
          const float4 u0 = *reinterpret_cast<const float4*>(fp), u1 = *reinterpret_cast<const float4*>(fp + 4);
          f[0] = u0.x; f[1] = u0.y; f[2] = u0.z; f[3] = u0.w; f[4] = u1.x; f[5] = u1.y; f[6] = u1.z; f[7] = u1.w;
        }
        if (pe_vec) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[2 * i] = static_cast<float>(__dadd_rn(static_cast<double>(f[2 * i]), t4[i].x));
            f[2 * i + 1] = static_cast<float>(__dadd_rn(static_cast<double>(f[2 * i + 1]), t4[i].y));
          }
        } else if (pe.scale != 0.) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = pe_add(f[i], c0 + i);
        }
        float* op = out_tok + row_out * D + c0;
        __stcs(reinterpret_cast<float4*>(op), make_float4(f[0], f[1], f[2], f[3]));
        __stcs(reinterpret_cast<float4*>(op + 4), make_float4(f[4], f[5], f[6], f[7]));
      }
    } else {  // any D / alignment: one element per lane
      for (int c = lane; c < D; c += 32) {
        float f = FEAT_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(feat)[src_row * ld_feat + c])
                            : static_cast<const float*>(feat)[src_row * ld_feat + c];
        if (pe.scale != 0.) f = pe_add(f, c);
        out_tok[row_out * D + c] = f;
      }
    }
  }
}

// ----------------------------------------------------------------------------- G2
__global__ void bbox_init_kernel(int32_t* bbox) {
  if (threadIdx.x < 6) bbox[threadIdx.x] = (threadIdx.x & 1) ? -1 : 0x7fffffff;
}

// kRowCol = false: the reference's xy-meshgrid convention (xi = (n/S) % H, yi = n/(H*S));  true: (col, row) of voxel n,
// i.e. the bounding box of the union mask over slices that generate_features starts from (tfds_dense_descriptor.py:257-260).
template <bool kRowCol>
__global__ void __launch_bounds__(256) voxel_bbox_kernel(const uint8_t* __restrict__ mask, int H, int W, int S,
                                                         int64_t total, int32_t* __restrict__ bbox) {
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
  // 16 mask bytes per thread per step
  const int64_t nvec = total >> 4;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvec + 1; v += (int64_t)gridDim.x * blockDim.x) {
    uint32_t wds[4] = {0, 0, 0, 0};
    int cnt = 16;
    if (v < nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(mask) + v);
      wds[0] = u.x; wds[1] = u.y; wds[2] = u.z; wds[3] = u.w;
    } else {  // tail
      cnt = static_cast<int>(total - (nvec << 4));
      for (int i = 0; i < cnt; ++i) wds[i >> 2] |= static_cast<uint32_t>(mask[(nvec << 4) + i]) << ((i & 3) * 8);
    }
    if ((wds[0] | wds[1] | wds[2] | wds[3]) == 0) continue;
    for (int i = 0; i < cnt; ++i) {
      if ((wds[i >> 2] >> ((i & 3) * 8)) & 0xffu) {
        const int64_t n = (v << 4) + i;
        const int zi = static_cast<int>(n % S);
        const int64_t q = n / S;
        const int xi = kRowCol ? static_cast<int>(q % W) : static_cast<int>(q % H);
        const int yi = kRowCol ? static_cast<int>(q / W) : static_cast<int>(q / H);
        lo[0] = min(lo[0], xi); hi[0] = max(hi[0], xi);
        lo[1] = min(lo[1], yi); hi[1] = max(hi[1], yi);
        lo[2] = min(lo[2], zi); hi[2] = max(hi[2], zi);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0 && hi[a] >= 0) {
      atomicMin(bbox + 2 * a, lo[a]);
      atomicMax(bbox + 2 * a + 1, hi[a]);
    }
  }
}

// Voxels inside the box, ascending flat index n = (yi*H + xi)*S + zi: a box in (yi, xi, zi) is
// lexicographically ordered exactly like n, so output row j maps to its voxel in closed form.
__global__ void __launch_bounds__(256)
voxel_gather_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, int H, int S,
                    const int32_t* __restrict__ bbox, int32_t* __restrict__ out_flat, float* __restrict__ out_raw,
                    uint8_t* __restrict__ out_mask, int32_t* __restrict__ out_count, int cap) {
  const int x0 = bbox[0], x1 = bbox[1], y0 = bbox[2], y1 = bbox[3], z0 = bbox[4], z1 = bbox[5];
  const int64_t ex = (x1 >= x0) ? (x1 - x0 + 1) : 0, ey = (y1 >= y0) ? (y1 - y0 + 1) : 0, ez = (z1 >= z0) ? (z1 - z0 + 1) : 0;
  const int64_t count = ex * ey * ez;
  if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = static_cast<int32_t>(count);
  const int64_t lim = count < cap ? count : cap;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < lim; j += (int64_t)gridDim.x * blockDim.x) {
    const int zi = z0 + static_cast<int>(j % ez);
    const int64_t t = j / ez;
    const int xi = x0 + static_cast<int>(t % ex);
    const int yi = y0 + static_cast<int>(t / ex);
    const int64_t n = (static_cast<int64_t>(yi) * H + xi) * S + zi;
    out_flat[j] = static_cast<int32_t>(n);
    out_raw[j] = __ldg(img + n);
    out_mask[j] = __ldg(mask + n);
  }
}

}  // namespace vdr

static size_t g1_scan_bytes(int S, int h, int w) {   // tile counts + offsets, rounded up to 16 bytes
  const int64_t total = (int64_t)S * h * w;
  const int64_t tiles = (total + vdr::kTile - 1) / vdr::kTile;
  return (((size_t)(tiles > 0 ? tiles : 1) * 2 * sizeof(int32_t)) + 15) & ~(size_t)15;
}

extern "C" size_t vdr_mask_gather_workspace_bytes(int S, int h, int w, int D) {
  return g1_scan_bytes(S, h, w) + (size_t)(h + w + S) * (size_t)(2 * (D / 6)) * sizeof(double);
}

extern "C" int vdr_mask_gather(const void* feat, int feat_dtype, int64_t ld_feat, int64_t feat_slice_rows,
                               int64_t feat_row_pitch, int64_t feat_row0, const uint8_t* mask,
                               int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                               const int32_t* row_map, const int32_t* col_map, int S, int h, int w, int D,
                               float* out_tok, int32_t* out_src, int32_t* out_count, int cap, double pe_scale,
                               const double* pe_div, const double* coef_host, void* workspace,
                               size_t workspace_bytes, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(feat && mask && row_map && col_map && out_tok && out_src && out_count && workspace, VDR_EINVAL, "vdr_mask_gather: null pointer");
  VDR_CHECK_ARG(S > 0 && h > 0 && w > 0 && D > 0 && cap >= 0, VDR_EINVAL, "vdr_mask_gather: bad shape");
  VDR_CHECK_ARG(feat_slice_rows >= 0 && feat_row_pitch >= w && feat_row0 >= 0 && mask_slice_stride >= 0 && mask_row_stride >= 0 && mask_col_stride >= 1, VDR_EINVAL, "vdr_mask_gather: bad strides");
  VDR_CHECK_ARG(ld_feat >= D, VDR_EINVAL, "vdr_mask_gather: ld_feat (%lld) smaller than D (%d)", (long long)ld_feat, D);
  // 16-byte vector path when the rows allow it; otherwise an element-wise path (any D, e.g. the reference's D = 12 tests)
  const bool vec_ok = D % 8 == 0 && ld_feat % 8 == 0 && aligned16(feat) && aligned16(out_tok);
  VDR_CHECK_ARG(feat_dtype == VDR_DTYPE_BF16 || feat_dtype == VDR_DTYPE_F32, VDR_EINVAL, "vdr_mask_gather: bad feat_dtype");
  VDR_CHECK_ARG((int64_t)S * h * w < 0x7fffffffLL, VDR_EINVAL, "vdr_mask_gather: too many candidates");
  VDR_CHECK_ARG(workspace_bytes >= vdr_mask_gather_workspace_bytes(S, h, w, D), VDR_EWORKSPACE, "vdr_mask_gather: workspace too small (%zu < %zu)", workspace_bytes, vdr_mask_gather_workspace_bytes(S, h, w, D));
  VDR_CHECK_ARG(aligned16(workspace), VDR_EALIGN, "vdr_mask_gather: workspace must be 16-byte aligned");
  VDR_CHECK_ARG(pe_scale == 0.0 || (pe_div && coef_host), VDR_EINVAL, "vdr_mask_gather: positional encoding needs pe_div and coef_host");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  G1Geom g{mask, row_map, col_map, S, h, w, mask_slice_stride, mask_row_stride, mask_col_stride, feat_slice_rows, feat_row_pitch, feat_row0, (int64_t)S * h * w};
  const int tiles = (int)((g.total + kTile - 1) / kTile);
  int32_t* counts = static_cast<int32_t*>(workspace);
  int32_t* offsets = counts + tiles;
  G1Pe pe{};
  pe.scale = pe_scale;
  pe.div = pe_div;
  pe.table = reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + g1_scan_bytes(S, h, w));
  const int npair2 = 2 * (D / 6);
  if (pe_scale != 0.0) {
    pe.w_orig = coef_host[0]; pe.h_orig = coef_host[1];
    pe.res0 = coef_host[2]; pe.res1 = coef_host[3]; pe.res2 = coef_host[4];
    pe.noise0 = coef_host[5]; pe.noise1 = coef_host[6]; pe.noise2 = coef_host[7];
    pe.mean_x = coef_host[8]; pe.mean_y = coef_host[9]; pe.mean_z = coef_host[10];
  }
  if (pe_scale != 0.0 && npair2 > 0) {
    const int64_t entries = (int64_t)(h + w + S) * npair2;
    g1_pe_table_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, s>>>(pe, S, h, w, npair2);
    count_launch();
    VDR_CHECK_LAUNCH("g1_pe_table_kernel");
  }
  g1_count_kernel<<<tiles, kGThreads, 0, s>>>(g, counts);
  VDR_CHECK_LAUNCH("g1_count_kernel");
  tile_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, tiles, out_count);
  VDR_CHECK_LAUNCH("tile_scan_kernel");
  g1_scatter_kernel<<<tiles, kGThreads, 0, s>>>(g, offsets, out_src, cap);
  VDR_CHECK_LAUNCH("g1_scatter_kernel");
  int emit_blocks = (cap + (kGThreads / 32) - 1) / (kGThreads / 32);
  if (emit_blocks > num_sms() * 8) emit_blocks = num_sms() * 8;
  if (emit_blocks < 1) emit_blocks = 1;
  if (feat_dtype == VDR_DTYPE_BF16)
    g1_emit_kernel<true><<<emit_blocks, kGThreads, 0, s>>>(g, feat, ld_feat, D, out_src, out_count, out_tok, cap, pe, vec_ok);
  else
    g1_emit_kernel<false><<<emit_blocks, kGThreads, 0, s>>>(g, feat, ld_feat, D, out_src, out_count, out_tok, cap, pe, vec_ok);
  count_launch(4);
  VDR_CHECK_LAUNCH("g1_emit_kernel");
  return VDR_OK;
}

extern "C" int vdr_voxel_bbox(const uint8_t* mask, int H, int W, int S, int32_t* bbox, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(mask && bbox, VDR_EINVAL, "vdr_voxel_bbox: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0, VDR_EINVAL, "vdr_voxel_bbox: bad shape");
  VDR_CHECK_ARG((int64_t)H * W * S < 0x7fffffffLL, VDR_EINVAL, "vdr_voxel_bbox: volume too large for int32 flat indices");
  VDR_CHECK_ARG(aligned16(mask), VDR_EALIGN, "vdr_voxel_bbox: mask must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)H * W * S;
  bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
  int64_t blocks = ((total >> 4) + 1 + 255) / 256;
  const int64_t capb = (int64_t)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  voxel_bbox_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(mask, H, W, S, total, bbox);
  count_launch(2);
  VDR_CHECK_LAUNCH("voxel_bbox_kernel");
  return VDR_OK;
}

extern "C" int vdr_mask_bbox(const uint8_t* mask, int H, int W, int S, int32_t* bbox, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(mask && bbox, VDR_EINVAL, "vdr_mask_bbox: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0, VDR_EINVAL, "vdr_mask_bbox: bad shape");
  VDR_CHECK_ARG((int64_t)H * W * S < 0x7fffffffLL, VDR_EINVAL, "vdr_mask_bbox: volume too large for int32 flat indices");
  VDR_CHECK_ARG(aligned16(mask), VDR_EALIGN, "vdr_mask_bbox: mask must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)H * W * S;
  bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
  int64_t blocks = ((total >> 4) + 1 + 255) / 256;
  const int64_t capb = (int64_t)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  voxel_bbox_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(mask, H, W, S, total, bbox);
  count_launch(2);
  VDR_CHECK_LAUNCH("voxel_bbox_kernel");
  return VDR_OK;
}

extern "C" int vdr_voxel_gather(const float* img, const uint8_t* mask, int H, int W, int S, const int32_t* bbox,
                                int32_t* out_flat, float* out_raw, uint8_t* out_mask, int32_t* out_count, int cap,
                                vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(img && mask && bbox && out_flat && out_raw && out_mask && out_count, VDR_EINVAL, "vdr_voxel_gather: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0 && cap >= 0, VDR_EINVAL, "vdr_voxel_gather: bad shape");
  VDR_CHECK_ARG((int64_t)H * W * S < 0x7fffffffLL, VDR_EINVAL, "vdr_voxel_gather: volume too large for int32 flat indices");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = ((int64_t)cap + 255) / 256;
  const int64_t capb = (int64_t)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  if (blocks < 1) blocks = 1;
  voxel_gather_kernel<<<(unsigned)blocks, 256, 0, s>>>(img, mask, H, S, bbox, out_flat, out_raw, out_mask, out_count, cap);
  count_launch();
  VDR_CHECK_LAUNCH("voxel_gather_kernel");
  return VDR_OK;
}
