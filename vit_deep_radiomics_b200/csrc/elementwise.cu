// Patch extraction (im2col fused with prepare_image's gray2rgb / NCHW / cast) and CLS rows.
#include "common.cuh"

namespace vdr {

// One thread writes 8 consecutive K-elements (16 bytes) of one patch row.
// A[m][k]: m = (b, py, px), k = (c, iy, ix);  K = 3*p*p, row pitch ldk = roundup(K, 8).
__global__ void __launch_bounds__(256)
im2col_patches_kernel(const float* __restrict__ src, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                      int gh, int gw, int patch, int K, int ldk, int64_t total_vec,
                      __nv_bfloat16* __restrict__ A) {
  const int vec_per_row = ldk >> 3;
  const int pp = patch * patch;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total_vec;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = v / vec_per_row;
    const int k0 = static_cast<int>(v - m * vec_per_row) << 3;
    const int px = static_cast<int>(m % gw);
    const int64_t t = m / gw;
    const int py = static_cast<int>(t % gh);
    const int64_t b = t / gh;
    const float* base = src + b * sb + (int64_t)(py * patch) * sy + (int64_t)(px * patch) * sx;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i;
      float val = 0.f;
      if (k < K) {
        const int c = k / pp;
        const int rem = k - c * pp;
        const int iy = rem / patch;
        const int ix = rem - iy * patch;
        val = __ldg(base + c * sc + iy * sy + ix * sx);
      }
      f[i] = val;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(A + m * ldk + k0) = o;
  }
}

// Destination layout of the staged slices.  Plain: (S, OH, OW).  Cell-padded (cin < cout): every cin x cin patch of the slice sits in
// the top-left corner of a cout x cout cell of a (S, OH / cin * cout, OW / cin * cout) image whose other pixels stay zero -- a 14-pixel
// patch grid becomes a 16-pixel one, which the TMA im2col view of the patch embedding can address (32-byte box rows; the weights get
// zero columns at the pad positions), vdr_volume_to_slices_cells.
struct CellMap {
  int cin, cout;   // 0, 0: plain
  __device__ __forceinline__ int64_t index(int s, int y, int x, int OH, int OW) const {
    if (cin == 0) return (static_cast<int64_t>(s) * OH + y) * OW + x;
    const int Y = y / cin * cout + y % cin, X = x / cin * cout + x % cin;
    const int OHp = OH / cin * cout, OWp = OW / cin * cout;
    return (static_cast<int64_t>(s) * OHp + Y) * OWp + X;
  }
};

// (H, W, S) f32 volume (slice index fastest, as np.dstack lays it out) -> (S, ch, cw) bf16 slices of
// the crop window, through a 32x33 smem tile so that both the reads (along S) and the writes (along x)
// are coalesced.  prepare_image's float32 cast becomes the bf16 operand cast here.
__global__ void __launch_bounds__(256)
volume_to_slices_kernel(const float* __restrict__ src, int W_full, int S, int y0, int x0, int ch, int cw,
                        __nv_bfloat16* __restrict__ dst, CellMap cm) {
  __shared__ float tile[32][33];
  const int y = blockIdx.y;
  const int xb = blockIdx.x * 32, sb = blockIdx.z * 32;
  const float* row = src + (static_cast<int64_t>(y0 + y) * W_full + x0) * S;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {          // j: x within tile, threadIdx.x: s within tile
    const int x = xb + j, sidx = sb + threadIdx.x;
    tile[j][threadIdx.x] = (x < cw && sidx < S) ? __ldg(row + static_cast<int64_t>(x) * S + sidx) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {          // j: s within tile, threadIdx.x: x within tile
    const int sidx = sb + j, x = xb + threadIdx.x;
    if (sidx < S && x < cw) dst[cm.index(sidx, y, x, ch, cw)] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// ---- prepare_image's resize on the device (scope row N2) ---------------------------------------------------------
// skimage.transform.resize(img, (OH, OW)) of a float image (tfds_dense_descriptor.py:42/44) = optional Gaussian
// anti-aliasing (only when an axis shrinks: sigma = (in/out - 1)/2, truncate 4, mode 'mirror') followed by
// scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True): output pixel o samples input coordinate
// (o + 0.5) * in/out - 0.5 with linear weights, indices mirrored about the edge pixel centres (d c b | a b c d | c b a).
__device__ __forceinline__ int mirror_idx(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n - 2;
  i %= p;
  if (i < 0) i += p;
  return i >= n ? p - i : i;
}

// One Gaussian pass along x (kAlongY = false) or y (true) over the crop window, layout (rows, cols, S) with the slice index
// fastest: consecutive threads walk s, so every tap is a coalesced read.  Weights are built once per block.
template <bool kAlongY>
__global__ void __launch_bounds__(256)
gauss_pass_kernel(const float* __restrict__ src, int64_t src_row_pitch, int64_t src_col_pitch, int rows, int cols, int S,
                  float sigma, int radius, float* __restrict__ dst) {
  extern __shared__ float wts[];   // [2 * radius + 1]
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int j = -radius; j <= radius; ++j) {
      const float x = static_cast<float>(j) / sigma;
      wts[j + radius] = expf(-0.5f * x * x);
      sum += wts[j + radius];
    }
    for (int j = 0; j <= 2 * radius; ++j) wts[j] /= sum;
  }
  __syncthreads();
  const int64_t total = static_cast<int64_t>(rows) * cols * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = static_cast<int>(i % S);
    const int64_t t = i / S;
    const int x = static_cast<int>(t % cols), y = static_cast<int>(t / cols);
    float acc = 0.f;
    for (int j = -radius; j <= radius; ++j) {
      const int yy = kAlongY ? mirror_idx(y + j, rows) : y, xx = kAlongY ? x : mirror_idx(x + j, cols);
      acc = fmaf(wts[j + radius], __ldg(src + yy * src_row_pitch + xx * src_col_pitch + s), acc);
    }
    dst[i] = acc;
  }
}

// Bilinear resample of the (rows, cols, S) window to (S, OH, OW) bf16 slices, transposed through a 32x33 smem tile like
// volume_to_slices_kernel (reads coalesced along s, writes along x).
__global__ void __launch_bounds__(256)
resize_to_slices_kernel(const float* __restrict__ src, int64_t src_row_pitch, int64_t src_col_pitch, int rows, int cols, int S,
                        int OH, int OW, float sy, float sx, __nv_bfloat16* __restrict__ dst, CellMap cm) {
  __shared__ float tile[32][33];
  const int Y = blockIdx.y;
  const int xb = blockIdx.x * 32, sb = blockIdx.z * 32;
  const float fy = (Y + 0.5f) * sy - 0.5f;
  const float fy0 = floorf(fy);
  const float wy = fy - fy0;
  const int y0 = mirror_idx(static_cast<int>(fy0), rows), y1 = mirror_idx(static_cast<int>(fy0) + 1, rows);
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {          // j: X within tile, threadIdx.x: s within tile
    const int X = xb + j, sidx = sb + threadIdx.x;
    float v = 0.f;
    if (X < OW && sidx < S) {
      const float fx = (X + 0.5f) * sx - 0.5f;
      const float fx0 = floorf(fx);
      const float wx = fx - fx0;
      const int x0 = mirror_idx(static_cast<int>(fx0), cols), x1 = mirror_idx(static_cast<int>(fx0) + 1, cols);
      const float a00 = __ldg(src + y0 * src_row_pitch + x0 * src_col_pitch + sidx), a01 = __ldg(src + y0 * src_row_pitch + x1 * src_col_pitch + sidx);
      const float a10 = __ldg(src + y1 * src_row_pitch + x0 * src_col_pitch + sidx), a11 = __ldg(src + y1 * src_row_pitch + x1 * src_col_pitch + sidx);
      const float top = fmaf(wx, a01 - a00, a00), bot = fmaf(wx, a11 - a10, a10);
      v = fmaf(wy, bot - top, top);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {          // j: s within tile, threadIdx.x: X within tile
    const int sidx = sb + j, X = xb + threadIdx.x;
    if (sidx < S && X < OW) dst[cm.index(sidx, Y, X, OH, OW)] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// im2col from bf16 slices (B, H, W), gray replicated to 3 channels: 16-byte copies when patch % 8 == 0.
__global__ void __launch_bounds__(256)
im2col_gray_bf16_kernel(const __nv_bfloat16* __restrict__ src, int H, int W, int gh, int gw, int patch, int K,
                        int ldk, int64_t total_vec, __nv_bfloat16* __restrict__ A) {
  const int vec_per_row = ldk >> 3;
  const int pp = patch * patch;
  const bool vec = (patch & 7) == 0 && (W & 7) == 0;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total_vec;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = v / vec_per_row;
    const int k0 = static_cast<int>(v - m * vec_per_row) << 3;
    const int px = static_cast<int>(m % gw);
    const int64_t t = m / gw;
    const int py = static_cast<int>(t % gh);
    const int64_t b = t / gh;
    const __nv_bfloat16* base = src + (b * H + static_cast<int64_t>(py) * patch) * W + px * patch;
    uint4 o;
    if (vec && k0 + 8 <= K) {
      const int rem = k0 % pp;                           // channel ignored: all three are the same image
      const int iy = rem / patch, ix = rem - iy * patch;
      o = __ldg(reinterpret_cast<const uint4*>(base + iy * W + ix));
    } else {
      __nv_bfloat16 e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        if (k < K) {
          const int rem = k % pp;
          const int iy = rem / patch, ix = rem - iy * patch;
          e[i] = base[iy * W + ix];
        } else {
          e[i] = __float2bfloat16_rn(0.f);
        }
      }
      o = *reinterpret_cast<uint4*>(e);
    }
    *reinterpret_cast<uint4*>(A + m * ldk + k0) = o;
  }
}

__global__ void write_cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos0,
                                      __nv_bfloat16* __restrict__ X, int B, int N, int d) {
  const int64_t total = (int64_t)B * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d;
    const int c = static_cast<int>(i - b * d);
    X[b * N * d + c] = __float2bfloat16_rn(cls[c] + pos0[c]);
  }
}

}  // namespace vdr

extern "C" int vdr_im2col_patches(const float* src, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int B, int H,
                                  int W, int patch, void* A_bf16, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(src && A_bf16, VDR_EINVAL, "vdr_im2col_patches: null pointer");
  VDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && patch > 0, VDR_EINVAL, "vdr_im2col_patches: bad shape");
  VDR_CHECK_ARG(H % patch == 0 && W % patch == 0, VDR_EINVAL, "vdr_im2col_patches: H (%d), W (%d) must be multiples of patch (%d)", H, W, patch);
  VDR_CHECK_ARG(aligned16(A_bf16), VDR_EALIGN, "vdr_im2col_patches: output must be 16-byte aligned");
  const int gh = H / patch, gw = W / patch;
  const int K = 3 * patch * patch;
  const int ldk = (K + 7) & ~7;
  const int64_t total_vec = (int64_t)B * gh * gw * (ldk >> 3);
  const int threads = 256;
  int64_t blocks = (total_vec + threads - 1) / threads;
  const int64_t max_blocks = (int64_t)num_sms() * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  im2col_patches_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, sb, sc, sy, sx, gh, gw, patch, K, ldk, total_vec, static_cast<__nv_bfloat16*>(A_bf16));
  count_launch();
  VDR_CHECK_LAUNCH("im2col_patches_kernel");
  return VDR_OK;
}

static int volume_to_slices_impl(const char* who, const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw, void* slices_bf16,
                                 vdr::CellMap cm, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(vol && slices_bf16, VDR_EINVAL, "%s: null pointer", who);
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0 && ch > 0 && cw > 0 && y0 >= 0 && x0 >= 0 && y0 + ch <= H && x0 + cw <= W, VDR_EINVAL,
                "%s: crop window (%d:%d, %d:%d) outside the %dx%d volume", who, y0, y0 + ch, x0, x0 + cw, H, W);
  VDR_CHECK_ARG(ch <= 65535 && (S + 31) / 32 <= 65535, VDR_EINVAL, "%s: volume too large for the launch grid", who);
  dim3 grid((cw + 31) / 32, ch, (S + 31) / 32), block(32, 8);
  volume_to_slices_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      vol, W, S, y0, x0, ch, cw, static_cast<__nv_bfloat16*>(slices_bf16), cm);
  count_launch();
  VDR_CHECK_LAUNCH("volume_to_slices_kernel");
  return VDR_OK;
}

extern "C" int vdr_volume_to_slices(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw,
                                    void* slices_bf16, vdr_stream_t stream) {
  return volume_to_slices_impl("vdr_volume_to_slices", vol, H, W, S, y0, x0, ch, cw, slices_bf16, vdr::CellMap{0, 0}, stream);
}

static void resize_sigmas(int ch, int cw, int OH, int OW, float* sig_y, float* sig_x) {
  // skimage: anti-aliasing is on when ANY axis shrinks; per-axis sigma = max(0, (in/out - 1) / 2)
  const bool aa = OH < ch || OW < cw;
  const float fy = static_cast<float>(ch) / OH, fx = static_cast<float>(cw) / OW;
  *sig_y = aa ? fmaxf(0.f, (fy - 1.f) * 0.5f) : 0.f;
  *sig_x = aa ? fmaxf(0.f, (fx - 1.f) * 0.5f) : 0.f;
}
static int gauss_radius(float sigma) { return sigma > 1e-15f ? static_cast<int>(4.0f * sigma + 0.5f) : 0; }

extern "C" size_t vdr_volume_to_slices_resized_workspace_bytes(int S, int ch, int cw, int OH, int OW) {
  float sy, sx;
  resize_sigmas(ch, cw, OH, OW, &sy, &sx);
  const int passes = (gauss_radius(sy) > 0 ? 1 : 0) + (gauss_radius(sx) > 0 ? 1 : 0);
  return static_cast<size_t>(passes) * ch * cw * S * sizeof(float);
}

static int volume_to_slices_resized_impl(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw, int OH, int OW,
                                         void* slices_bf16, void* workspace, size_t workspace_bytes, vdr::CellMap cm,
                                         vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(vol && slices_bf16, VDR_EINVAL, "vdr_volume_to_slices_resized: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0 && ch > 0 && cw > 0 && y0 >= 0 && x0 >= 0 && y0 + ch <= H && x0 + cw <= W, VDR_EINVAL,
                "vdr_volume_to_slices_resized / _cells: crop window (%d:%d, %d:%d) outside the %dx%d volume", y0, y0 + ch, x0, x0 + cw, H, W);
  VDR_CHECK_ARG(OH > 0 && OW > 0 && OH <= 65535 && (S + 31) / 32 <= 65535, VDR_EINVAL, "vdr_volume_to_slices_resized: bad output size %dx%d", OH, OW);
  const size_t need = vdr_volume_to_slices_resized_workspace_bytes(S, ch, cw, OH, OW);
  VDR_CHECK_ARG(need == 0 || (workspace && workspace_bytes >= need), VDR_EWORKSPACE,
                "vdr_volume_to_slices_resized: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float sig_y, sig_x;
  resize_sigmas(ch, cw, OH, OW, &sig_y, &sig_x);
  const int ry = gauss_radius(sig_y), rx = gauss_radius(sig_x);
  VDR_CHECK_ARG(ry <= 4096 && rx <= 4096, VDR_EINVAL, "vdr_volume_to_slices_resized: shrink factor too large");
  // source window: rows y0.., cols x0.. of the (H, W, S) volume
  const float* src = vol + (static_cast<int64_t>(y0) * W + x0) * S;
  int64_t row_pitch = static_cast<int64_t>(W) * S, col_pitch = S;
  float* tmp = static_cast<float*>(workspace);
  const int64_t total = static_cast<int64_t>(ch) * cw * S;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  if (rx > 0) {
    gauss_pass_kernel<false><<<(unsigned)blocks, 256, (2 * rx + 1) * sizeof(float), s>>>(src, row_pitch, col_pitch, ch, cw, S, sig_x, rx, tmp);
    count_launch();
    VDR_CHECK_LAUNCH("gauss_pass_kernel<x>");
    src = tmp; row_pitch = static_cast<int64_t>(cw) * S; col_pitch = S;
    tmp += total;
  }
  if (ry > 0) {
    gauss_pass_kernel<true><<<(unsigned)blocks, 256, (2 * ry + 1) * sizeof(float), s>>>(src, row_pitch, col_pitch, ch, cw, S, sig_y, ry, tmp);
    count_launch();
    VDR_CHECK_LAUNCH("gauss_pass_kernel<y>");
    src = tmp; row_pitch = static_cast<int64_t>(cw) * S; col_pitch = S;
  }
  dim3 grid((OW + 31) / 32, OH, (S + 31) / 32), block(32, 8);
  resize_to_slices_kernel<<<grid, block, 0, s>>>(src, row_pitch, col_pitch, ch, cw, S, OH, OW, static_cast<float>(ch) / OH,
                                                static_cast<float>(cw) / OW, static_cast<__nv_bfloat16*>(slices_bf16), cm);
  count_launch();
  VDR_CHECK_LAUNCH("resize_to_slices_kernel");
  return VDR_OK;
}

extern "C" int vdr_volume_to_slices_resized(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw, int OH,
                                            int OW, void* slices_bf16, void* workspace, size_t workspace_bytes,
                                            vdr_stream_t stream) {
  return volume_to_slices_resized_impl(vol, H, W, S, y0, x0, ch, cw, OH, OW, slices_bf16, workspace, workspace_bytes, vdr::CellMap{0, 0}, stream);
}

// The same staging into the CELL-PADDED layout (see CellMap): output pixel (y, x) of the (OH, OW) slice goes to
// (y / cell_in * cell_out + y % cell_in, x / cell_in * cell_out + x % cell_in) of a (S, OH / cell_in * cell_out, OW / cell_in * cell_out)
// image; the caller zeroes that buffer once (the pad pixels are never written).  OH x OW == ch x cw: no resize, workspace unused.
extern "C" int vdr_volume_to_slices_cells(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw, int OH, int OW,
                                          int cell_in, int cell_out, void* slices_bf16, void* workspace, size_t workspace_bytes,
                                          vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(cell_in > 0 && cell_out >= cell_in && OH > 0 && OW > 0 && OH % cell_in == 0 && OW % cell_in == 0, VDR_EINVAL,
                "vdr_volume_to_slices_cells: %dx%d slices do not tile into %d-pixel cells (cell_out %d)", OH, OW, cell_in, cell_out);
  const CellMap cm{cell_in, cell_out};
  if (OH == ch && OW == cw) return volume_to_slices_impl("vdr_volume_to_slices_cells", vol, H, W, S, y0, x0, ch, cw, slices_bf16, cm, stream);
  return volume_to_slices_resized_impl(vol, H, W, S, y0, x0, ch, cw, OH, OW, slices_bf16, workspace, workspace_bytes, cm, stream);
}

extern "C" int vdr_im2col_gray_bf16(const void* slices_bf16, int B, int H, int W, int patch, void* A_bf16,
                                    vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(slices_bf16 && A_bf16, VDR_EINVAL, "vdr_im2col_gray_bf16: null pointer");
  VDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && patch > 0 && H % patch == 0 && W % patch == 0, VDR_EINVAL,
                "vdr_im2col_gray_bf16: H (%d), W (%d) must be positive multiples of patch (%d)", H, W, patch);
  VDR_CHECK_ARG(aligned16(slices_bf16) && aligned16(A_bf16), VDR_EALIGN, "vdr_im2col_gray_bf16: pointers must be 16-byte aligned");
  const int gh = H / patch, gw = W / patch;
  const int K = 3 * patch * patch;
  const int ldk = (K + 7) & ~7;
  const int64_t total_vec = (int64_t)B * gh * gw * (ldk >> 3);
  int64_t blocks = (total_vec + 255) / 256;
  const int64_t max_blocks = (int64_t)num_sms() * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  im2col_gray_bf16_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(slices_bf16), H, W, gh, gw, patch, K, ldk, total_vec,
      static_cast<__nv_bfloat16*>(A_bf16));
  count_launch();
  VDR_CHECK_LAUNCH("im2col_gray_bf16_kernel");
  return VDR_OK;
}

extern "C" int vdr_write_cls_rows(const float* cls, const float* pos0, void* X_bf16, int B, int N, int d,
                                  vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(cls && pos0 && X_bf16, VDR_EINVAL, "vdr_write_cls_rows: null pointer");
  VDR_CHECK_ARG(B > 0 && N > 0 && d > 0, VDR_EINVAL, "vdr_write_cls_rows: bad shape");
  const int64_t total = (int64_t)B * d;
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  if (blocks > 1184) blocks = 1184;
  write_cls_rows_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cls, pos0, static_cast<__nv_bfloat16*>(X_bf16), B, N, d);
  count_launch();
  VDR_CHECK_LAUNCH("write_cls_rows_kernel");
  return VDR_OK;
}
