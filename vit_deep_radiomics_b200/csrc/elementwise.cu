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

// (H, W, S) f32 volume (slice index fastest, as np.dstack lays it out) -> (S, ch, cw) bf16 slices of
// the crop window, through a 32x33 smem tile so that both the reads (along S) and the writes (along x)
// are coalesced.  prepare_image's float32 cast becomes the bf16 operand cast here.
__global__ void __launch_bounds__(256)
volume_to_slices_kernel(const float* __restrict__ src, int W_full, int S, int y0, int x0, int ch, int cw,
                        __nv_bfloat16* __restrict__ dst) {
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
    if (sidx < S && x < cw) dst[(static_cast<int64_t>(sidx) * ch + y) * cw + x] = __float2bfloat16_rn(tile[threadIdx.x][j]);
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

extern "C" int vdr_volume_to_slices(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw,
                                    void* slices_bf16, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(vol && slices_bf16, VDR_EINVAL, "vdr_volume_to_slices: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0 && ch > 0 && cw > 0 && y0 >= 0 && x0 >= 0 && y0 + ch <= H && x0 + cw <= W, VDR_EINVAL,
                "vdr_volume_to_slices: crop window (%d:%d, %d:%d) outside the %dx%d volume", y0, y0 + ch, x0, x0 + cw, H, W);
  VDR_CHECK_ARG(ch <= 65535 && (S + 31) / 32 <= 65535, VDR_EINVAL, "vdr_volume_to_slices: volume too large for the launch grid");
  dim3 grid((cw + 31) / 32, ch, (S + 31) / 32), block(32, 8);
  volume_to_slices_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      vol, W, S, y0, x0, ch, cw, static_cast<__nv_bfloat16*>(slices_bf16));
  count_launch();
  VDR_CHECK_LAUNCH("volume_to_slices_kernel");
  return VDR_OK;
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
