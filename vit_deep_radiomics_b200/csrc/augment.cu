// Offline augmentation of the extraction loop on the device:
//   flip_image   (tfds_dense_descriptor.py:306-325)  image[:, ::-1] / image[::-1]           -> an index transform of the source read
//   rotate_image (tfds_dense_descriptor.py:328-350)  scipy.ndimage.rotate(x, angle, axes=(0, 1), reshape=False, mode='nearest')
//                                                    (cubic B-spline), image clipped to [0, 1], mask re-binarised (> 0)
// The rotation restates what scipy.ndimage does for that call, operation for operation in IEEE double without contraction
// (pinned bit for bit against scipy 1.18 on the CPU, see tests/test_gpu_augment.py and oracle/rotate_np.py):
//   1. every (row, col) plane is padded by 12 pixels of edge values (ndimage._prepad_for_spline_filter, mode 'nearest');
//   2. cubic spline prefilter along axis 0, then axis 1 (ni_splines.c: apply_filter with the 'reflect' initialisers that scipy
//      uses for mode 'nearest'): c *= gain; c[0] = z (c[0] + z^n c[n-1] + sum_i z^i (c[i] + z^n c[n-1-i])) / (1 - z^2n) + c[0];
//      causal c[i] += z c[i-1]; c[n-1] *= z / (z - 1); anticausal c[i] = z (c[i+1] - c[i]);  z = the pole constant of the binary;
//   3. NI_GeometricTransform: input coordinate cc = ((offset + i m0) + j m1) + 12 per axis, NOT mapped into the array for mode
//      'nearest'; start = floor(cc) - 1; cubic weights from y = cc - floor(cc); the four taps per axis are clamped into the padded
//      plane; t = sum_a sum_b (c[a][b] w_a) w_b in that order;
//   4. output conversion: float image = (float)t, then clip(0, 1);  bool mask = (unsigned char)t != 0, i.e. |t| >= 1;
//      uint8 mask = (t > 0 ? t + 0.5 : 0) truncated, then > 0.
// The rotation matrix / offset come from the host (special.cosdg / sindg and NumPy, as scipy computes them).
// HBM-bound: the padded f64 planes are written once and swept four times per axis (strided lines, coalesced across planes).
#include "common.cuh"

namespace vdr {

constexpr int kRotPad = 12;
// sqrt(3) - 2 as folded into scipy's binary (2 ulp from the double-precision evaluation of sqrt(3.0) - 2.0)
constexpr double kRotPole = -0x1.126145e9ecd56p-2;

struct RotGeom {
  int H, W, NP;          // plane extents and number of planes (slices x channels); element (r, c, p) at (r*W + c)*NP + p
  int HP, WP;            // padded extents
  int flip;              // 0 none, 1 horizontal (columns reversed), 2 vertical (rows reversed): applied to the SOURCE read
};

template <typename T>
__device__ __forceinline__ double rot_src(const T* __restrict__ src, const RotGeom& g, int rp, int cp, int p) {
  int r = min(max(rp - kRotPad, 0), g.H - 1), c = min(max(cp - kRotPad, 0), g.W - 1);
  if (g.flip == 1) c = g.W - 1 - c;
  if (g.flip == 2) r = g.H - 1 - r;
  return static_cast<double>(src[(static_cast<int64_t>(r) * g.W + c) * g.NP + p]);
}

// One thread = one line of the padded array along `axis`; lines are enumerated with the plane index fastest, so a warp's
// accesses are 32 consecutive doubles at every step of the recursion.  kFirst: the line is read from the (flipped, edge-padded)
// source; otherwise from P itself (second axis).
template <typename T, bool kFirst>
__global__ void __launch_bounds__(128) rot_prefilter_kernel(const T* __restrict__ src, double* __restrict__ P, RotGeom g, int axis, double z_n) {
  const int64_t other = axis == 0 ? g.WP : g.HP;
  const int64_t line = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (line >= other * g.NP) return;
  const int p = static_cast<int>(line % g.NP), o = static_cast<int>(line / g.NP);
  const int n = axis == 0 ? g.HP : g.WP;
  const int64_t stride = axis == 0 ? static_cast<int64_t>(g.WP) * g.NP : g.NP;
  double* c = P + (axis == 0 ? static_cast<int64_t>(o) * g.NP : static_cast<int64_t>(o) * g.WP * g.NP) + p;
  const double z = kRotPole;
  const double gain = __dmul_rn(__dsub_rn(1.0, __ddiv_rn(1.0, z)), __dsub_rn(1.0, z));
  // Element i of the line after scipy's gain pass: the (flipped, edge-padded) source sample (first axis) or what the first axis left in
  // P (second axis), times the gain.  Evaluated on the fly wherever it is needed -- the same product, so the same double -- instead of
  // a read-modify-write sweep of its own.  The lines are latency-bound recurrences (64 k threads, one dependent chain each): the loads
  // of kB consecutive elements are issued together, ahead of the arithmetic that consumes them and of the stores of the batch (the
  // compiler cannot hoist a load of c[] above a store to c[] itself).
  auto raw = [&](int i) -> double { return kFirst ? (axis == 0 ? rot_src(src, g, i, o, p) : rot_src(src, g, o, i, p)) : c[i * stride]; };
  constexpr int kB = 8;
  // causal initialisation (_init_causal_reflect): c0 + z^n c[n-1] + sum_i z^i (c[i] + z^n c[n-1-i]), terms added in index order
  const double c0 = __dmul_rn(raw(0), gain);
  double acc = __dadd_rn(__dmul_rn(__dmul_rn(raw(n - 1), gain), z_n), c0);
  double z_i = z;
  for (int i0 = 1; i0 < n; i0 += kB) {
    double a[kB], r[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      const int i = min(i0 + k, n - 1);
      a[k] = raw(i);
      r[k] = raw(n - 1 - i);
    }
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      if (i0 + k < n) {
        const double t = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(r[k], gain), z_n), __dmul_rn(a[k], gain)), z_i);
        z_i = __dmul_rn(z_i, z);
        acc = __dadd_rn(acc, t);
      }
    }
  }
  double prev = __dadd_rn(__ddiv_rn(__dmul_rn(z, acc), __dsub_rn(1.0, __dmul_rn(z_n, z_n))), c0);
  c[0] = prev;
  // causal: c[i] = c[i-1] z + c[i]
  for (int i0 = 1; i0 < n; i0 += kB) {
    double a[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) a[k] = raw(min(i0 + k, n - 1));
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      if (i0 + k < n) {
        prev = __dadd_rn(__dmul_rn(prev, z), __dmul_rn(a[k], gain));
        c[(i0 + k) * stride] = prev;
      }
    }
  }
  // anticausal: c[n-1] *= z / (z - 1); c[i] = z (c[i+1] - c[i])
  prev = __dmul_rn(__ddiv_rn(z, __dsub_rn(z, 1.0)), prev);
  c[(n - 1) * stride] = prev;
  for (int i0 = n - 2; i0 >= 0; i0 -= kB) {
    double a[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) a[k] = c[max(i0 - k, 0) * stride];
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      if (i0 - k >= 0) {
        prev = __dmul_rn(__dsub_rn(prev, a[k]), z);
        c[(i0 - k) * stride] = prev;
      }
    }
  }
}

struct RotXform {
  double m00, m01, m10, m11, off0, off1;
};

__device__ __forceinline__ void rot_weights(double cc, int& start, double (&w)[4]) {
  const double fl = floor(cc);
  const double y = __dsub_rn(cc, fl), zz = __dsub_rn(1.0, y);
  start = static_cast<int>(fl) - 1;
  w[0] = __ddiv_rn(__dmul_rn(zz, __dmul_rn(zz, zz)), 6.0);
  w[1] = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(y, -2.0), __dmul_rn(y, y)), 3.0), 4.0), 6.0);
  w[2] = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(zz, -2.0), __dmul_rn(zz, zz)), 3.0), 4.0), 6.0);
  w[3] = __dsub_rn(__dsub_rn(__dsub_rn(1.0, w[0]), w[1]), w[2]);
}

// kOut: 0 = f32 image clipped to [0, 1]; 1 = mask from a bool input; 2 = mask from a uint8 input
// One warp per output pixel (i, j), lanes over the planes: the input coordinate, the two sets of cubic weights (six f64 divisions) and the
// clamped tap indices depend on the pixel only and are evaluated once per warp instead of once per element; every tap is then one coalesced
// 256-byte read of 32 consecutive planes.
template <int kOut>
__global__ void __launch_bounds__(256) rot_interp_kernel(const double* __restrict__ P, RotGeom g, RotXform x, void* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t pixels = static_cast<int64_t>(g.H) * g.W;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5, nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t rc = warp0; rc < pixels; rc += nwarps) {
    const int j = static_cast<int>(rc % g.W), i = static_cast<int>(rc / g.W);
    const double di = static_cast<double>(i), dj = static_cast<double>(j);
    const double cc0 = __dadd_rn(__dadd_rn(__dadd_rn(x.off0, __dmul_rn(di, x.m00)), __dmul_rn(dj, x.m01)), static_cast<double>(kRotPad));
    const double cc1 = __dadd_rn(__dadd_rn(__dadd_rn(x.off1, __dmul_rn(di, x.m10)), __dmul_rn(dj, x.m11)), static_cast<double>(kRotPad));
    int s0, s1;
    double w0[4], w1[4];
    rot_weights(cc0, s0, w0);
    rot_weights(cc1, s1, w1);
    int64_t row[4];
    int col[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      row[a] = static_cast<int64_t>(min(max(s0 + a, 0), g.HP - 1)) * g.WP;
      col[a] = min(max(s1 + a, 0), g.WP - 1);
    }
    for (int p = lane; p < g.NP; p += 32) {
      double t = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double c = P[(row[a] + col[b]) * g.NP + p];
          t = __dadd_rn(t, __dmul_rn(__dmul_rn(c, w0[a]), w1[b]));
        }
      }
      const int64_t e = rc * g.NP + p;
      if (kOut == 0) {
        const float v = __double2float_rn(t);
        static_cast<float*>(out)[e] = fminf(fmaxf(v, 0.f), 1.f);                         // np.clip(image_rot, 0, 1)
      } else if (kOut == 1) {
        static_cast<uint8_t*>(out)[e] = (t >= 1.0 || t <= -1.0) ? 1 : 0;                 // (npy_bool)t, then > 0
      } else {
        static_cast<uint8_t*>(out)[e] = (t > 0.0 && __dadd_rn(t, 0.5) >= 1.0) ? 1 : 0;   // round-half-up to uint8, then > 0
      }
    }
  }
}

// flips without rotation (angle 0): a permuting copy
template <typename T>
__global__ void __launch_bounds__(256) flip_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W, int NP, int flip, int binarise) {
  const int64_t total = static_cast<int64_t>(H) * W * NP;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(e % NP);
    const int64_t rc = e / NP;
    int c = static_cast<int>(rc % W), r = static_cast<int>(rc / W);
    if (flip == 1) c = W - 1 - c;
    if (flip == 2) r = H - 1 - r;
    T v = src[(static_cast<int64_t>(r) * W + c) * NP + p];
    if (binarise) v = v != T(0) ? T(1) : T(0);
    dst[e] = v;
  }
}

}  // namespace vdr

extern "C" size_t vdr_rotate_workspace_bytes(int H, int W, int planes) {
  return static_cast<size_t>(H + 2 * vdr::kRotPad) * static_cast<size_t>(W + 2 * vdr::kRotPad) * static_cast<size_t>(planes) * sizeof(double);
}

static int rot_grid(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(vdr::num_sms()) * 16;
  return static_cast<int>(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int vdr_flip_rotate_volume(const void* src, int src_kind, void* dst, int H, int W, int planes, int flip, int rotate,
                                      const double* xform_host, double z_n_rows, double z_n_cols, void* workspace, size_t workspace_bytes,
                                      vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(src && dst, VDR_EINVAL, "vdr_flip_rotate_volume: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && planes > 0 && (int64_t)H * W * planes < (1ll << 40), VDR_EINVAL, "vdr_flip_rotate_volume: bad shape");
  VDR_CHECK_ARG(flip >= 0 && flip <= 2, VDR_EINVAL, "vdr_flip_rotate_volume: flip must be 0 (none), 1 (horizontal) or 2 (vertical)");
  VDR_CHECK_ARG(src_kind >= 0 && src_kind <= 2, VDR_EINVAL, "vdr_flip_rotate_volume: src_kind must be 0 (f32 image), 1 (bool mask) or 2 (uint8 mask)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)H * W * planes;
  if (!rotate) {
    if (src_kind == 0)
      flip_copy_kernel<float><<<rot_grid(total, 256), 256, 0, s>>>(static_cast<const float*>(src), static_cast<float*>(dst), H, W, planes, flip, 0);
    else
      flip_copy_kernel<uint8_t><<<rot_grid(total, 256), 256, 0, s>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), H, W, planes, flip, 1);
    count_launch();
    VDR_CHECK_LAUNCH("flip_copy_kernel");
    return VDR_OK;
  }
  VDR_CHECK_ARG(xform_host && workspace, VDR_EINVAL, "vdr_flip_rotate_volume: a rotation needs xform_host (m00, m01, m10, m11, off0, off1) and a workspace");
  VDR_CHECK_ARG(workspace_bytes >= vdr_rotate_workspace_bytes(H, W, planes), VDR_EWORKSPACE, "vdr_flip_rotate_volume: workspace too small (%zu < %zu)",
                workspace_bytes, vdr_rotate_workspace_bytes(H, W, planes));
  VDR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, VDR_EALIGN, "vdr_flip_rotate_volume: workspace must be 8-byte aligned");
  RotGeom g{H, W, planes, H + 2 * kRotPad, W + 2 * kRotPad, flip};
  double* P = static_cast<double*>(workspace);
  const int64_t lines0 = (int64_t)g.WP * planes, lines1 = (int64_t)g.HP * planes;
  if (src_kind == 0)
    rot_prefilter_kernel<float, true><<<(unsigned)((lines0 + 127) / 128), 128, 0, s>>>(static_cast<const float*>(src), P, g, 0, z_n_rows);
  else
    rot_prefilter_kernel<uint8_t, true><<<(unsigned)((lines0 + 127) / 128), 128, 0, s>>>(static_cast<const uint8_t*>(src), P, g, 0, z_n_rows);
  VDR_CHECK_LAUNCH("rot_prefilter_kernel (axis 0)");
  rot_prefilter_kernel<float, false><<<(unsigned)((lines1 + 127) / 128), 128, 0, s>>>(nullptr, P, g, 1, z_n_cols);
  VDR_CHECK_LAUNCH("rot_prefilter_kernel (axis 1)");
  const RotXform x{xform_host[0], xform_host[1], xform_host[2], xform_host[3], xform_host[4], xform_host[5]};
  const int igrid = rot_grid((int64_t)H * W * 32, 256);              // one warp per pixel
  if (src_kind == 0) rot_interp_kernel<0><<<igrid, 256, 0, s>>>(P, g, x, dst);
  else if (src_kind == 1) rot_interp_kernel<1><<<igrid, 256, 0, s>>>(P, g, x, dst);
  else rot_interp_kernel<2><<<igrid, 256, 0, s>>>(P, g, x, dst);
  count_launch(3);
  VDR_CHECK_LAUNCH("rot_interp_kernel");
  return VDR_OK;
}
