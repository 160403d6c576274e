// Windowed attention with decomposed relative-position bias of SAM / MedSAM's image encoder (14 x 14 windows, head_dim 64) on
// tcgen05 -- the eight windowed blocks of the reference's default backbone (segment_anything ImageEncoderViT, Block.forward with
// window_size 14; what load_medsam builds, tfds_dense_descriptor.py:91-107).  Replaces the mma.sync kernel for this shape.
//
// One CTA = (window, head, half): 98 query tokens (7 of the window's 14 rows) against all 196 keys, everything in one shot (no
// key loop, no online softmax).  Two CTAs per SM.
//
//   TMA      Q (98 x 64), K, V (196 x 64) straight from the un-partitioned qkv matrix through 4-D tensor maps
//            (dim, x, y, image): positions beyond the image arrive as zeros and are overwritten with bf16(qkv bias) -- the
//            reference pads AFTER norm1, so a pad token is exactly "q = k = v = bias" and takes part in the softmax.
//   T        = Q [R_hi ; R_lo]^T   (128 x 64 x 64, twice, one accumulator): q . rel_pos_h[i], q . rel_pos_w[i] for all 27 + 27
//            relative offsets (the R tiles borrow the E tile's space); a query thread picks its 14 + 14 terms (a barrel shift by its
//            own (qh, qw) over registers).
//   S'       = [Q | E] [K | onehot]^T  (128 x 208 x 128): the bias is folded INTO the score MMA.  E holds the query row's
//            28 bias terms / scale as bf16 hi + lo parts (fp32-class accuracy) and a constant 1; the matching 64 extra K columns
//            are one-hot in (kh, kw) (a constant tile, the same for every window) and -30000 in the constant's column for the
//            12 pad keys (196 -> 208), which masks them.  The softmax threads never touch the bias.
//   softmax  one thread per query row, two passes over TMEM (max, then exp2 / sum); P (bf16 pairs) overwrites the S' columns
//            it was computed from, O = P V (128 x 64 x 208, V MN-major straight from its TMA tile) lands in S' columns 128..191.
//
// 5 warps: 0-3 softmax / fix-ups (TMEM lanes 32 w ..), 4 = TMA + MMA issue (elect.sync).  TMEM 256 columns.
#include "common.cuh"

namespace vdr {

constexpr int kW14 = 14, kW14N = 196, kW14NP = 208, kW14Half = 98;
constexpr int kW14Threads = 160;
constexpr int kW14Tile128 = 128 * 128;                      // 16 KB: 128 rows x 64 bf16
constexpr int kW14Tile208 = kW14NP * 128;                   // 26 KB
constexpr int kW14OffQm = 0, kW14OffQx = kW14Tile128, kW14OffKm = 2 * kW14Tile128, kW14OffKx = kW14OffKm + kW14Tile208,
              kW14OffV = kW14OffKx + kW14Tile208, kW14OffBar = kW14OffV + kW14Tile208;
constexpr int kW14Smem = kW14OffBar + 128;
constexpr int kW14TmemCols = 256;
constexpr float kW14MaskValue = -30000.f;

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, uint32_t dst_smem, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ float w14_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// MN-major B operand (V tile: kv rows x 64 d, 128-byte rows, SWIZZLE_128B), 16 kv rows per MMA K step (see attention.cu)
__device__ __forceinline__ uint64_t w14_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// The constant half of the extended key tile: row k = (kh, kw) is one-hot at kh, 14 + kw (hi parts) and 28 + kh, 42 + kw (lo
// parts); column 56 pairs with the constant 1 of the query side: 0 for real keys, the mask value for the 12 pad rows.
__device__ __nv_bfloat16 g_w14_kx[kW14NP * 64];
__global__ void w14_init_kx_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kW14NP * 64) return;
  const int k = i >> 6, c = i & 63;
  float v = 0.f;
  if (k < kW14N) {
    const int kh = k / kW14, kw = k % kW14;
    if (c == kh || c == 14 + kw || c == 28 + kh || c == 42 + kw) v = 1.f;
  } else if (c == 56) {
    v = kW14MaskValue;
  }
  g_w14_kx[i] = __float2bfloat16_rn(v);
}

struct Win14Params {
  const float* qkv_bias;         // (3 * heads * 64) f32
  __nv_bfloat16* out;
  int64_t ld_out;
  int gh, gw, nwh, nww, heads;
  float scale_log2, inv_scale;
};

// rel[k] = t[base + 13 + q - k], k = 0..13, q in [0, 14): a barrel shift over registers (static indices only)
template <int kBase>
__device__ __forceinline__ void w14_pick(const float (&t)[64], int q, float (&rel)[14]) {
  float a[29];
#pragma unroll
  for (int i = 0; i < 27; ++i) a[i] = t[kBase + i];
  a[27] = 0.f;
  a[28] = 0.f;
  const bool b8 = q & 8, b4 = q & 4, b2 = q & 2, b1 = q & 1;
#pragma unroll
  for (int i = 0; i <= 20; ++i) a[i] = b8 ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i <= 16; ++i) a[i] = b4 ? a[i + 4] : a[i];
#pragma unroll
  for (int i = 0; i <= 14; ++i) a[i] = b2 ? a[i + 2] : a[i];
#pragma unroll
  for (int i = 0; i <= 13; ++i) a[i] = b1 ? a[i + 1] : a[i];
#pragma unroll
  for (int k = 0; k < 14; ++k) rel[k] = a[13 - k];
}

__global__ void __launch_bounds__(kW14Threads, 2)
attn_win14_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmRhi, const __grid_constant__ CUtensorMap tmRlo,
                     const __grid_constant__ CUtensorMap tmKx, const Win14Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sQm = base + kW14OffQm, sQx = base + kW14OffQx, sKm = base + kW14OffKm, sKx = base + kW14OffKx, sV = base + kW14OffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kW14OffBar);
  uint64_t* bar_ld = bars;         // Q, K, V landed
  uint64_t* bar_r = bars + 1;      // R_hi, R_lo landed
  uint64_t* bar_t = bars + 2;      // T complete
  uint64_t* bar_kx = bars + 3;     // the constant key columns landed
  uint64_t* bar_qx = bars + 4;     // E written, fix-ups done (128 arrivals)
  uint64_t* bar_s = bars + 5;      // S' complete
  uint64_t* bar_p = bars + 6;      // P stored (128 arrivals)
  uint64_t* bar_o = bars + 7;      // O complete
  uint64_t* bar_fix = bars + 8;    // pad tokens overwritten (128 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = blockIdx.x, head = blockIdx.y;
  const int win = blockIdx.z;
  const int wx = win % p.nww, wy = (win / p.nww) % p.nwh, b = win / (p.nww * p.nwh);
  const int d = p.heads * 64;
  const int x0 = wx * kW14, y0 = wy * kW14;

  if (tid == 128) {
    if (base & 1023u) { printf("vdr: window attention smem base not 1024-byte aligned\n"); __trap(); }
    mbar_init(bar_ld, 1);
    mbar_init(bar_r, 1);
    mbar_init(bar_t, 1);
    mbar_init(bar_kx, 1);
    mbar_init(bar_qx, 128);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_fix, 128);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_ld, kW14Half * 128 + 2 * kW14N * 128);
    tma_load_4d(&tmQ, bar_ld, sQm, head * 64, x0, y0 + 7 * half, b);
    mbar_arrive_expect_tx(bar_r, 2 * 64 * 128);
    tma_load_2d_addr(&tmRhi, bar_r, sQx, 0, 0);            // R_hi / R_lo borrow the E tile's space: E is written only after T is complete
    tma_load_2d_addr(&tmRlo, bar_r, sQx + 8192, 0, 0);
    mbar_arrive_expect_tx(bar_kx, kW14Tile208);
    tma_load_2d_addr(&tmKx, bar_kx, sKx, 0, 0);
    tma_load_4d(&tmKV, bar_ld, sKm, d + head * 64, x0, y0, b);
    tma_load_4d(&tmKV, bar_ld, sV, 2 * d + head * 64, x0, y0, b);
  }
  if (warp == 0) tmem_alloc<kW14TmemCols>(tmem_ptr);
  // the 12 pad keys: K rows zero (the mask column does the rest), V rows zero (0 x garbage could be NaN)
  if (tid < 128) {
    for (int i = tid; i < 2 * 12 * 8; i += 128) {
      const int which = i / 96, r = kW14N + (i % 96) / 8, c = i & 7;
      st_shared_v4((which ? sV : sKm) + r * 128 + ((c ^ (r & 7)) << 4), 0u, 0u, 0u, 0u);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 4) {
    // =============================================================== issuer
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64);
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kW14NP);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    mbar_wait(bar_fix, 0);                   // implies bar_ld: the query threads arrive after the loads (and their fix-ups)
    mbar_wait(bar_r, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t dq = umma_desc_kmajor_sw128(sQm), dh = umma_desc_kmajor_sw128(sQx), dl = umma_desc_kmajor_sw128(sQx + 8192);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_S, dq + 2 * k, dh + 2 * k, idesc_t, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_S, dq + 2 * k, dl + 2 * k, idesc_t, 1u);
      umma_commit(bar_t);
    }
    __syncwarp();
    mbar_wait(bar_kx, 0);
    mbar_wait(bar_qx, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t dq = umma_desc_kmajor_sw128(sQm), dk = umma_desc_kmajor_sw128(sKm);
      const uint64_t dqx = umma_desc_kmajor_sw128(sQx), dkx = umma_desc_kmajor_sw128(sKx);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_S, dqx + 2 * k, dkx + 2 * k, idesc_s, 1u);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t dv = w14_desc_mnmajor_sw128(sV);
#pragma unroll
      for (int k = 0; k < kW14NP / 16; ++k) umma_ts(tmem_O, tmem_S + k * 8, dv + static_cast<uint64_t>(k * 128), idesc_o, k != 0);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // =============================================================== query rows
    const int r = tid;                                   // row of the tile = TMEM lane
    const int t = kW14Half * half + r;                   // token of the window (valid: r < 98)
    const int ty = t / kW14, tx = t - ty * kW14;         // = (qh, qw)
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    mbar_wait(bar_ld, 0);
    // pad tokens (beyond the image): q / k / v = bf16(bias)
    const bool edge = (x0 + kW14 > p.gw) || (y0 + kW14 > p.gh);
    if (edge) {
      auto fix_row = [&](uint32_t tile, int row, int tok, const float* bias) {
        const int yy = tok / kW14, xx = tok - yy * kW14;
        if (y0 + yy < p.gh && x0 + xx < p.gw) return;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 f0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c), f1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c + 1);
          st_shared_v4(tile + row * 128 + ((c ^ (row & 7)) << 4), pack_bf16x2(f0.x, f0.y), pack_bf16x2(f0.z, f0.w), pack_bf16x2(f1.x, f1.y),
                       pack_bf16x2(f1.z, f1.w));
        }
      };
      if (r < kW14Half) fix_row(sQm, r, t, p.qkv_bias + head * 64);
      for (int k = r; k < kW14N; k += 128) {
        fix_row(sKm, k, k, p.qkv_bias + d + head * 64);
        fix_row(sV, k, k, p.qkv_bias + 2 * d + head * 64);
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(bar_fix);
    mbar_wait(bar_t, 0);
    tc_fence_after();
    float tt[64];
    {
      uint32_t u[32];
      tmem_ld_32x32b_x32(tmem_S + lane_sel, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) tt[i] = __uint_as_float(u[i]);
      tmem_ld_32x32b_x32(tmem_S + lane_sel + 32, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) tt[32 + i] = __uint_as_float(u[i]);
    }
    {
      float rh[14], rw[14];
      w14_pick<0>(tt, ty < kW14 ? ty : 0, rh);
      w14_pick<27>(tt, tx, rw);
      uint32_t e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = 0u;
      // columns: [0,14) h hi  [14,28) w hi  [28,42) h lo  [42,56) w lo  56: 1.0
      float hi[28], lo[28];
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        const float vh = rh[k] * p.inv_scale, vw = rw[k] * p.inv_scale;
        hi[k] = __bfloat162float(__float2bfloat16_rn(vh));
        hi[14 + k] = __bfloat162float(__float2bfloat16_rn(vw));
        lo[k] = vh - hi[k];
        lo[14 + k] = vw - hi[14 + k];
      }
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        e[i] = pack_bf16x2(hi[2 * i], hi[2 * i + 1]);
        e[14 + i] = pack_bf16x2(lo[2 * i], lo[2 * i + 1]);
      }
      e[28] = pack_bf16x2(1.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) st_shared_v4(sQx + r * 128 + ((c ^ (r & 7)) << 4), e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_qx);

    // ---- softmax over the 208 columns of this row
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const uint32_t tS = tmem_S + lane_sel;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t u[32];
      tmem_ld_32x32b_x32(tS + 32 * c, u);
      tmem_ld_wait();
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(u[i]));
        m1 = fmaxf(m1, __uint_as_float(u[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(u[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(u[i + 3]));
      }
      mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    {
      uint32_t u[16];
      tmem_ld_32x32b_x16(tS + 192, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(u[i]));
    }
    const float negm = -mx * p.scale_log2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t u[32], w[16];
      tmem_ld_32x32b_x32(tS + 32 * c, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = w14_ex2(fmaf(__uint_as_float(u[i]), p.scale_log2, negm));
        const float p1 = w14_ex2(fmaf(__uint_as_float(u[i + 1]), p.scale_log2, negm));
        l0 += p0;
        l1 += p1;
        w[i >> 1] = pack_bf16x2(p0, p1);
      }
      tmem_st_32x32b_x16(tS + 16 * c, w);               // P over the S' columns already consumed
    }
    {
      uint32_t u[16], w[8];
      tmem_ld_32x32b_x16(tS + 192, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = w14_ex2(fmaf(__uint_as_float(u[i]), p.scale_log2, negm));
        const float p1 = w14_ex2(fmaf(__uint_as_float(u[i + 1]), p.scale_log2, negm));
        l0 += p0;
        l1 += p1;
        w[i >> 1] = pack_bf16x2(p0, p1);
      }
      tmem_st_32x32b_x8(tS + 96, w);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar_p);
    const float inv = 1.f / (l0 + l1);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    const bool valid = r < kW14Half && y0 + ty < p.gh && x0 + tx < p.gw;
    __nv_bfloat16* orow = p.out + ((static_cast<int64_t>(b) * p.gh + (y0 + ty)) * p.gw + (x0 + tx)) * p.ld_out + head * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t u[32];
      tmem_ld_32x32b_x32(tmem_O + lane_sel + 32 * c, u);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(u[i]) * inv, __uint_as_float(u[i + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(u[i + 2]) * inv, __uint_as_float(u[i + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(u[i + 4]) * inv, __uint_as_float(u[i + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(u[i + 6]) * inv, __uint_as_float(u[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + 32 * c + i) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kW14TmemCols>(tmem_base);
  }
}

int launch_attn_win14_tc(const void* qkv, int64_t ld_qkv, const float* qkv_bias, const void* rcat_hi, const void* rcat_lo, void* out,
                         int64_t ld_out, int B, int gh, int gw, int heads, float scale, cudaStream_t s) {
  const int nwh = (gh + kW14 - 1) / kW14, nww = (gw + kW14 - 1) / kW14;
  static DeviceFlags configured;
  if (!configured.current()) {
    cudaError_t e = cudaFuncSetAttribute(attn_win14_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW14Smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_win14_tc_kernel)");
    w14_init_kx_kernel<<<(kW14NP * 64 + 255) / 256, 256, 0, s>>>();
    count_launch();
    VDR_CHECK_LAUNCH("w14_init_kx_kernel");
    configured.current() = true;
  }
  void* kx = nullptr;
  cudaError_t e = cudaGetSymbolAddress(&kx, g_w14_kx);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetSymbolAddress(g_w14_kx)");
  const uint64_t cols = 3ull * heads * 64;
  const uint64_t dims[4] = {cols, (uint64_t)gw, (uint64_t)gh, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)ld_qkv * 2, (uint64_t)gw * ld_qkv * 2, (uint64_t)gh * gw * ld_qkv * 2};
  const uint32_t box_q[4] = {64, kW14, 7, 1}, box_kv[4] = {64, kW14, kW14, 1};
  CUtensorMap tmQ, tmKV, tmRhi, tmRlo, tmKx;
  int rc = make_tmap_nd_bf16(&tmQ, qkv, 4, dims, strides, box_q);
  if (rc == VDR_OK) rc = make_tmap_nd_bf16(&tmKV, qkv, 4, dims, strides, box_kv);
  if (rc == VDR_OK) rc = make_tmap_2d_bf16(&tmRhi, rcat_hi, 2 * (2 * kW14 - 1), 64, 64, 64, 64);
  if (rc == VDR_OK) rc = make_tmap_2d_bf16(&tmRlo, rcat_lo, 2 * (2 * kW14 - 1), 64, 64, 64, 64);
  if (rc == VDR_OK) rc = make_tmap_2d_bf16(&tmKx, kx, kW14NP, 64, 64, kW14NP, 64);
  if (rc != VDR_OK) return rc;
  Win14Params p;
  p.qkv_bias = qkv_bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.gh = gh; p.gw = gw; p.nwh = nwh; p.nww = nww; p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.inv_scale = 1.f / scale;
  attn_win14_tc_kernel<<<dim3(2, heads, (unsigned)(B * nwh * nww)), kW14Threads, kW14Smem, s>>>(tmQ, tmKV, tmRhi, tmRlo, tmKx, p);
  count_launch();
  VDR_CHECK_LAUNCH("attn_win14_tc_kernel");
  return VDR_OK;
}

}  // namespace vdr
