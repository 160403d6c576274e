// the pieces of the MedSAM / SAM ViT-B image encoder (the backbone the reference loads by
// default, src/tfds_dense_descriptor.py:104,123; code in the un-vendored segment_anything package) that the plain-ViT path
// does not have: 14x14 window partition with zero padding, decomposed relative-position bias, attention with that bias,
// and the 3x3 convolution of the neck as an im2col over token-major maps.  The attention runs on mma.sync (bf16 m16n8k16,
// fp32 accumulation; cp.async + ldmatrix feed) with the bias built and added on chip; the GEMMs either side are the tcgen05 kernel.
#include "common.cuh"

namespace vdr {

// ------------------------------------------------------------------------------------------------ window (un)partition
// to_windows: dst row (b, wy, wx, ty, tx) <- src row (b, wy*ws+ty, wx*ws+tx), zeros outside H x W (the padding is applied to
//   the NORMALISED tokens, so pad tokens are exact zeros and take part in the window's softmax with q = k = v = bias);
// else:       dst row (b, y, x) <- src row (b, y/ws, x/ws, y%ws, x%ws)   (pad rows are dropped).
// One warp per destination row, 16-byte vectors.
__global__ void __launch_bounds__(256)
window_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst, int64_t ld_dst,
                   int H, int W, int ws, int nwh, int nww, int d, int64_t dst_rows, int to_windows) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int wt = ws * ws;
  for (int64_t r = warp0; r < dst_rows; r += nwarps) {
    int64_t srow = -1;
    if (to_windows) {
      const int64_t win = r / wt;
      const int t = static_cast<int>(r - win * wt);
      const int64_t b = win / (nwh * nww);
      const int wi = static_cast<int>(win - b * (nwh * nww));
      const int y = (wi / nww) * ws + t / ws, x = (wi % nww) * ws + t % ws;
      if (y < H && x < W) srow = (b * H + y) * W + x;
    } else {
      const int64_t b = r / ((int64_t)H * W);
      const int rem = static_cast<int>(r - b * H * W);
      const int y = rem / W, x = rem - y * W;
      srow = ((b * nwh + y / ws) * nww + x / ws) * wt + (y % ws) * ws + x % ws;
    }
    uint4* o = reinterpret_cast<uint4*>(dst + r * ld_dst);
    if (srow >= 0) {
      const uint4* s = reinterpret_cast<const uint4*>(src + srow * ld_src);
      for (int v = lane; v < (d >> 3); v += 32) o[v] = __ldg(s + v);
    } else {
      for (int v = lane; v < (d >> 3); v += 32) o[v] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention + bias
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// cp.async / ldmatrix wrappers (sm_80+ forms; the data path of this kernel is plain LDGSTS + LDSM + HMMA)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

constexpr int kRpBQ = 128, kRpBK = 64, kRpPitch = 72;   // 72 bf16 = 144-byte rows: ldmatrix's 8 row segments hit 32 distinct banks
constexpr int kRpTile = kRpBK * kRpPitch;               // elements of one K / V tile in shared memory

// acc[nt] (16 query rows x 8 tile rows) += Q (A fragments, 16 x 64) * T^T for the 64 rows of a K-major bf16 tile in smem.
// One ldmatrix.x4 feeds two n-tiles of one k-step, so consecutive MMAs never accumulate into the same registers
// (a dependent HMMA chain stalls for the full pipe latency).
// `rows` (warp-uniform) = tile rows that hold data: 16-row groups beyond it are skipped (their accumulators stay 0).
// `need` (warp-uniform bit mask, bit p = the 16-row group p is wanted): further groups to skip.
__device__ __forceinline__ void qk_tile_mma(float (&acc)[8][4], const uint32_t (&qa)[4][4], uint32_t tile_smem, int lane, int rows = kRpBK,
                                            uint32_t need = 0xfu) {
  const uint32_t lane_off = (((lane & 7) + (lane >> 4) * 8) * kRpPitch + ((lane >> 3) & 1) * 8) * 2;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int ntp = 0; ntp < 4; ++ntp) {
      if (ntp * 16 >= rows || !((need >> ntp) & 1u)) continue;
      uint32_t b[4];     // b0/b1 of n-tile 2*ntp, b0/b1 of n-tile 2*ntp + 1, k-step ks
      ldmatrix_x4(b, tile_smem + lane_off + (ntp * 16 * kRpPitch + ks * 16) * 2);
      mma_bf16_16816(acc[2 * ntp], qa[ks], b[0], b[1]);
      mma_bf16_16816(acc[2 * ntp + 1], qa[ks], b[2], b[3]);
    }
  }
}

// softmax(q k^T * scale + rel_h[q, kh] + rel_w[q, kw]) v   for one (128-query block, head, image-or-window).
//   8 warps x 16 query rows; 64-key tiles, cp.async double buffering; K fragments by ldmatrix, V fragments by ldmatrix.trans;
//   online softmax in the log2 domain; P rounded to bf16 for the second MMA.
//   The decomposed relative-position terms are built in the prologue by the same MMA path: the block's queries times the
//   concatenated tables [rel_pos_h ; rel_pos_w] (rows j), each table split into bf16 hi + lo parts (fp32-class accuracy),
//   and the entries a query needs -- kh = qh + Sh-1 - j, kw = qw + Sw-1 - j' -- are kept in shared memory.
//   kRowTiles: Sw == 64 and N % 64 == 0 (the 64 x 64 token grid of a 1024^2 image): a key tile is exactly one grid row, so
//   kh = tile index, kw = column -- no index arithmetic, no key mask, the rel_w terms come as aligned float2 loads.
template <bool kRowTiles>
__global__ void __launch_bounds__(256, 2)
attn_relpos_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ rcat_hi,
                   const __nv_bfloat16* __restrict__ rcat_lo, __nv_bfloat16* __restrict__ out, int64_t ld_out, int N, int heads,
                   int Sh, int Sw, int RP, uint32_t magic_sw, float scale_log2e) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Kb = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [2][64 keys][72]
  __nv_bfloat16* Vb = Kb + 2 * kRpTile;                                 // [2][64 keys][72]
  float* relS = reinterpret_cast<float*>(Vb + 2 * kRpTile);            // [128 query rows][RP >= Sh + Sw], times log2(e); RP % 32 == 8
  const int qb = blockIdx.x, h = blockIdx.y, bw = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = heads * 64;
  const int q0 = qb * kRpBQ;
  const __nv_bfloat16* base = qkv + (int64_t)bw * N * ld + h * 64;     // q at +0, k at +d, v at +2d
  const uint32_t kb_s = smem_u32(Kb), vb_s = smem_u32(Vb);
  // Q fragments (A operand, 16 rows x 64): rows g and g + 8 of this warp's 16
  const int rl0 = warp * 16 + g, rl1 = rl0 + 8;
  const int r0 = q0 + rl0, r1 = q0 + rl1;
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + ks * 16 + 2 * t);
    const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + ks * 16 + 2 * t);
    qa[ks][0] = r0 < N ? __ldg(p0) : 0u;
    qa[ks][1] = r1 < N ? __ldg(p1) : 0u;
    qa[ks][2] = r0 < N ? __ldg(p0 + 4) : 0u;
    qa[ks][3] = r1 < N ? __ldg(p1 + 4) : 0u;
  }
  const bool warp_active = q0 + warp * 16 < N;      // a warp whose 16 query rows lie beyond N only helps with the loads and barriers
  float* rel0 = relS + rl0 * RP;
  float* rel1 = relS + rl1 * RP;
  const int WO = (Sh + 1) & ~1;                // rel_w terms start at an even column: float2 loads in the row-tile variant

  // ---- prologue: the block's bias entries.  Table chunks of 64 rows go through the K buffers (hi -> Kb[0], lo -> Kb[1]).
  {
    const int RH = 2 * Sh - 1, RT = RH + 2 * Sw - 1;
    const int qh0 = r0 / Sw, qh1 = r1 / Sw;
    const int offh0 = qh0 + Sh - 1, offw0 = r0 - qh0 * Sw + Sw - 1;
    const int offh1 = qh1 + Sh - 1, offw1 = r1 - qh1 * Sw + Sw - 1;
    for (int c0 = 0; c0 < RT; c0 += 64) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * 256, row = idx >> 3, seg = idx & 7, j = c0 + row;
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
        if (j < RT) {
          hi = __ldg(reinterpret_cast<const uint4*>(rcat_hi + (int64_t)j * 64 + seg * 8));
          lo = __ldg(reinterpret_cast<const uint4*>(rcat_lo + (int64_t)j * 64 + seg * 8));
        }
        *reinterpret_cast<uint4*>(Kb + row * kRpPitch + seg * 8) = hi;
        *reinterpret_cast<uint4*>(Kb + kRpTile + row * kRpPitch + seg * 8) = lo;
      }
      __syncthreads();
      if (!warp_active) continue;
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
      qk_tile_mma(acc, qa, kb_s + kRpTile * 2, lane, RT - c0);      // low parts first, then the high parts on top
      qk_tile_mma(acc, qa, kb_s, lane, RT - c0);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = c0 + nt * 8 + 2 * t + e;
          if (j < RH) {
            const int kh0 = offh0 - j, kh1 = offh1 - j;
            if (r0 < N && kh0 >= 0 && kh0 < Sh) rel0[kh0] = acc[nt][e] * 1.4426950408889634f;
            if (r1 < N && kh1 >= 0 && kh1 < Sh) rel1[kh1] = acc[nt][2 + e] * 1.4426950408889634f;
          } else if (j < RT) {
            const int kw0 = offw0 - (j - RH), kw1 = offw1 - (j - RH);
            if (r0 < N && kw0 >= 0 && kw0 < Sw) rel0[WO + kw0] = acc[nt][e] * 1.4426950408889634f;
            if (r1 < N && kw1 >= 0 && kw1 < Sw) rel1[WO + kw1] = acc[nt][2 + e] * 1.4426950408889634f;
          }
        }
      }
    }
    __syncthreads();      // table chunks consumed: the K buffers may now receive key tiles; relS visible
  }

  auto issue_tile = [&](int kt) {
    const int k0 = kt * kRpBK, buf = kt & 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, row = idx >> 3, seg = idx & 7, key = k0 + row;
      const int bytes = key < N ? 16 : 0;                                  // out-of-range keys are zero-filled
      const __nv_bfloat16* src = base + (int64_t)(key < N ? key : N - 1) * ld + seg * 8;
      const uint32_t off = (buf * kRpTile + row * kRpPitch + seg * 8) * 2;
      cp_async16(kb_s + off, src + d, bytes);
      cp_async16(vb_s + off, src + 2 * d, bytes);
    }
    cp_async_commit();
  };

  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int n_tiles = (N + kRpBK - 1) / kRpBK;
  issue_tile(0);
  for (int kt = 0; kt < n_tiles; ++kt) {
    const int k0 = kt * kRpBK, buf = kt & 1;
    if (kt + 1 < n_tiles) {
      issue_tile(kt + 1);          // into the buffer the previous iteration finished reading (barrier at its end)
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (warp_active) {
    const int rem = kRowTiles ? kRpBK : N - k0;      // keys in this tile (>= 64 except on a ragged last tile)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
    qk_tile_mma(s, qa, kb_s + buf * kRpTile * 2, lane, rem);
    // scale + bias (+ key mask on a ragged last tile); row maxima
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (kRowTiles) {
      const float bh0 = rel0[kt], bh1 = rel1[kt];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 w0 = *reinterpret_cast<const float2*>(rel0 + WO + nt * 8 + 2 * t);
        const float2 w1 = *reinterpret_cast<const float2*>(rel1 + WO + nt * 8 + 2 * t);
        s[nt][0] = fmaf(s[nt][0], scale_log2e, bh0 + w0.x);
        s[nt][1] = fmaf(s[nt][1], scale_log2e, bh0 + w0.y);
        s[nt][2] = fmaf(s[nt][2], scale_log2e, bh1 + w1.x);
        s[nt][3] = fmaf(s[nt][3], scale_log2e, bh1 + w1.y);
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    } else {
      const bool ragged = k0 + kRpBK > N;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = k0 + nt * 8 + 2 * t + e;
          if (!ragged || key < N) {
            const int kh = Sw == 1 ? key : (int)__umulhi((uint32_t)key, magic_sw);
            const int kw = key - kh * Sw;
            s[nt][e] = fmaf(s[nt][e], scale_log2e, rel0[kh] + rel0[WO + kw]);
            s[nt][2 + e] = fmaf(s[nt][2 + e], scale_log2e, rel1[kh] + rel1[WO + kw]);
          } else {
            s[nt][e] = -INFINITY;
            s[nt][2 + e] = -INFINITY;
          }
          mx0 = fmaxf(mx0, s[nt][e]);
          mx1 = fmaxf(mx1, s[nt][2 + e]);
        }
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);      // finite: key k0 < N is valid in every tile
    const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);   // ex2(-inf) = 0 on the first tile
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2_approx(s[nt][0] - mn0); s[nt][1] = ex2_approx(s[nt][1] - mn0);
      s[nt][2] = ex2_approx(s[nt][2] - mn1); s[nt][3] = ex2_approx(s[nt][3] - mn1);
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    if (__any_sync(0xffffffffu, c0 != 1.f || c1 != 1.f)) {      // the running maxima settle after the first tiles
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
    }
    l0 = l0 * c0 + ps0;                               // per-thread partial sums; reduced over the quad at the end
    l1 = l1 * c1 + ps1;
    // O += P V : P (16 x 64 keys) from the S accumulators; V (keys x d, row-major tile) through ldmatrix.trans
    const uint32_t v_lane = vb_s + (buf * kRpTile + ((lane & 7) + ((lane >> 3) & 1) * 8) * kRpPitch + (lane >> 4) * 8) * 2;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 >= rem) continue;                  // P is zero beyond the last key
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int ntp = 0; ntp < 4; ++ntp) {
        uint32_t b[4];   // b0/b1 of d-tile 2*ntp, b0/b1 of d-tile 2*ntp + 1
        ldmatrix_x4_trans(b, v_lane + (kk * 16 * kRpPitch + ntp * 16) * 2);
        mma_bf16_16816(o[2 * ntp], pa, b[0], b[1]);
        mma_bf16_16816(o[2 * ntp + 1], pa, b[2], b[3]);
      }
    }
    }
    __syncthreads();      // everyone is done with this tile's buffers before the next iteration refills them
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __nv_bfloat16* ob = out + (int64_t)bw * N * ld_out + h * 64;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (r0 < N) *reinterpret_cast<uint32_t*>(ob + (int64_t)r0 * ld_out + nt * 8 + 2 * t) = pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < N) *reinterpret_cast<uint32_t*>(ob + (int64_t)r1 * ld_out + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------ resident-key variant
// Small extents (N <= 208 tokens, 2*Sh-1 + 2*Sw-1 <= 64 table rows: SAM's 14 x 14 windows): two CTAs per (window, head), 7 warps
// x 16 query rows each (rows 0..111 and 112..207), ALL keys / values and both table parts resident in shared memory after one
// cp.async wave -- a single CTA barrier instead of two per key tile, no query rows beyond the 16-row granule, keys processed as
// a 112- and a 96-wide tile (7 + 6 MMA k-steps) instead of four 64-wide ones.  Two CTAs fit an SM (registers and 96 KB of
// shared memory each), so one CTA's load wave overlaps the other's arithmetic.
constexpr int kWinRows = 208, kWinQ = 112, kWinWarps = 7, kWinThreads = kWinWarps * 32;

template <int NT>   // 8-key n-tiles in this key tile (even)
__device__ __forceinline__ void win_key_tile(float (&o)[8][4], float& m0, float& m1, float& l0, float& l1, const uint32_t (&qa)[4][4],
                                             uint32_t k_s, uint32_t v_s, int k0, int N, const float* rel0, const float* rel1, int WO, int Sw,
                                             uint32_t magic_sw, float scale_log2e, int lane) {
  const int t = lane & 3;
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
  const uint32_t k_lane = k_s + ((k0 + (lane & 7) + (lane >> 4) * 8) * kRpPitch + ((lane >> 3) & 1) * 8) * 2;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int ntp = 0; ntp < NT / 2; ++ntp) {
      uint32_t b[4];
      ldmatrix_x4(b, k_lane + (ntp * 16 * kRpPitch + ks * 16) * 2);
      mma_bf16_16816(s[2 * ntp], qa[ks], b[0], b[1]);
      mma_bf16_16816(s[2 * ntp + 1], qa[ks], b[2], b[3]);
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = k0 + nt * 8 + 2 * t + e;
      if (key < N) {
        const int kh = Sw == 1 ? key : (int)__umulhi((uint32_t)key, magic_sw);
        const int kw = key - kh * Sw;
        s[nt][e] = fmaf(s[nt][e], scale_log2e, rel0[kh] + rel0[WO + kw]);
        s[nt][2 + e] = fmaf(s[nt][2 + e], scale_log2e, rel1[kh] + rel1[WO + kw]);
      } else {
        s[nt][e] = -INFINITY;
        s[nt][2 + e] = -INFINITY;
      }
      mx0 = fmaxf(mx0, s[nt][e]);
      mx1 = fmaxf(mx1, s[nt][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);      // finite: key k0 < N is valid in every tile the caller passes
  const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);
  m0 = mn0; m1 = mn1;
  float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = ex2_approx(s[nt][0] - mn0); s[nt][1] = ex2_approx(s[nt][1] - mn0);
    s[nt][2] = ex2_approx(s[nt][2] - mn1); s[nt][3] = ex2_approx(s[nt][3] - mn1);
    ps0 += s[nt][0] + s[nt][1];
    ps1 += s[nt][2] + s[nt][3];
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
  l0 = l0 * c0 + ps0;
  l1 = l1 * c1 + ps1;
  const uint32_t v_lane = v_s + ((k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kRpPitch + (lane >> 4) * 8) * 2;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int ntp = 0; ntp < 4; ++ntp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, v_lane + (kk * 16 * kRpPitch + ntp * 16) * 2);
      mma_bf16_16816(o[2 * ntp], pa, b[0], b[1]);
      mma_bf16_16816(o[2 * ntp + 1], pa, b[2], b[3]);
    }
  }
}

// kMapped: qkv / out hold the UN-partitioned token rows of B images of gh x gw tokens; window bw = (image, wy, wx) reads its
// Sh x Sw tokens in place (window_partition without the copy) and pad tokens -- positions beyond the image, whose normalised
// input is zero, so that q = k = v = the qkv bias -- are synthesised from `qkv_bias`; outputs of pad queries are dropped
// (window_unpartition without the copy).
struct WinMap {
  int gh, gw, nwh, nww;
  const float* qkv_bias;      // (3 * heads * 64) f32: the UNFOLDED bias of the qkv Linear
};

template <bool kMapped>
__global__ void __launch_bounds__(kWinThreads, 2)
attn_relpos_win_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ rcat_hi,
                       const __nv_bfloat16* __restrict__ rcat_lo, __nv_bfloat16* __restrict__ out, int64_t ld_out, int N, int heads,
                       int Sh, int Sw, int RP, uint32_t magic_sw, float scale_log2e, WinMap wm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [208 keys][72]
  __nv_bfloat16* Vs = Ks + kWinRows * kRpPitch;                         // [208 keys][72]
  __nv_bfloat16* Rs = Vs + kWinRows * kRpPitch;                         // [2][64 table rows][72]: hi, lo
  float* relS = reinterpret_cast<float*>(Rs + 2 * kRpTile);            // [112 query rows][RP], times log2(e)
  const int h = blockIdx.x, bw = blockIdx.y, q0 = blockIdx.z * kWinQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = heads * 64;
  const __nv_bfloat16* base = qkv + (kMapped ? 0 : (int64_t)bw * N * ld) + h * 64;
  const uint32_t ks_s = smem_u32(Ks), vs_s = smem_u32(Vs), rs_s = smem_u32(Rs);
  // token r of this window -> its row in qkv / out (kMapped: -1 for a pad token)
  int64_t img_row0 = 0;
  int y0 = 0, x0 = 0;
  if (kMapped) {
    const int per_img = wm.nwh * wm.nww, bi = bw / per_img, wi = bw - bi * per_img;
    img_row0 = (int64_t)bi * wm.gh * wm.gw;
    y0 = (wi / wm.nww) * Sh;
    x0 = (wi % wm.nww) * Sw;
  }
  auto token_row = [&](int r) -> int64_t {
    if (!kMapped) return r;
    const int ty = Sw == 1 ? r : (int)__umulhi((uint32_t)r, magic_sw), tx = r - ty * Sw;
    const int y = y0 + ty, x = x0 + tx;
    return (y < wm.gh && x < wm.gw) ? img_row0 + (int64_t)y * wm.gw + x : -1;
  };
  const int RH = 2 * Sh - 1, RT = RH + 2 * Sw - 1;
  // one cp.async wave: K, V (rows beyond N zero-filled) and the two table parts (rows beyond RT zero-filled)
  for (int idx = tid; idx < kWinRows * 8; idx += kWinThreads) {
    const int row = idx >> 3, seg = idx & 7;
    const uint32_t off = (row * kRpPitch + seg * 8) * 2;
    const int64_t trow = row < N ? token_row(row) : 0;
    if (kMapped && trow < 0) {      // pad token: k = bias_k, v = bias_v (rounded to bf16 as the qkv GEMM would have stored them)
      const float* bk = wm.qkv_bias + d + h * 64 + seg * 8;
      const float4 k0 = __ldg(reinterpret_cast<const float4*>(bk)), k1 = __ldg(reinterpret_cast<const float4*>(bk) + 1);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(bk + d)), v1 = __ldg(reinterpret_cast<const float4*>(bk + d) + 1);
      *reinterpret_cast<uint4*>(Ks + row * kRpPitch + seg * 8) =
          make_uint4(pack_bf16x2(k0.x, k0.y), pack_bf16x2(k0.z, k0.w), pack_bf16x2(k1.x, k1.y), pack_bf16x2(k1.z, k1.w));
      *reinterpret_cast<uint4*>(Vs + row * kRpPitch + seg * 8) =
          make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
      continue;
    }
    const int bytes = row < N ? 16 : 0;
    const __nv_bfloat16* src = base + (kMapped ? trow : (int64_t)(row < N ? row : N - 1)) * ld + seg * 8;
    cp_async16(ks_s + off, src + d, bytes);
    cp_async16(vs_s + off, src + 2 * d, bytes);
  }
  for (int idx = tid; idx < 64 * 8; idx += kWinThreads) {
    const int row = idx >> 3, seg = idx & 7;
    const int bytes = row < RT ? 16 : 0;
    const int64_t soff = (int64_t)(row < RT ? row : RT - 1) * 64 + seg * 8;
    const uint32_t off = (row * kRpPitch + seg * 8) * 2;
    cp_async16(rs_s + off, rcat_hi + soff, bytes);
    cp_async16(rs_s + kRpTile * 2 + off, rcat_lo + soff, bytes);
  }
  cp_async_commit();
  const int rl0 = q0 + warp * 16 + g, rl1 = rl0 + 8;     // this thread's query rows (tokens of the window)
  const int64_t tr0 = rl0 < N ? token_row(rl0) : -1, tr1 = rl1 < N ? token_row(rl1) : -1;   // -1: no such query (beyond N or pad)
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base + (tr0 < 0 ? 0 : tr0) * ld + ks * 16 + 2 * t);
    const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base + (tr1 < 0 ? 0 : tr1) * ld + ks * 16 + 2 * t);
    qa[ks][0] = tr0 >= 0 ? __ldg(p0) : 0u;
    qa[ks][1] = tr1 >= 0 ? __ldg(p1) : 0u;
    qa[ks][2] = tr0 >= 0 ? __ldg(p0 + 4) : 0u;
    qa[ks][3] = tr1 >= 0 ? __ldg(p1 + 4) : 0u;
  }
  cp_async_wait<0>();
  __syncthreads();                                       // the only CTA barrier
  if (q0 + warp * 16 >= N) return;
  float* rel0 = relS + (warp * 16 + g) * RP;
  float* rel1 = rel0 + 8 * RP;
  const int WO = (Sh + 1) & ~1;
  {
    const int qh0 = rl0 / Sw, qh1 = rl1 / Sw;
    const int offh0 = qh0 + Sh - 1, offw0 = rl0 - qh0 * Sw + Sw - 1;
    const int offh1 = qh1 + Sh - 1, offw1 = rl1 - qh1 * Sw + Sw - 1;
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
    qk_tile_mma(acc, qa, rs_s + kRpTile * 2, lane, RT);
    qk_tile_mma(acc, qa, rs_s, lane, RT);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        if (j < RH) {
          const int kh0 = offh0 - j, kh1 = offh1 - j;
          if (rl0 < N && kh0 >= 0 && kh0 < Sh) rel0[kh0] = acc[nt][e] * 1.4426950408889634f;
          if (rl1 < N && kh1 >= 0 && kh1 < Sh) rel1[kh1] = acc[nt][2 + e] * 1.4426950408889634f;
        } else if (j < RT) {
          const int kw0 = offw0 - (j - RH), kw1 = offw1 - (j - RH);
          if (rl0 < N && kw0 >= 0 && kw0 < Sw) rel0[WO + kw0] = acc[nt][e] * 1.4426950408889634f;
          if (rl1 < N && kw1 >= 0 && kw1 < Sw) rel1[WO + kw1] = acc[nt][2 + e] * 1.4426950408889634f;
        }
      }
    }
    __syncwarp();                                        // a warp reads only the bias rows it wrote itself
  }
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  win_key_tile<14>(o, m0, m1, l0, l1, qa, ks_s, vs_s, 0, N, rel0, rel1, WO, Sw, magic_sw, scale_log2e, lane);
  if (N > 112) win_key_tile<12>(o, m0, m1, l0, l1, qa, ks_s, vs_s, 112, N, rel0, rel1, WO, Sw, magic_sw, scale_log2e, lane);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __nv_bfloat16* ob = out + (kMapped ? 0 : (int64_t)bw * N * ld_out) + h * 64;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (tr0 >= 0) *reinterpret_cast<uint32_t*>(ob + tr0 * ld_out + nt * 8 + 2 * t) = pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
    if (tr1 >= 0) *reinterpret_cast<uint32_t*>(ob + tr1 * ld_out + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------ rel-pos tables
// rel[((bw*heads + h)*N + q)*(Sh+Sw) + j] = out_scale * q_vec . Rh[qh - j + Sh - 1]        (j <  Sh)
//                                         = out_scale * q_vec . Rw[qw - (j-Sh) + Sw - 1]   (j >= Sh)
// with q_vec the UNSCALED query of token q = qh*Sw + qw, head h (segment_anything add_decomposed_rel_pos).  Same MMA path as
// the prologue of attn_relpos_kernel (queries x [rel_pos_h ; rel_pos_w], bf16 hi + lo parts), written to global memory: the
// tcgen05 attention (vdr_flash_attn_relpos_fwd) reads its bias terms from this table.
__global__ void __launch_bounds__(256, 3)
relpos_table_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, const __nv_bfloat16* __restrict__ rcat_hi,
                    const __nv_bfloat16* __restrict__ rcat_lo, float* __restrict__ rel, int N, int heads, int Sh, int Sw, float out_scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Kb = reinterpret_cast<__nv_bfloat16*>(smem_raw);       // [4 sub-chunks of 64 table rows][hi, lo][64][72]
  const int qb = blockIdx.x, h = blockIdx.y, bw = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qb * kRpBQ;
  const __nv_bfloat16* base = qkv + (int64_t)bw * N * ld + h * 64;
  const uint32_t kb_s = smem_u32(Kb);
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base + (int64_t)r0 * ld + ks * 16 + 2 * t);
    const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base + (int64_t)r1 * ld + ks * 16 + 2 * t);
    qa[ks][0] = r0 < N ? __ldg(p0) : 0u;
    qa[ks][1] = r1 < N ? __ldg(p1) : 0u;
    qa[ks][2] = r0 < N ? __ldg(p0 + 4) : 0u;
    qa[ks][3] = r1 < N ? __ldg(p1 + 4) : 0u;
  }
  const int R = Sh + Sw, RH = 2 * Sh - 1, RT = RH + 2 * Sw - 1;
  float* o0 = rel + ((int64_t)(bw * heads + h) * N + r0) * R;
  float* o1 = rel + ((int64_t)(bw * heads + h) * N + r1) * R;
  const int qh0 = r0 / Sw, qh1 = r1 / Sw;
  const int offh0 = qh0 + Sh - 1, offw0 = r0 - qh0 * Sw + Sw - 1;
  const int offh1 = qh1 + Sh - 1, offw1 = r1 - qh1 * Sw + Sw - 1;
  // table rows this warp's 16 queries can select: rel_h rows [qh_min, qh_max + Sh - 1], rel_w rows RH + [qw_min, qw_max + Sw - 1]
  // (a query (qh, qw) reads rows qh + Sh-1 - kh and RH + qw + Sw-1 - kw, kh < Sh, kw < Sw); 16-row groups outside both are skipped
  const int rw0 = q0 + warp * 16, rw1 = (rw0 + 15 < N ? rw0 + 15 : N - 1);
  const int h_lo = rw0 / Sw, h_hi = rw1 / Sw + Sh - 1;
  const int w_lo = RH + (h_lo == rw1 / Sw ? rw0 - h_lo * Sw : 0), w_hi = RH + (h_lo == rw1 / Sw ? rw1 - h_lo * Sw : Sw - 1) + Sw - 1;
  // up to 256 table rows (four 64-row sub-chunks, both parts) land in shared memory in one cp.async wave behind one barrier:
  // the load -> barrier -> MMA -> store chain used to run once per 64 rows and left the kernel latency-bound
  for (int s0 = 0; s0 < RT; s0 += 256) {
    if (s0 > 0) __syncthreads();
    for (int idx = tid; idx < 256 * 8; idx += 256) {
      const int row = idx >> 3, seg = idx & 7, j = s0 + row;
      if (j - (j & 63) >= RT) break;                                   // whole sub-chunk beyond the tables (uniform per 64 rows)
      const int bytes = j < RT ? 16 : 0;
      const int64_t soff = (int64_t)(j < RT ? j : RT - 1) * 64 + seg * 8;
      const uint32_t off = ((row >> 6) * 2 * kRpTile + (row & 63) * kRpPitch + seg * 8) * 2;
      cp_async16(kb_s + off, rcat_hi + soff, bytes);
      cp_async16(kb_s + kRpTile * 2 + off, rcat_lo + soff, bytes);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  for (int c0 = s0; c0 < RT && c0 < s0 + 256; c0 += 64) {
    const uint32_t sub_s = kb_s + ((c0 - s0) >> 6) * 2 * kRpTile * 2;
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
    uint32_t need = 0u;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int g_lo = c0 + 16 * p, g_hi = g_lo + 15;
      if (rw0 < N && ((g_hi >= h_lo && g_lo <= h_hi) || (g_hi >= w_lo && g_lo <= w_hi))) need |= 1u << p;
    }
    if (need == 0u) continue;
    qk_tile_mma(acc, qa, sub_s + kRpTile * 2, lane, RT - c0, need);
    qk_tile_mma(acc, qa, sub_s, lane, RT - c0, need);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (!((need >> (nt >> 1)) & 1u)) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = c0 + nt * 8 + 2 * t + e;
        if (j < RH) {
          const int kh0 = offh0 - j, kh1 = offh1 - j;
          if (r0 < N && kh0 >= 0 && kh0 < Sh) o0[kh0] = acc[nt][e] * out_scale;
          if (r1 < N && kh1 >= 0 && kh1 < Sh) o1[kh1] = acc[nt][2 + e] * out_scale;
        } else if (j < RT) {
          const int kw0 = offw0 - (j - RH), kw1 = offw1 - (j - RH);
          if (r0 < N && kw0 >= 0 && kw0 < Sw) o0[Sh + kw0] = acc[nt][e] * out_scale;
          if (r1 < N && kw1 >= 0 && kw1 < Sw) o1[Sh + kw1] = acc[nt][2 + e] * out_scale;
        }
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------ 3x3 im2col (neck)
// A[(b,y,x), (ky*3+kx)*C + c] = X[(b, y+ky-1, x+kx-1), c], zero outside the map (Conv2d padding 1).  One 16-byte vector per thread.
__global__ void __launch_bounds__(256)
im2col3x3_tokens_kernel(const __nv_bfloat16* __restrict__ X, int64_t ldx, __nv_bfloat16* __restrict__ A, int64_t lda,
                        int H, int W, int C, int64_t total_vec) {
  const int vpt = C >> 3;                  // vectors per tap
  const int vpr = 9 * vpt;                 // vectors per output row
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total_vec; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = v / vpr;
    const int rem = static_cast<int>(v - m * vpr);
    const int tap = rem / vpt, cv = rem - tap * vpt;
    const int x = static_cast<int>(m % W);
    const int64_t tq = m / W;
    const int y = static_cast<int>(tq % H);
    const int64_t b = tq / H;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = __ldg(reinterpret_cast<const uint4*>(X + ((b * H + yy) * W + xx) * ldx) + cv);
    *reinterpret_cast<uint4*>(A + m * lda + (int64_t)tap * C + cv * 8) = val;
  }
}

}  // namespace vdr

// ================================================================================================ C ABI
extern "C" int vdr_window_rows(const void* src_bf16, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int B, int H, int W, int ws,
                               int d, int to_windows, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(src_bf16 && dst_bf16, VDR_EINVAL, "vdr_window_rows: null pointer");
  VDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && ws > 0 && d > 0 && d % 8 == 0, VDR_EINVAL,
                "vdr_window_rows: bad shape B=%d H=%d W=%d ws=%d d=%d (d must be a multiple of 8)", B, H, W, ws, d);
  VDR_CHECK_ARG(aligned16(src_bf16) && aligned16(dst_bf16) && ld_src % 8 == 0 && ld_dst % 8 == 0 && ld_src >= d && ld_dst >= d,
                VDR_EALIGN, "vdr_window_rows: pointers must be 16-byte aligned, leading dimensions multiples of 8 and >= d");
  const int nwh = (H + ws - 1) / ws, nww = (W + ws - 1) / ws;
  const int64_t dst_rows = to_windows ? (int64_t)B * nwh * nww * ws * ws : (int64_t)B * H * W;
  int64_t blocks = (dst_rows + 7) / 8;
  const int64_t max_blocks = (int64_t)num_sms() * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  window_rows_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src_bf16), ld_src, static_cast<__nv_bfloat16*>(dst_bf16), ld_dst, H, W, ws, nwh, nww, d,
      dst_rows, to_windows ? 1 : 0);
  count_launch();
  VDR_CHECK_LAUNCH("window_rows_kernel");
  return VDR_OK;
}

extern "C" int vdr_relpos_tables(const void* qkv_bf16, int64_t ld_qkv, const void* rcat_hi_bf16, const void* rcat_lo_bf16, float* rel,
                                 int BW, int Sh, int Sw, int heads, float out_scale, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(qkv_bf16 && rcat_hi_bf16 && rcat_lo_bf16 && rel, VDR_EINVAL, "vdr_relpos_tables: null pointer");
  VDR_CHECK_ARG(BW > 0 && BW <= 65535 && Sh > 0 && Sw > 0 && heads > 0 && heads <= 65535 && (int64_t)Sh * Sw < 65536, VDR_EINVAL,
                "vdr_relpos_tables: bad shape BW=%d Sh=%d Sw=%d heads=%d", BW, Sh, Sw, heads);
  VDR_CHECK_ARG(ld_qkv >= 3LL * heads * 64 && ld_qkv % 8 == 0 && aligned16(qkv_bf16) && aligned16(rcat_hi_bf16) && aligned16(rcat_lo_bf16),
                VDR_EALIGN, "vdr_relpos_tables: qkv must be (rows, >= 3*heads*64) with ld %% 8 == 0; tables 16-byte aligned");
  const int N = Sh * Sw;
  dim3 grid((N + kRpBQ - 1) / kRpBQ, heads, BW);
  constexpr int kTableSmem = 4 * 2 * kRpTile * (int)sizeof(__nv_bfloat16);      // 73,728 B
  static DeviceFlags configured;
  if (!configured.current()) {
    cudaError_t e = cudaFuncSetAttribute(relpos_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTableSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(relpos_table_kernel)");
    configured.current() = true;
  }
  relpos_table_kernel<<<grid, 256, kTableSmem, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv_bf16), ld_qkv, static_cast<const __nv_bfloat16*>(rcat_hi_bf16),
      static_cast<const __nv_bfloat16*>(rcat_lo_bf16), rel, N, heads, Sh, Sw, out_scale);
  count_launch();
  VDR_CHECK_LAUNCH("relpos_table_kernel");
  return VDR_OK;
}

extern "C" int vdr_attn_relpos_fwd(const void* qkv_bf16, int64_t ld_qkv, const void* rcat_hi_bf16, const void* rcat_lo_bf16,
                                   void* out_bf16, int64_t ld_out, int BW, int Sh, int Sw, int heads, float scale, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(qkv_bf16 && rcat_hi_bf16 && rcat_lo_bf16 && out_bf16, VDR_EINVAL, "vdr_attn_relpos_fwd: null pointer");
  VDR_CHECK_ARG(BW > 0 && BW <= 65535 && Sh > 0 && Sw > 0 && heads > 0 && heads <= 65535 && (int64_t)Sh * Sw < 65536, VDR_EINVAL,
                "vdr_attn_relpos_fwd: bad shape BW=%d Sh=%d Sw=%d heads=%d (Sh*Sw < 65536)", BW, Sh, Sw, heads);
  VDR_CHECK_ARG(ld_qkv >= 3LL * heads * 64 && ld_qkv % 8 == 0 && ld_out >= heads * 64LL && ld_out % 8 == 0 && aligned16(qkv_bf16) &&
                    aligned16(out_bf16) && aligned16(rcat_hi_bf16) && aligned16(rcat_lo_bf16),
                VDR_EALIGN, "vdr_attn_relpos_fwd: qkv (rows, >= 3*heads*64) / out (rows, >= heads*64), ld %% 8 == 0, 16-byte aligned");
  const int N = Sh * Sw;
  const int rp0 = ((Sh + 1) & ~1) + Sw;                          // rel_h terms, pad to an even column, rel_w terms
  const int RP = rp0 + (8 - rp0 % 32 + 32) % 32;                 // row pitch of the bias tile: RP % 32 == 8 spreads a warp's rows over the banks
  const size_t smem = 4 * (size_t)kRpTile * sizeof(__nv_bfloat16) + (size_t)kRpBQ * RP * sizeof(float);
  VDR_CHECK_ARG(smem <= 200 * 1024, VDR_EINVAL, "vdr_attn_relpos_fwd: Sh + Sw = %d too large for the shared-memory bias tile", Sh + Sw);
  const uint32_t magic = Sw > 1 ? (uint32_t)((0x100000000ULL + (uint64_t)Sw - 1) / (uint64_t)Sw) : 0u;   // key / Sw = umulhi(key, magic), key < 2^16
  if (N <= kWinRows && 2 * Sh - 1 + 2 * Sw - 1 <= 64) {
    // small extents (the 14 x 14 windows): two CTAs per (window, head) with everything resident
    const size_t smem_w = (2 * (size_t)kWinRows * kRpPitch + 2 * (size_t)kRpTile) * sizeof(__nv_bfloat16) + (size_t)kWinQ * RP * sizeof(float);
    static DeviceFlags configured_w;
    if (!configured_w.current()) {
      cudaError_t e = cudaFuncSetAttribute(attn_relpos_win_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_relpos_win_kernel)");
      configured_w.current() = true;
    }
    attn_relpos_win_kernel<false><<<dim3(heads, BW, N > kWinQ ? 2 : 1), kWinThreads, smem_w, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv_bf16), ld_qkv, static_cast<const __nv_bfloat16*>(rcat_hi_bf16),
        static_cast<const __nv_bfloat16*>(rcat_lo_bf16), static_cast<__nv_bfloat16*>(out_bf16), ld_out, N, heads, Sh, Sw, RP, magic,
        scale * 1.4426950408889634f, WinMap{});
    count_launch();
    VDR_CHECK_LAUNCH("attn_relpos_win_kernel");
    return VDR_OK;
  }
  const bool row_tiles = Sw == 64 && N % 64 == 0;
  auto kernel = row_tiles ? attn_relpos_kernel<true> : attn_relpos_kernel<false>;
  static DeviceFlags configured[2];
  if (!configured[row_tiles].current()) {   // opt in to the device maximum once per device: smem varies with the extent
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_relpos_kernel)");
    configured[row_tiles].current() = true;
  }
  dim3 grid((N + kRpBQ - 1) / kRpBQ, heads, BW);
  kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv_bf16), ld_qkv, static_cast<const __nv_bfloat16*>(rcat_hi_bf16),
      static_cast<const __nv_bfloat16*>(rcat_lo_bf16), static_cast<__nv_bfloat16*>(out_bf16), ld_out, N, heads, Sh, Sw, RP, magic,
      scale * 1.4426950408889634f);
  count_launch();
  VDR_CHECK_LAUNCH("attn_relpos_kernel");
  return VDR_OK;
}

extern "C" int vdr_attn_relpos_windows_fwd(const void* qkv_bf16, int64_t ld_qkv, const float* qkv_bias, const void* rcat_hi_bf16,
                                           const void* rcat_lo_bf16, void* out_bf16, int64_t ld_out, int B, int gh, int gw, int ws, int heads,
                                           float scale, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(qkv_bf16 && qkv_bias && rcat_hi_bf16 && rcat_lo_bf16 && out_bf16, VDR_EINVAL, "vdr_attn_relpos_windows_fwd: null pointer");
  VDR_CHECK_ARG(B > 0 && gh > 0 && gw > 0 && ws > 0 && heads > 0 && heads <= 65535, VDR_EINVAL,
                "vdr_attn_relpos_windows_fwd: bad shape B=%d grid %dx%d ws=%d heads=%d", B, gh, gw, ws, heads);
  const int N = ws * ws;
  VDR_CHECK_ARG(N <= kWinRows && 4 * ws - 2 <= 64, VDR_EINVAL,
                "vdr_attn_relpos_windows_fwd: windows of %dx%d tokens do not fit the resident-key kernel (<= %d tokens); use vdr_window_rows + vdr_attn_relpos_fwd",
                ws, ws, kWinRows);
  VDR_CHECK_ARG(ld_qkv >= 3LL * heads * 64 && ld_qkv % 8 == 0 && ld_out >= heads * 64LL && ld_out % 8 == 0 && aligned16(qkv_bf16) &&
                    aligned16(out_bf16) && aligned16(rcat_hi_bf16) && aligned16(rcat_lo_bf16) && aligned16(qkv_bias),
                VDR_EALIGN, "vdr_attn_relpos_windows_fwd: qkv (rows, >= 3*heads*64) / out (rows, >= heads*64), ld %% 8 == 0, 16-byte aligned");
  const int nwh = (gh + ws - 1) / ws, nww = (gw + ws - 1) / ws;
  const int64_t BW = (int64_t)B * nwh * nww;
  VDR_CHECK_ARG(BW <= 65535, VDR_EINVAL, "vdr_attn_relpos_windows_fwd: %lld windows exceed the grid limit (65535): split the batch", (long long)BW);
  // SAM's own window size: the tcgen05 kernel (sam_window_tc.cu); VDR_SAM_WIN_MMA_SYNC=1 keeps the mma.sync kernel below for A/B runs
  static const bool legacy = getenv("VDR_SAM_WIN_MMA_SYNC") != nullptr;
  if (ws == 14 && !legacy)
    return launch_attn_win14_tc(qkv_bf16, ld_qkv, qkv_bias, rcat_hi_bf16, rcat_lo_bf16, out_bf16, ld_out, B, gh, gw, heads, scale,
                                reinterpret_cast<cudaStream_t>(stream));
  const int rp0 = ((ws + 1) & ~1) + ws;
  const int RP = rp0 + (8 - rp0 % 32 + 32) % 32;
  const size_t smem_w = (2 * (size_t)kWinRows * kRpPitch + 2 * (size_t)kRpTile) * sizeof(__nv_bfloat16) + (size_t)kWinQ * RP * sizeof(float);
  static DeviceFlags configured;
  if (!configured.current()) {
    cudaError_t e = cudaFuncSetAttribute(attn_relpos_win_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_relpos_win_kernel<mapped>)");
    configured.current() = true;
  }
  const uint32_t magic = ws > 1 ? (uint32_t)((0x100000000ULL + (uint64_t)ws - 1) / (uint64_t)ws) : 0u;
  WinMap wm{gh, gw, nwh, nww, qkv_bias};
  attn_relpos_win_kernel<true><<<dim3(heads, (unsigned)BW, N > kWinQ ? 2 : 1), kWinThreads, smem_w, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv_bf16), ld_qkv, static_cast<const __nv_bfloat16*>(rcat_hi_bf16),
      static_cast<const __nv_bfloat16*>(rcat_lo_bf16), static_cast<__nv_bfloat16*>(out_bf16), ld_out, N, heads, ws, ws, RP, magic,
      scale * 1.4426950408889634f, wm);
  count_launch();
  VDR_CHECK_LAUNCH("attn_relpos_win_kernel<mapped>");
  return VDR_OK;
}

extern "C" int vdr_im2col3x3_tokens(const void* X_bf16, int64_t ldx, void* A_bf16, int64_t lda, int B, int H, int W, int C,
                                    vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(X_bf16 && A_bf16, VDR_EINVAL, "vdr_im2col3x3_tokens: null pointer");
  VDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, VDR_EINVAL, "vdr_im2col3x3_tokens: bad shape B=%d H=%d W=%d C=%d (C %% 8 == 0)", B, H, W, C);
  VDR_CHECK_ARG(aligned16(X_bf16) && aligned16(A_bf16) && ldx % 8 == 0 && lda % 8 == 0 && ldx >= C && lda >= 9LL * C, VDR_EALIGN,
                "vdr_im2col3x3_tokens: pointers must be 16-byte aligned, ldx >= C, lda >= 9*C, both multiples of 8");
  const int64_t total_vec = (int64_t)B * H * W * 9 * (C >> 3);
  int64_t blocks = (total_vec + 255) / 256;
  const int64_t max_blocks = (int64_t)num_sms() * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  im2col3x3_tokens_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(X_bf16), ldx, static_cast<__nv_bfloat16*>(A_bf16), lda, H, W, C, total_vec);
  count_launch();
  VDR_CHECK_LAUNCH("im2col3x3_tokens_kernel");
  return VDR_OK;
}
