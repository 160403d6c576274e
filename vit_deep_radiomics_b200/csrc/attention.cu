// Fused flash-style self-attention forward on tcgen05 (head_dim 64), sm_100a.
//
// One CTA = 128 query rows of one (batch, head); 128 threads, thread t owns query row t = TMEM lane t
// so row max / row sum need no shuffles.  Per 128-key block:
//   S = Q K^T      tcgen05.mma 128x128x64 (both operands K-major, TMA SWIZZLE_128B tiles) -> TMEM
//   softmax        tcgen05.ld S, online max/sum in registers, P (bf16) -> swizzled smem
//   O_j = P V      tcgen05.mma 128x64x128 (A = P K-major, B = V MN-major straight from the TMA tile)
//   O = O*alpha + O_j in registers (no TMEM round trip for the rescale)
// K/V tiles stream through a 3-slot TMA ring; two CTAs are resident per SM so the tensor pipe of one
// overlaps the softmax of the other.  Q, K, V are read in place from the packed QKV GEMM output.
#include "common.cuh"

namespace vdr {

constexpr int kAttnThreads = 128;
constexpr int kBQ = 128, kBKV = 128, kHD = 64;
constexpr int kTileBytes = 128 * kHD * 2;               // 16 KB: one 128 x 64 bf16 tile
constexpr int kAttnSmem = 6 * kTileBytes /*Q, 3 ring slots, P lo/hi*/ + 1024 /*align*/ + 128 /*barriers*/;
constexpr int kAttnTmemCols = 256;                       // S: [0,128)  O: [128,192)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major B operand (V tile: 128 kv rows x 64 d, 128-byte rows, SWIZZLE_128B), 16 kv rows per MMA:
// canonical layout ((8,8,1),(8,2)):((1,8,LBO),(64,SBO)) elements -> SBO = 1024 B between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (single 64-element MN group: unused)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct AttnParams {
  __nv_bfloat16* out;
  float* lse;
  int64_t ld_out;
  int B, N, heads, d;
  float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads, 2)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  // layout: Q | ring0 | ring1 | ring2 | P_lo | P_hi | barriers
  const uint32_t sQ = base, sRing = base + kTileBytes, sP = base + 4 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);
  uint64_t* bar_q = bars;          // Q landed
  uint64_t* bar_kv = bars + 1;     // [3] ring slot landed
  uint64_t* bar_s = bars + 4;      // S = QK^T complete
  uint64_t* bar_o = bars + 5;      // O_j = PV complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * kBQ, head = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.N;                       // first token row of this image in the qkv matrix
  const int colQ = head * kHD, colK = p.d + head * kHD, colV = 2 * p.d + head * kHD;
  const int nkv = (p.N + kBKV - 1) / kBKV;
  const int ntiles = 2 * nkv;                         // ring tiles: K0 V0 K1 V1 ...

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAttnTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;

  auto issue_tile = [&](int t) {   // ring tile t: even = K block t/2, odd = V block t/2
    const int slot = t % 3;
    const int kv0 = (t >> 1) * kBKV;
    mbar_arrive_expect_tx(&bar_kv[slot], kTileBytes);
    tma_load_2d(&tmQKV, &bar_kv[slot], smem + kTileBytes * (1 + slot), (t & 1) ? colV : colK, row_base + kv0);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, kTileBytes);
    tma_load_2d(&tmQKV, bar_q, smem, colQ, row_base + q0);
    for (int t = 0; t < 3 && t < ntiles; ++t) issue_tile(t);
  }

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // B (= V) is MN-major

  float o_acc[kHD];
#pragma unroll
  for (int i = 0; i < kHD; ++i) o_acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < nkv; ++j) {
    const int kv0 = j * kBKV;
    // ---- S = Q K_j^T
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      const int t = 2 * j;
      mbar_wait(&bar_kv[t % 3], (t / 3) & 1);
      tc_fence_after();
      const uint64_t dq = umma_desc_kmajor_sw128(sQ);
      const uint64_t dk = umma_desc_kmajor_sw128(sRing + (t % 3) * kTileBytes);
#pragma unroll
      for (int k = 0; k < kHD / 16; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    if (tid == 0 && 2 * j + 3 < ntiles) issue_tile(2 * j + 3);   // K_j's slot is free again
    __syncwarp();

    // ---- online softmax over this thread's row (two passes over TMEM: max, then exp)
    const bool tail = kv0 + kBKV > p.N;
    float m_blk = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_S + lane_sel + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float s = __uint_as_float(r[i]);
        if (tail && kv0 + c * 32 + i >= p.N) s = -INFINITY;
        m_blk = fmaxf(m_blk, s);
      }
    }
    const float m_new = fmaxf(m_run, m_blk * p.scale_log2);
    const float alpha = ex2(m_run - m_new);
    float l_blk = 0.f;
    const uint32_t prow = sP + tid * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_S + lane_sel + c * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
        float p0 = ex2(fmaf(s0, p.scale_log2, -m_new));
        float p1 = ex2(fmaf(s1, p.scale_log2, -m_new));
        if (tail) {
          if (kv0 + c * 32 + i >= p.N) p0 = 0.f;
          if (kv0 + c * 32 + i + 1 >= p.N) p1 = 0.f;
        }
        // the row sum uses the bf16-rounded probabilities that the P V product actually sees
        const uint32_t w = pack_bf16x2(p0, p1);
        const float2 pr = unpack_bf16x2(w);
        l_blk += pr.x + pr.y;
        pk[i >> 1] = w;
      }
      // P[row][kv] bf16, K-major, 128B swizzle: halves of 64 kv columns (16 KB each)
      const uint32_t half_base = prow + (c >> 1) * kTileBytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q) ^ static_cast<uint32_t>(tid & 7);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(half_base + chunk * 16), "r"(pk[q * 4]),
                     "r"(pk[q * 4 + 1]), "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                     : "memory");
      }
    }
    l_run = l_run * alpha + l_blk;
    m_run = m_new;
    fence_proxy_async_smem();   // P stores -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();

    // ---- O_j = P V_j
    if (tid == 0) {
      const int t = 2 * j + 1;
      mbar_wait(&bar_kv[t % 3], (t / 3) & 1);
      tc_fence_after();
      const uint32_t sV = sRing + (t % 3) * kTileBytes;
#pragma unroll
      for (int k = 0; k < kBKV / 16; ++k) {
        const uint64_t dp = umma_desc_kmajor_sw128(sP + (k >> 2) * kTileBytes) + 2 * (k & 3);
        const uint64_t dv = umma_desc_mnmajor_sw128(sV + k * 2048);   // 16 kv rows x 128 B
        umma_ss(tmem_O, dp, dv, idesc_o, k != 0);
      }
      umma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
    if (tid == 0 && 2 * j + 4 < ntiles) issue_tile(2 * j + 4);   // V_j's slot is free again
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_O + lane_sel + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(r[i]));
    }
    tc_fence_before();   // ordered before the next iteration's barrier -> S / O may be overwritten
  }

  // ---- normalise and store
  const int q = q0 + tid;
  if (q < p.N) {
    const float inv = 1.f / l_run;
    __nv_bfloat16* op = p.out + static_cast<int64_t>(row_base + q) * p.ld_out + head * kHD;
#pragma unroll
    for (int i = 0; i < kHD; i += 8) {
      uint4 o;
      o.x = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv);
      o.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
      o.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
      o.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
      *reinterpret_cast<uint4*>(op + i) = o;
    }
    if (p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (m_run + log2f(l_run)) * 0.69314718055994531f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kAttnTmemCols>(tmem_base);
  }
}

}  // namespace vdr

extern "C" int vdr_flash_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int B,
                                  int N, int heads, float scale, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(qkv && out, VDR_EINVAL, "vdr_flash_attn_fwd: null pointer");
  VDR_CHECK_ARG(B > 0 && N > 0 && heads > 0, VDR_EINVAL, "vdr_flash_attn_fwd: bad shape B=%d N=%d heads=%d", B, N, heads);
  const int d = heads * kHD;
  VDR_CHECK_ARG(ld_qkv >= 3 * d && ld_qkv % 8 == 0 && ld_out >= d && ld_out % 8 == 0, VDR_EALIGN, "vdr_flash_attn_fwd: ld_qkv (%lld) / ld_out (%lld) too small or not multiples of 8", (long long)ld_qkv, (long long)ld_out);
  VDR_CHECK_ARG(aligned16(qkv) && aligned16(out), VDR_EALIGN, "vdr_flash_attn_fwd: pointers must be 16-byte aligned");
  VDR_CHECK_ARG(B <= 65535 && heads <= 65535, VDR_EINVAL, "vdr_flash_attn_fwd: B and heads must be <= 65535");
  CUtensorMap tm;
  int rc = make_tmap_2d_bf16(&tm, qkv, (uint64_t)B * N, (uint64_t)3 * d, (uint64_t)ld_qkv, 128, kHD);
  if (rc != VDR_OK) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(flash_attn_fwd)");
    configured = true;
  }
  AttnParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.ld_out = ld_out;
  p.B = B; p.N = N; p.heads = heads; p.d = d;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((N + kBQ - 1) / kBQ, heads, B);
  flash_attn_fwd_kernel<<<grid, kAttnThreads, kAttnSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
  count_launch();
  VDR_CHECK_LAUNCH("flash_attn_fwd_kernel");
  return VDR_OK;
}
