// Fused flash-style self-attention BACKWARD on tcgen05 (head_dim 64), sm_100a.
// Replaces the autograd of the SDPA inside nn.TransformerEncoderLayer on the classifier's training path
// (reference src/models_archs.py:146 driven by loss.backward(), src/train_models.py:683).
//
// One CTA = one 128-key block j of one (image, head); it walks the query blocks i and keeps dK_j and dV_j in TMEM:
//   S  = Q_i K_j^T,  dP = dO_i V_j^T                      tcgen05.mma 128x128x64 (K-major operands from TMA tiles)
//   P  = exp2(S*scale*log2e - lse*log2e),  dS = P (dP - delta) scale        softmax warps: TMEM -> registers -> bf16 tiles in smem
//   dV_j += P^T dO_i,  dK_j += dS^T Q_i                    tcgen05.mma 128x64x128 (A = P / dS MN-major, B = dO / Q MN-major)
//   dQ_i  = dS K_j                                         tcgen05.mma 128x64x128 (A = dS K-major, B = K_j MN-major) -> red.add to f32 dQ
// Scores are never written to HBM (the previous backward materialised S, dP (f32) and P, dS (bf16) per head).
// Warps 0-3: softmax / epilogue (thread t = TMEM lane t = query row, later key row); warp 4 lane 0: TMA + MMA issue.
#include "common.cuh"

namespace vdr {

constexpr int kBwdThreads = 160;
constexpr int kBT = 128, kBD = 64;
constexpr int kTile = kBT * kBD * 2;                    // 16 KB: one 128 x 64 bf16 tile
// K_j | V_j | Q_i x2 | dO_i x2 | P (two 128x64 sub-tiles) | dS (two sub-tiles) | barriers
constexpr int kBwdSmem = 10 * kTile + 256;
constexpr int kBwdTmemCols = 512;                        // S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ [384,448)

struct AttnBwdParams {
  const float* lse;      // (B, heads, N) natural-log sum-exp from the forward
  const float* delta;    // (heads, B*N)  rowsum(dO * O)
  float* dq_acc;         // (B*N, d) f32, zero before the launch
  __nv_bfloat16* dqkv;   // (B*N, 3d): the dK and dV parts are written here
  int64_t ld_dqkv;
  int B, N, heads, d;
  float scale, scale_log2;
  DropSpec drop;         // attention dropout of the forward (thr16 == 0: none); the mask is regenerated here
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MN-major operand, 128-byte rows of 64 elements (SWIZZLE_128B), 8-row groups 1024 B apart along K, 64-element groups `lbo`
// bytes apart along M/N (two [128][64] sub-tiles for a 128-wide M)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
flash_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + kTile, sQ = base + 2 * kTile, sDO = base + 4 * kTile, sP = base + 6 * kTile, sDS = base + 8 * kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 10 * kTile);
  uint64_t* bar_kv = bars;           // K_j, V_j landed
  uint64_t* bar_ld = bars + 1;       // [2] Q_i, dO_i landed
  uint64_t* bar_sdp = bars + 3;      // S and dP complete                  (tcgen05.commit)
  uint64_t* bar_pds = bars + 4;      // P and dS are in shared memory       (4 softmax warps)
  uint64_t* bar_grad = bars + 5;     // dV, dK, dQ MMAs complete            (tcgen05.commit)
  uint64_t* bar_free = bars + 6;     // dQ read, S / dP / P / dS reusable   (4 softmax warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.N;
  const int colQ = head * kBD, colK = p.d + head * kBD, colV = 2 * p.d + head * kBD;
  const int nq = (p.N + kBT - 1) / kBT;
  const int kv0 = j * kBT;

  if (tid == 128) {
    if (base & 1023u) { printf("vdr: attention-backward smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_kv, 1);
    mbar_init(&bar_ld[0], 1);
    mbar_init(&bar_ld[1], 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 4);
    mbar_init(bar_grad, 1);
    mbar_init(bar_free, 4);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_kv, 2 * kTile);
    tma_load_2d(&tmQKV, bar_kv, smem, colK, row_base + kv0);
    tma_load_2d(&tmQKV, bar_kv, smem + kTile, colV, row_base + kv0);
    mbar_arrive_expect_tx(&bar_ld[0], 2 * kTile);
    tma_load_2d(&tmQKV, &bar_ld[0], smem + 2 * kTile, colQ, row_base);
    tma_load_2d(&tmDO, &bar_ld[0], smem + 4 * kTile, head * kBD, row_base);
  }
  if (warp == 0) tmem_alloc<kBwdTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;

  if (warp == 4) {
    // the whole warp runs the loop and waits; TMA / tcgen05 instructions under elect.sync (one lane, operands straight to uniform
    // registers: under `lane == 0` each of the 32 MMAs of a block was wrapped in an ELECT / BRA.U.ANY waterfall)
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);     // S, dP: A, B K-major
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, 64, 1, 1);      // dV, dK: A MN-major (P^T / dS^T), B MN-major (dO / Q)
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);      // dQ: A = dS K-major, B = K_j MN-major
      mbar_wait(bar_kv, 0);
      for (int i = 0; i < nq; ++i) {
        const int buf = i & 1;
        mbar_wait(&bar_ld[buf], (i >> 1) & 1);
        if (i > 0) mbar_wait(bar_free, (i - 1) & 1);       // block i - 1 is done: S / dP columns, P / dS tiles, its Q / dO buffer are free
        tc_fence_after();
        if (i + 1 < nq && elect_one()) {   // prefetch the next query block into the buffer block i - 1 used
          mbar_arrive_expect_tx(&bar_ld[buf ^ 1], 2 * kTile);
          tma_load_2d(&tmQKV, &bar_ld[buf ^ 1], smem + (2 + (buf ^ 1)) * kTile, colQ, row_base + (i + 1) * kBT);
          tma_load_2d(&tmDO, &bar_ld[buf ^ 1], smem + (4 + (buf ^ 1)) * kTile, head * kBD, row_base + (i + 1) * kBT);
        }
        const uint32_t q_t = sQ + buf * kTile, do_t = sDO + buf * kTile;
        const uint64_t dq = umma_desc_kmajor_sw128(q_t), dk = umma_desc_kmajor_sw128(sK);
        const uint64_t ddo = umma_desc_kmajor_sw128(do_t), dv = umma_desc_kmajor_sw128(sV);
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBD / 16; ++k) umma_ss(tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < kBD / 16; ++k) umma_ss(tDP, ddo + 2 * k, dv + 2 * k, idesc_s, k != 0);
          umma_commit(bar_sdp);
        }
        __syncwarp();
        mbar_wait(bar_pds, i & 1);
        tc_fence_after();
        // K dimension = the 128 query rows of the block: 8 steps of 16 rows (2048 B = 128 address units per step in every tile)
        const uint64_t a_p = umma_desc_mn_sw128(sP, kTile), a_ds = umma_desc_mn_sw128(sDS, kTile);
        const uint64_t b_do = umma_desc_mn_sw128(do_t, kTile), b_q = umma_desc_mn_sw128(q_t, kTile);
        // dQ_i = dS K_j: K dimension = the 128 keys: dS K-major as two [128][64] sub-tiles, K_j MN-major (16 key rows per step)
        const uint64_t a_dsk0 = umma_desc_kmajor_sw128(sDS), a_dsk1 = umma_desc_kmajor_sw128(sDS + kTile);
        const uint64_t b_k = umma_desc_mn_sw128(sK, kTile);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBT / 16; ++k) umma_ss(tDV, a_p + 128 * k, b_do + 128 * k, idesc_g, (i | k) != 0);
#pragma unroll
          for (int k = 0; k < kBT / 16; ++k) umma_ss(tDK, a_ds + 128 * k, b_q + 128 * k, idesc_g, (i | k) != 0);
#pragma unroll
          for (int k = 0; k < kBT / 16; ++k)
            umma_ss(tDQ, (k < 4 ? a_dsk0 + 2 * k : a_dsk1 + 2 * (k - 4)), b_k + 128 * k, idesc_q, k != 0);
          umma_commit(bar_grad);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================================================== softmax / epilogue warps: thread = TMEM lane = row
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const float log2e = 1.4426950408889634f;
    for (int i = 0; i < nq; ++i) {
      const int q = i * kBT + tid;                          // query row of this thread in block i
      const bool q_ok = q < p.N;
      const float nlse = q_ok ? -p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] * log2e : 0.f;
      const float dlt = q_ok ? p.delta[static_cast<int64_t>(head) * p.B * p.N + row_base + q] : 0.f;
      mbar_wait(bar_sdp, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                         // 32 key columns at a time
        uint32_t s[32], dp[32];
        tmem_ld_32x32b_x32(tS + lane_sel + c * 32, s);
        tmem_ld_32x32b_x32(tDP + lane_sel + c * 32, dp);
        tmem_ld_wait();
        uint32_t pw[16], dw[16];
        // forward: O = (P (.) M) V with M = mask / (1 - p)  =>  dV = (P (.) M)^T dO,  dP = M (.) (dO V^T),  dS = P (.) (dP - delta) * scale
        // (delta = rowsum(dO (.) O) already contains the mask)
        const bool dropping = p.drop.thr16 != 0;
        const uint64_t drow = (static_cast<uint64_t>(b) * p.heads + head) * p.N + q;
        const float dsc = dropping ? drop_scale(p.drop) : 1.f;
        uint4 dbits = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const bool ok0 = q_ok && (kv0 + c * 32 + e) < p.N, ok1 = q_ok && (kv0 + c * 32 + e + 1) < p.N;
          const float p0 = ok0 ? ex2f(fmaf(__uint_as_float(s[e]), p.scale_log2, nlse)) : 0.f;
          const float p1 = ok1 ? ex2f(fmaf(__uint_as_float(s[e + 1]), p.scale_log2, nlse)) : 0.f;
          float k0 = 1.f, k1 = 1.f;
          if (dropping) {
            if ((e & 7) == 0) dbits = drop_bits8(p.drop, drow, static_cast<uint32_t>((kv0 + c * 32 + e) >> 3));
            k0 = drop_lane16(dbits, e & 7) >= p.drop.thr16 ? dsc : 0.f;
            k1 = drop_lane16(dbits, (e & 7) + 1) >= p.drop.thr16 ? dsc : 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dp[e]) * k0 - dlt) * p.scale, d1 = p1 * (__uint_as_float(dp[e + 1]) * k1 - dlt) * p.scale;
          pw[e >> 1] = pack_bf16x2(p0 * k0, p1 * k1);
          dw[e >> 1] = pack_bf16x2(d0, d1);
        }
        // row tid of sub-tile (c >> 1), 16-byte chunks (c & 1) * 4 .. + 3, SWIZZLE_128B
        const uint32_t rowoff = (c >> 1) * kTile + tid * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = rowoff + ((((c & 1) * 4 + g) ^ (tid & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off), "r"(pw[g * 4]), "r"(pw[g * 4 + 1]), "r"(pw[g * 4 + 2]), "r"(pw[g * 4 + 3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(dw[g * 4]), "r"(dw[g * 4 + 1]), "r"(dw[g * 4 + 2]), "r"(dw[g * 4 + 3]) : "memory");
        }
      }
      fence_proxy_async_smem();                             // generic-proxy stores -> visible to the tensor core's smem reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pds);
      // ---- dQ_i (this block's contribution) -> f32 accumulator in global memory
      mbar_wait(bar_grad, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tDQ + lane_sel + c * 32, r);
        tmem_ld_wait();
        if (q_ok) {
          float* dst = p.dq_acc + static_cast<int64_t>(row_base + q) * p.d + head * kBD + c * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + e), "f"(__uint_as_float(r[e])), "f"(__uint_as_float(r[e + 1])),
                         "f"(__uint_as_float(r[e + 2])), "f"(__uint_as_float(r[e + 3])) : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
    }
    // ---- epilogue: dV_j, dK_j (lane = key row) -> bf16
    const int kv = kv0 + tid;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32((which == 0 ? tDV : tDK) + lane_sel + c * 32, r);
        tmem_ld_wait();
        if (kv < p.N) {
          __nv_bfloat16* dst = p.dqkv + static_cast<int64_t>(row_base + kv) * p.ld_dqkv + (which == 0 ? 2 * p.d : p.d) + head * kBD + c * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[e]), __uint_as_float(r[e + 1]));
            o.y = pack_bf16x2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
            o.z = pack_bf16x2(__uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]));
            o.w = pack_bf16x2(__uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]));
            *reinterpret_cast<uint4*>(dst + e) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kBwdTmemCols>(tmem_base);
  }
}

// dQ accumulator (f32) -> the q part of dqkv (bf16)
__global__ void __launch_bounds__(256) dq_to_bf16_kernel(const float* __restrict__ acc, int64_t rows, int d, __nv_bfloat16* __restrict__ dqkv, int64_t ld) {
  const int64_t total = rows * (d >> 3);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (d >> 3);
    const int c = static_cast<int>(i % (d >> 3)) << 3;
    const float4 a = *reinterpret_cast<const float4*>(acc + r * d + c), bq = *reinterpret_cast<const float4*>(acc + r * d + c + 4);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(bq.x, bq.y); o.w = pack_bf16x2(bq.z, bq.w);
    *reinterpret_cast<uint4*>(dqkv + r * ld + c) = o;
  }
}

}  // namespace vdr

static size_t bwd_delta_bytes(int B, int N, int heads) { return (((size_t)B * N * heads * sizeof(float)) + 255) & ~(size_t)255; }

extern "C" size_t vdr_flash_attn_bwd_workspace_bytes(int B, int N, int heads) {
  return bwd_delta_bytes(B, N, heads) + (size_t)B * N * heads * 64 * sizeof(float);
}

extern "C" int vdr_flash_attn_bwd(const void* qkv, int64_t ld_qkv, const void* O, const void* dO, int64_t ld_o, const float* lse,
                                  void* dqkv, int64_t ld_dqkv, int B, int N, int heads, float scale, const vdr_dropout* drop, void* workspace,
                                  size_t workspace_bytes, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(qkv && O && dO && lse && dqkv && workspace, VDR_EINVAL, "vdr_flash_attn_bwd: null pointer");
  VDR_CHECK_ARG(B > 0 && N > 0 && heads > 0 && B <= 65535 && heads <= 65535, VDR_EINVAL, "vdr_flash_attn_bwd: bad shape B=%d N=%d heads=%d", B, N, heads);
  const int d = heads * kBD;
  VDR_CHECK_ARG(ld_qkv >= 3 * d && ld_qkv % 8 == 0 && ld_o >= d && ld_o % 8 == 0 && ld_dqkv >= 3 * d && ld_dqkv % 8 == 0, VDR_EALIGN,
                "vdr_flash_attn_bwd: leading dimensions too small or not multiples of 8");
  VDR_CHECK_ARG(aligned16(qkv) && aligned16(O) && aligned16(dO) && aligned16(dqkv) && aligned16(workspace), VDR_EALIGN, "vdr_flash_attn_bwd: pointers must be 16-byte aligned");
  VDR_CHECK_ARG(workspace_bytes >= vdr_flash_attn_bwd_workspace_bytes(B, N, heads), VDR_EWORKSPACE, "vdr_flash_attn_bwd: workspace too small (%zu < %zu)",
                workspace_bytes, vdr_flash_attn_bwd_workspace_bytes(B, N, heads));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* delta = static_cast<float*>(workspace);
  float* dq_acc = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + bwd_delta_bytes(B, N, heads));
  int rc = vdr_attn_delta(dO, O, ld_o, B * N, heads, delta, stream);     // delta[h][b*N + i]
  if (rc != VDR_OK) return rc;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)B * N * d * sizeof(float), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(dQ accumulator)");
  CUtensorMap tmQKV, tmDO;
  rc = make_tmap_2d_bf16(&tmQKV, qkv, (uint64_t)B * N, (uint64_t)3 * d, (uint64_t)ld_qkv, 128, kBD);
  if (rc != VDR_OK) return rc;
  rc = make_tmap_2d_bf16(&tmDO, dO, (uint64_t)B * N, (uint64_t)d, (uint64_t)ld_o, 128, kBD);
  if (rc != VDR_OK) return rc;
  static DeviceFlags configured;
  if (!configured.current()) {
    e = cudaFuncSetAttribute(flash_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(flash_attn_bwd)");
    configured.current() = true;
  }
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.dq_acc = dq_acc;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv); p.ld_dqkv = ld_dqkv;
  p.B = B; p.N = N; p.heads = heads; p.d = d;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.drop = DropSpec{0ull, 0u, 0u, nullptr};
  if (drop != nullptr && drop->thr16 != 0) {
    VDR_CHECK_ARG(drop->thr16 < 65536u, VDR_EINVAL, "vdr_flash_attn_bwd: dropout threshold must be < 65536");
    p.drop = DropSpec{drop->seed, drop->site, drop->thr16, reinterpret_cast<const unsigned long long*>(drop->seed_offset)};
  }
  dim3 grid((N + kBT - 1) / kBT, heads, B);
  flash_attn_bwd_kernel<<<grid, kBwdThreads, kBwdSmem, s>>>(tmQKV, tmDO, p);
  count_launch();
  VDR_CHECK_LAUNCH("flash_attn_bwd_kernel");
  const int64_t rows = (int64_t)B * N;
  int64_t blocks = (rows * (d >> 3) + 255) / 256;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  dq_to_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(dq_acc, rows, d, static_cast<__nv_bfloat16*>(dqkv), ld_dqkv);
  count_launch();
  VDR_CHECK_LAUNCH("dq_to_bf16_kernel");
  return VDR_OK;
}
