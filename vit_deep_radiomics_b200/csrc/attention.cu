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
#include <type_traits>

#include "common.cuh"

namespace vdr {

constexpr int kAttnThreads = 256;   // softmax warpgroup + issuer warpgroup
constexpr int kBQ = 128, kBKV = 128, kHD = 64;
constexpr int kTileBytes = 128 * kHD * 2;               // 16 KB: one 128 x 64 bf16 tile
constexpr int kAttnSmem = 7 * kTileBytes /*Q, 4 ring slots, P lo/hi*/ + 256 /*barriers*/;
constexpr int kAttnTmemCols = 256;                       // S: [0,128)  O: [128,192)  P (bf16 pairs): [192,256)

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
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  __nv_bfloat16* out;
  float* lse;
  int64_t ld_out;
  int B, N, heads, d;
  float scale_log2;
  unsigned long long* trace;   // debug: per-iteration timestamps of CTA (0,0,0) thread 32
};

__device__ __forceinline__ unsigned long long attn_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ATT_TRACE(ev)                                                                                              \
  do {                                                                                                             \
    if (p.trace != nullptr && tid == 32 && blockIdx.x + blockIdx.y + blockIdx.z == 0 && j < 16) p.trace[j * 16 + (ev)] = attn_gtime(); \
  } while (0)

#define ISS_TRACE(ev)                                                                                              \
  do {                                                                                                             \
    if (p.trace != nullptr && blockIdx.x + blockIdx.y + blockIdx.z == 0 && j < 16) p.trace[j * 16 + 8 + (ev)] = attn_gtime(); \
  } while (0)

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// exp2 of two non-positive arguments on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial,
// max relative error 7.7e-5 -- far below the bf16 rounding of P): relieves the MUFU, which is the ceiling of
// head_dim-64 attention (128 x 128 exponentials per 4.2 MFLOP block).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t magic2 = pack2(12582912.f, 12582912.f);           // 1.5 * 2^23: t = x + magic rounds x to an integer
  const uint64_t t2 = add2(x2, magic2);
  const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), x2);              // f = x - round(x) in [-0.5, 0.5]
  uint64_t q2 = fma2(f2, pack2(0.05508868396282196f, 0.05508868396282196f), pack2(0.24260404706001282f, 0.24260404706001282f));
  q2 = fma2(q2, f2, pack2(0.6932762265205383f, 0.6932762265205383f));
  q2 = fma2(q2, f2, pack2(0.9999289512634277f, 0.9999289512634277f));
  float t0, t1, q0, q1;
  unpack2(t2, t0, t1);
  unpack2(q2, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));   // * 2^round(x) through the exponent field
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// Warp roles: warps 0-3 = softmax warpgroup (thread t owns query row t = TMEM lane t);
//             warp 4    = issuer (TMA ring + tcgen05.mma), warps 5-7 idle (setmaxnreg works per warpgroup).
// Pipeline per 128-key block j (no CTA-wide barrier in the loop):
//   issuer : wait S_j consumed -> prefetch K_{j+2}, issue S_{j+1};  wait P_j ready -> issue O_j = P_j V_j
//   softmax: wait S_j -> registers -> signal "consumed" -> max / exp2 / sum -> fold O_{j-1} (finished long ago)
//            -> write P_j -> signal "ready"
// so the tensor pipe computes S_{j+1} and O_{j-1} while the exponentials of block j are evaluated.
__global__ void __launch_bounds__(kAttnThreads, 2)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  // layout: Q | ring0..3 | P_lo | P_hi | barriers
  const uint32_t sQ = base, sRing = base + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
  uint64_t* bar_q = bars;            // Q landed
  uint64_t* bar_kv = bars + 1;       // [4] ring slot landed
  uint64_t* bar_s = bars + 5;        // S_j = Q K_j^T complete            (tcgen05.commit)
  uint64_t* bar_o = bars + 6;        // O_j = P_j V_j complete            (tcgen05.commit)
  uint64_t* bar_sfree = bars + 7;    // S_j is in registers               (4 softmax warps arrive)
  uint64_t* bar_pready = bars + 8;   // P_j is in smem, O_{j-1} consumed  (4 softmax warps arrive)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBQ, head = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.N;                       // first token row of this image in the qkv matrix
  const int colQ = head * kHD, colK = p.d + head * kHD, colV = 2 * p.d + head * kHD;
  // Key blocks: full 128-key blocks on the tensor cores; a ragged last block either runs as a narrow MMA block
  // (16-column granularity) or, when it holds only a few keys (<= 8, e.g. the 1025th token of a 32x32-patch image
  // + CLS), is folded into the epilogue on the CUDA cores instead of costing every CTA one more pipeline round trip.
  const int nkv_all = (p.N + kBKV - 1) / kBKV;
  const int last_keys = p.N - (nkv_all - 1) * kBKV;              // keys in the last block (1..128)
  const int tail_keys = (last_keys <= 8 && nkv_all > 1) ? last_keys : 0;
  const int nkv = tail_keys ? nkv_all - 1 : nkv_all;             // blocks that go through the MMA pipeline
  const int valid_last = tail_keys ? kBKV : last_keys;           // keys in the last MMA block
  const int ntail = (valid_last + 15) & ~15;                     // ... rounded to the MMA's 16-column granularity

  if (tid == 0) {
    if (base & 1023u) { printf("vdr: attention smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_sfree, 4);
    mbar_init(bar_pready, 4);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<kAttnTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128, tmem_P = tmem_base + 192;

  if (warp >= 4) {
    // =============================================================== issuer warpgroup
    reg_dec<48>();
    // Two issuing threads so that the two dependent MMA chains of a block (S_{j+1} = Q K^T: 4 steps, O += P V:
    // 8 steps, each step waiting ~130 cycles on the accumulator of the previous one) are dispatched concurrently.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // B (= V) is MN-major
    auto issue_tile = [&](int t) {   // ring tile t: even = K block t/2 (slots 0/2), odd = V block t/2 (slots 1/3)
      const int slot = t & 3;
      mbar_arrive_expect_tx(&bar_kv[slot], kTileBytes);
      tma_load_2d(&tmQKV, &bar_kv[slot], smem + kTileBytes * (1 + slot), (t & 1) ? colV : colK, row_base + (t >> 1) * kBKV);
    };
    if (warp == 4 && lane == 0) {
      // ---- K tiles + S = Q K^T
      auto issue_s = [&](int j) {
        const int t = 2 * j;
        mbar_wait(&bar_kv[t & 3], (t >> 2) & 1);
        tc_fence_after();
        const uint64_t dq = umma_desc_kmajor_sw128(sQ);
        const uint64_t dk = umma_desc_kmajor_sw128(sRing + (t & 3) * kTileBytes);
        // the last key block only computes the (16-column granular) part of S that has keys behind it
        const uint32_t idesc = (j == nkv - 1) ? umma_idesc_bf16(128, ntail) : idesc_s;
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc, k != 0);
        umma_commit(bar_s);
      };
      mbar_arrive_expect_tx(bar_q, kTileBytes);
      tma_load_2d(&tmQKV, bar_q, smem, colQ, row_base + q0);
      issue_tile(0);
      if (nkv > 1) issue_tile(2);
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        ISS_TRACE(0);
        mbar_wait(bar_sfree, j & 1);                       // S_j is in registers -> K_j's slot and the S columns are free
        tc_fence_after();
        ISS_TRACE(1);
        if (j + 1 < nkv) issue_s(j + 1);
        if (j + 2 < nkv) issue_tile(2 * j + 4);            // K_{j+2} into K_j's slot
        ISS_TRACE(2);
      }
    } else if (warp == 5 && lane == 0) {
      // ---- V tiles + O += P V
      issue_tile(1);
      if (nkv > 1) issue_tile(3);
      for (int j = 0; j < nkv; ++j) {
        ISS_TRACE(3);
        mbar_wait(bar_pready, j & 1);                      // P_j is in TMEM (and O rescaled if the maximum moved)
        tc_fence_after();
        ISS_TRACE(4);
        if (j >= 1 && j + 1 < nkv) {                       // V_{j+1} goes into V_{j-1}'s slot: O_{j-1} must be complete.
          mbar_wait(bar_o, (j - 1) & 1);                   // (waited BEFORE O_j is committed: a parity wait must never
          issue_tile(2 * j + 3);                           //  be two phases behind its barrier)
        }
        const int t = 2 * j + 1;
        mbar_wait(&bar_kv[t & 3], (t >> 2) & 1);
        tc_fence_after();
        const uint32_t sV = sRing + (t & 3) * kTileBytes;
        const int ksteps = (j == nkv - 1) ? ntail / 16 : kBKV / 16;
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t dv = umma_desc_mnmajor_sw128(sV + k * 2048);   // 16 kv rows x 128 B
          // A = P from TMEM (16 bf16 = 8 columns per K step); O accumulates in TMEM across all key blocks
          umma_ts(tmem_O, tmem_P + k * 8, dv, idesc_o, (j > 0 || k != 0) ? 1u : 0u);
        }
        umma_commit(bar_o);
        ISS_TRACE(5);
      }
    }
  } else {
    // =============================================================== softmax warpgroup
    reg_inc<208>();
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    // Online softmax with LAZY rescaling: probabilities are taken relative to a reference maximum m_ref that is
    // only moved when the running maximum exceeds it by more than 2^8; O then accumulates in TMEM across key
    // blocks (the P V MMAs run with accumulate = 1) and is touched by this warpgroup only on those rare moves.
    float m_ref = -INFINITY, l_run = 0.f;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);

    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * kBKV;
      ATT_TRACE(0);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      ATT_TRACE(1);
      float alpha = 1.f;
      bool moved = false;
      uint64_t lsum2 = 0ull;
      uint32_t pk[64];
      if (j == nkv - 1 && ntail < kBKV) {
        // ---- last, partial key block: only ntail (multiple of 16) columns exist in S; runtime loop over 16-column groups
        float mx = -INFINITY;
        for (int c = 0; c < ntail; c += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_S + lane_sel + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c + i < valid_last) ? __uint_as_float(r[i]) : -INFINITY);
        }
        const float m_new = fmaxf(m_ref, mx * p.scale_log2);
        moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);
        if (moved) {
          alpha = ex2(m_ref - m_new);
          m_ref = m_new;
        }
        float lsum = 0.f;
        if (j > 0) {   // P_{j-1} must have been consumed before its columns are overwritten
          mbar_wait(bar_o, (j - 1) & 1);
          tc_fence_after();
        }
        for (int c = 0; c < ntail; c += 16) {
          uint32_t r[16], w[8];
          tmem_ld_32x32b_x16(tmem_S + lane_sel + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = (c + i < valid_last) ? ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_ref)) : 0.f;
            const float p1 = (c + i + 1 < valid_last) ? ex2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_ref)) : 0.f;
            lsum += p0 + p1;
            w[i >> 1] = cvt_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x8(tmem_P + lane_sel + (c >> 1), w);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfree);
        lsum2 = pack2(lsum, 0.f);
      } else {
      // the whole 128-wide score row of this thread -> registers, then hand the S columns back
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tmem_S + lane_sel + c * 32, sr[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree);
      ATT_TRACE(2);
      {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1]));
        const float m_new = fmaxf(m_ref, mx * p.scale_log2);
        moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);   // warp-uniform: TMEM accesses are warp-wide
        if (moved) {
          alpha = ex2(m_ref - m_new);
          m_ref = m_new;
        }
        const uint64_t negm2 = pack2(-m_ref, -m_ref);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t x2 = fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), scale2, negm2);
            float p0, p1;
            if (((i >> 1) & 7) == 1 || ((i >> 1) & 7) == 4 || ((i >> 1) & 7) == 6) {   // 3 of every 8 pairs: FMA-pipe exp2
              exp2_poly2(x2, p0, p1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            lsum2 = add2(lsum2, pack2(p0, p1));
            pk[c * 16 + (i >> 1)] = cvt_bf16x2(p0, p1);
          }
        }
      }
      }
      const bool narrow = (j == nkv - 1 && ntail < kBKV);
      float l0, l1;
      unpack2(lsum2, l0, l1);
      l_run = l_run * alpha + (l0 + l1);
      ATT_TRACE(3);
      if (j > 0) {   // P_{j-1} must have been consumed (and O_{j-1} accumulated) before P / O are touched
        mbar_wait(bar_o, (j - 1) & 1);
        tc_fence_after();
        if (moved) {   // rare after the first blocks: rescale the TMEM accumulator in place
          const uint64_t alpha2 = pack2(alpha, alpha);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_O + lane_sel + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a, bq;
              unpack2(fma2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2, 0ull), a, bq);
              r[i] = __float_as_uint(a);
              r[i + 1] = __float_as_uint(bq);
            }
            tmem_st_32x32b_x32(tmem_O + lane_sel + c * 32, r);
          }
        }
      }
      ATT_TRACE(4);
      // P (bf16 pairs) -> TMEM columns [192, 256): the A operand of O += P V  (the narrow tail path stored its own)
      if (!narrow) {
        tmem_st_32x32b_x32(tmem_P + lane_sel, pk);
        tmem_st_32x32b_x32(tmem_P + lane_sel + 32, pk + 32);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pready);
      ATT_TRACE(5);
    }
    mbar_wait(bar_o, (nkv - 1) & 1);
    tc_fence_after();

    // ---- epilogue: O (TMEM) -> registers, fold the few trailing keys (if any), normalise, store
    const int q = q0 + tid;
    float o[kHD];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_O + lane_sel + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(r[i]);
    }
    if (tail_keys) {
      float qv[kHD];   // this thread's query row from the swizzled Q tile
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(sQ + tid * 128 + ((c ^ (tid & 7)) << 4)));
        const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
        qv[c * 8 + 0] = a0.x; qv[c * 8 + 1] = a0.y; qv[c * 8 + 2] = a1.x; qv[c * 8 + 3] = a1.y;
        qv[c * 8 + 4] = a2.x; qv[c * 8 + 5] = a2.y; qv[c * 8 + 6] = a3.x; qv[c * 8 + 7] = a3.y;
      }
      for (int t = 0; t < tail_keys; ++t) {
        const __nv_bfloat16* krow = p.qkv + static_cast<int64_t>(row_base + nkv * kBKV + t) * p.ld_qkv;
        float sdot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(krow + colK) + c);
          const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
          sdot = fmaf(qv[c * 8 + 0], a0.x, sdot); sdot = fmaf(qv[c * 8 + 1], a0.y, sdot);
          sdot = fmaf(qv[c * 8 + 2], a1.x, sdot); sdot = fmaf(qv[c * 8 + 3], a1.y, sdot);
          sdot = fmaf(qv[c * 8 + 4], a2.x, sdot); sdot = fmaf(qv[c * 8 + 5], a2.y, sdot);
          sdot = fmaf(qv[c * 8 + 6], a3.x, sdot); sdot = fmaf(qv[c * 8 + 7], a3.y, sdot);
        }
        sdot *= p.scale_log2;
        const float m_new = fmaxf(m_ref, sdot);
        const float a = ex2(m_ref - m_new);
        const float pj = __bfloat162float(__float2bfloat16_rn(ex2(sdot - m_new)));   // same bf16 rounding of P as the MMA path
        m_ref = m_new;
        l_run = l_run * a + pj;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(krow + colV) + c);
          const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
          o[c * 8 + 0] = fmaf(o[c * 8 + 0], a, pj * a0.x); o[c * 8 + 1] = fmaf(o[c * 8 + 1], a, pj * a0.y);
          o[c * 8 + 2] = fmaf(o[c * 8 + 2], a, pj * a1.x); o[c * 8 + 3] = fmaf(o[c * 8 + 3], a, pj * a1.y);
          o[c * 8 + 4] = fmaf(o[c * 8 + 4], a, pj * a2.x); o[c * 8 + 5] = fmaf(o[c * 8 + 5], a, pj * a2.y);
          o[c * 8 + 6] = fmaf(o[c * 8 + 6], a, pj * a3.x); o[c * 8 + 7] = fmaf(o[c * 8 + 7], a, pj * a3.y);
        }
      }
    }
    const float inv = 1.f / l_run;
    if (q < p.N) {
      __nv_bfloat16* op = p.out + static_cast<int64_t>(row_base + q) * p.ld_out + head * kHD;
#pragma unroll
      for (int i = 0; i < kHD; i += 8) {
        uint4 w;
        w.x = cvt_bf16x2(o[i] * inv, o[i + 1] * inv);
        w.y = cvt_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        w.z = cvt_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        w.w = cvt_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(op + i) = w;
      }
    }
    if (q < p.N && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (m_ref + log2f(l_run)) * 0.69314718055994531f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kAttnTmemCols>(tmem_base);
  }
}

// A handful of trailing query rows (N mod 128 <= 8, e.g. the 1025th token of a 32x32-patch image + CLS) would
// otherwise occupy a whole 128-row tensor-core tile per (image, head): one warp per row instead.  Each lane
// walks keys lane, lane+32, ... with its own online softmax (fp32), then the 32 partial states are merged.
constexpr int kTailWarps = 8;
__global__ void __launch_bounds__(kTailWarps * 32)
attn_tail_rows_kernel(AttnParams p, int row0, int nrows) {
  // One CTA per (image, head, row).  8 lanes share a key (16 bytes of the 128-byte K / V row each), so a warp-wide
  // load touches 4 rows x 128 contiguous bytes; the 32 lane-groups of the CTA each run an online softmax over the
  // keys dealt to them and the partial states are merged at the end.
  __shared__ float s_m[kTailWarps * 4], s_l[kTailWarps * 4], s_o[kTailWarps * 4][kHD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, grp = warp * 4 + (lane >> 3);          // 16-byte chunk of the row, key group 0..31
  const int t = blockIdx.x % nrows;
  const int head = (blockIdx.x / nrows) % p.heads;
  const int b = blockIdx.x / (nrows * p.heads);
  const int q = row0 + t;
  const __nv_bfloat16* base = p.qkv + static_cast<int64_t>(b) * p.N * p.ld_qkv + head * kHD + sub * 8;
  float qv[8], o[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(q) * p.ld_qkv));
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
    qv[0] = a0.x * p.scale_log2; qv[1] = a0.y * p.scale_log2; qv[2] = a1.x * p.scale_log2; qv[3] = a1.y * p.scale_log2;
    qv[4] = a2.x * p.scale_log2; qv[5] = a2.y * p.scale_log2; qv[6] = a3.x * p.scale_log2; qv[7] = a3.y * p.scale_log2;
  }
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  const uint32_t gmask = 0xffu << (lane & 24);
  for (int j = grp; j < p.N; j += kTailWarps * 4) {
    const __nv_bfloat16* row = base + static_cast<int64_t>(j) * p.ld_qkv;
    const uint4 ku = __ldg(reinterpret_cast<const uint4*>(row + p.d));
    const uint4 vu = __ldg(reinterpret_cast<const uint4*>(row + 2 * p.d));
    const float2 k0 = unpack_bf16x2(ku.x), k1 = unpack_bf16x2(ku.y), k2 = unpack_bf16x2(ku.z), k3 = unpack_bf16x2(ku.w);
    float sdot = qv[0] * k0.x + qv[1] * k0.y + qv[2] * k1.x + qv[3] * k1.y + qv[4] * k2.x + qv[5] * k2.y + qv[6] * k3.x + qv[7] * k3.y;
    sdot += __shfl_xor_sync(gmask, sdot, 1);   // reduce inside the 8-lane group (groups may leave the loop at
    sdot += __shfl_xor_sync(gmask, sdot, 2);   //  different trip counts, so the mask names only this group)
    sdot += __shfl_xor_sync(gmask, sdot, 4);
    const float m_new = fmaxf(m, sdot);
    const float a = ex2(m - m_new), pj = ex2(sdot - m_new);
    m = m_new;
    l = l * a + pj;
    const float2 v0 = unpack_bf16x2(vu.x), v1 = unpack_bf16x2(vu.y), v2 = unpack_bf16x2(vu.z), v3 = unpack_bf16x2(vu.w);
    o[0] = fmaf(o[0], a, pj * v0.x); o[1] = fmaf(o[1], a, pj * v0.y); o[2] = fmaf(o[2], a, pj * v1.x); o[3] = fmaf(o[3], a, pj * v1.y);
    o[4] = fmaf(o[4], a, pj * v2.x); o[5] = fmaf(o[5], a, pj * v2.y); o[6] = fmaf(o[6], a, pj * v3.x); o[7] = fmaf(o[7], a, pj * v3.y);
  }
  if (sub == 0) { s_m[grp] = m; s_l[grp] = l; }
#pragma unroll
  for (int d = 0; d < 8; ++d) s_o[grp][sub * 8 + d] = o[d];
  __syncthreads();
  if (warp == 0) {
    float mt = -INFINITY;
    for (int g = 0; g < kTailWarps * 4; ++g) mt = fmaxf(mt, s_m[g]);
    float lt = 0.f, o0 = 0.f, o1 = 0.f;
    for (int g = 0; g < kTailWarps * 4; ++g) {
      const float f = (s_m[g] == -INFINITY) ? 0.f : ex2(s_m[g] - mt);
      lt += s_l[g] * f;
      o0 += s_o[g][lane] * f;
      o1 += s_o[g][lane + 32] * f;
    }
    const float inv = 1.f / lt;
    __nv_bfloat16* op = p.out + (static_cast<int64_t>(b) * p.N + q) * p.ld_out + head * kHD;
    op[lane] = __float2bfloat16_rn(o0 * inv);
    op[lane + 32] = __float2bfloat16_rn(o1 * inv);
    if (lane == 0 && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (mt + log2f(lt)) * 0.69314718055994531f;
  }
}

}  // namespace vdr

static unsigned long long* g_attn_trace = nullptr;
extern "C" void vdr_debug_set_attn_trace(void* device_buf) { g_attn_trace = static_cast<unsigned long long*>(device_buf); }

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
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.ld_qkv = ld_qkv;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.ld_out = ld_out;
  p.B = B; p.N = N; p.heads = heads; p.d = d;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  // full 128-row query tiles on the tensor cores; a short tail of rows (<= 8) on a warp-per-row kernel
  const int tail_rows = N % kBQ;
  const bool vector_tail = tail_rows > 0 && tail_rows <= 8;
  const int q_tiles = vector_tail ? N / kBQ : (N + kBQ - 1) / kBQ;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (q_tiles > 0) {
    dim3 grid(q_tiles, heads, B);
    flash_attn_fwd_kernel<<<grid, kAttnThreads, kAttnSmem, s>>>(tm, p);
    count_launch();
    VDR_CHECK_LAUNCH("flash_attn_fwd_kernel");
  }
  if (vector_tail) {
    attn_tail_rows_kernel<<<(unsigned)(B * heads * tail_rows), kTailWarps * 32, 0, s>>>(p, N - tail_rows, tail_rows);
    count_launch();
    VDR_CHECK_LAUNCH("attn_tail_rows_kernel");
  }
  return VDR_OK;
}
