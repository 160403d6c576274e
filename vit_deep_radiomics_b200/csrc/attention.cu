// Fused flash-style self-attention forward on tcgen05 (head_dim 64), sm_100a.
//
// One CTA = 128 query rows of one (batch, head), two CTAs per SM.  Per 128-key block j:
//   S_j = Q K_j^T    tcgen05.mma 128x128x64 (both operands K-major, TMA SWIZZLE_128B tiles) -> TMEM
//   softmax          tcgen05.ld S, online max / sum in registers, P_j (bf16) -> TMEM
//   O  += P_j V_j    tcgen05.mma 128x64x128 (A = P from TMEM, B = V MN-major straight from the TMA tile)
// Warp roles (12 warps):
//   warps 0-7   softmax: warp w owns TMEM lanes 32*(w&3).. (one query row per lane) and the 64 score columns
//               [64*(w>>2), +64) of the block -- TWO threads per query row, so that four softmax warps (two CTAs)
//               are resident per scheduler: the exponentials are latency-bound per warp (MUFU / tcgen05.ld / barrier
//               round trips), not issue-bound, and the extra warps fill those stalls.  The two half-row threads
//               exchange their block maximum through shared memory (one 64-thread named barrier per block).
//   warp 8      K tiles + S MMAs, warp 9: V tiles + PV MMAs (one lane each), warps 10-11 idle (setmaxnreg is per warpgroup)
// Online softmax with LAZY rescaling: probabilities are taken relative to a reference maximum that only moves
// when the running maximum exceeds it by 2^8, so O accumulates in TMEM across key blocks and is rescaled rarely.
// Q, K, V are read in place from the packed QKV GEMM output through one 2-D tensor map.
#include <cstdlib>
#include <type_traits>

#include <cuda_fp16.h>

#include "common.cuh"

#ifndef VDR_ATTN_POLY_MASK
#define VDR_ATTN_POLY_MASK 0x22u   // 2 pairs of every 8.  With maximum-free blocks (N = 1024 alone): 0/8 0.585, 1/8 0.562, 2/8 0.540, 3/8 0.554, 4/8 0.605 ms; bench.py on one box: 1/8 4,336-4,373, 2/8 4,344-4,401, 3/8 4,298-4,316, 4/8 4,191-4,214 slices/s (round 1, with the maximum in every block: 1/8 was the optimum)
#endif

namespace vdr {

constexpr int kSoftmaxWarps = 8;
constexpr int kAttnThreads = (kSoftmaxWarps + 4) * 32;   // 2 softmax warpgroups + issuer warpgroup
constexpr int kBQ = 128, kBKV = 128, kHD = 64;
constexpr int kTileBytes = 128 * kHD * 2;               // 16 KB: one 128 x 64 bf16 tile
constexpr int kAttnSmem = 5 * kTileBytes /*Q, 4 ring slots*/ + 256 /*barriers*/ + 3 * 2 * 128 * 4 /*max exchange x2, sum exchange*/ + 8 * 2 * 128 /*trailing keys: K, V rows*/ + 8 * 128 * 4 /*trailing keys: q . k per query row*/;
constexpr uint32_t kPolyMask = VDR_ATTN_POLY_MASK;        // which of every 8 column pairs take the polynomial exp2 (bit i = pair i)
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

// instruction descriptor for fp16 x fp16 -> f32 (the rel_w accumulate): A / B format fields 0
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

struct AttnParams {
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  __nv_bfloat16* out;
  float* lse;
  int64_t ld_out;
  int B, N, heads, d;
  float scale_log2;
  int no_key_fold;             // 1: a ragged last key block always runs as a narrow MMA block (never folded in the epilogue)
  int dbg;                     // timing experiments only (VDR_ATTN_DBG): 1 = tail CTAs exit at once, 2 = skip the trailing-key fold
  int q_tiles, tail_rows;      // full 128-row query tiles (tensor cores) / trailing rows handled by the extra CUDA-core CTA
  unsigned long long* trace;   // debug (VDR_ATTN_TRACE builds only): per-iteration timestamps of CTA (0,0,0)
  // kBias instantiation only (SAM / MedSAM global attention over an Sh x 64 token grid, N % 128 == 0): additive bias
  // rel[(b*heads + head)*N + q][kh] + rel[...][Sh + kw] for key (kh, kw), already multiplied by log2(e)  (vdr_relpos_tables)
  const float* rel;
  int rel_pitch;               // Sh + 64
  // kFused instantiation: the split [rel_pos_h (2 Sh - 1) ; rel_pos_w (127)] x 64 tables as plain pointers (the kernel reads them through
  // TMA; the exact CUDA-core recomputation of a tile whose maximum-free blocks overflowed reads them directly)
  const __nv_bfloat16 *rcat_hi, *rcat_lo;
  DropSpec drop;               // kDrop instantiation only: attention dropout (thr16 == 0: off)
};
constexpr int kBiasPitch = 136;                          // bytes per query row of the rel_w terms in shared memory: 64 halfs + 8 pad
constexpr int kAttnSmemBias = kAttnSmem + 128 * kBiasPitch;
// kFused: the rel_w terms live in an fp16 UMMA operand tile (128 rows x 64, 1024-byte aligned; it takes the place of the trailing-key
// scratch, which N % 128 == 0 never needs) next to a 64 x 64 fp16 identity tile: S += Qw I^T adds them on the tensor cores
constexpr int kFusedQwOff = 86016, kFusedIdOff = kFusedQwOff + kTileBytes, kAttnSmemFused = kFusedIdOff + 8192;
static_assert(kFusedQwOff >= 5 * kTileBytes + 256 + 3 * 2 * 128 * 4 && kFusedQwOff % 1024 == 0, "fused rel-pos tiles overlap the barriers / exchange buffers");

#ifdef VDR_ATTN_TRACE
__device__ __forceinline__ unsigned long long attn_gtime() {
  unsigned long long t;
#ifdef VDR_ATTN_TRACE_CLOCK   // SM cycles (1-cycle resolution; %globaltimer ticks every 32 ns): all stamps of a CTA come from one SM
  asm volatile("mov.u64 %0, %clock64;" : "=l"(t));
#else
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
#endif
  return t;
}
#define ATT_TRACE(ev)                                                                                              \
  do {                                                                                                             \
    if (p.trace != nullptr && tid == 32 && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 60 && j < 16) p.trace[j * 16 + (ev)] = attn_gtime(); \
  } while (0)
#define ATT_TRACE_AT(row, ev)                                                                                      \
  do {                                                                                                             \
    if (p.trace != nullptr && threadIdx.x == 32 && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 60) p.trace[(row) * 16 + (ev)] = attn_gtime(); \
  } while (0)
#define ISS_TRACE(ev)                                                                                              \
  do {                                                                                                             \
    if (p.trace != nullptr && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 60 && j < 16) p.trace[j * 16 + 8 + (ev)] = attn_gtime(); \
  } while (0)
#else
#define ATT_TRACE(ev) do { } while (0)
#define ATT_TRACE_AT(row, ev) do { } while (0)
#define ISS_TRACE(ev) do { } while (0)
#endif

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
  // clamped on both sides: below, 2^x underflows cleanly; above (maximum-free blocks: x is not bounded by the running maximum any
  // more), the exponent-field addition at the end would wrap into the sign bit -- at 128 the result is >= 2^127 or not finite,
  // either of which trips the epilogue's row-sum check
  x2 = pack2(fminf(fmaxf(x0, -125.f), 128.f), fminf(fmaxf(x1, -125.f), 128.f));
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

// 64-thread named barrier shared by the two warps that own the same 32 query rows (ids 1..4; 0 is __syncthreads)
__device__ __forceinline__ void pair_bar_sync(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}
// Issuer-side wait: back off between polls so the spinning lane does not take issue slots from the softmax warps
// that share its scheduler (the MMAs it issues have a whole softmax block of slack).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 24)) {
      printf("vdr: attention issuer wait timeout block (%d,%d,%d) bar %p parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z, (void*)bar, parity);
      __trap();
    }
  }
}

// A handful of trailing query rows (N mod 128 <= 8, e.g. the 1025th token of a 32x32-patch image + CLS) would otherwise
// occupy a whole 128-row tensor-core tile per (image, head).  They are handled on the CUDA cores by one extra CTA per
// (image, head) of the SAME grid (blockIdx.x == number of full query tiles), so that its K / V reads hit the L2 lines its
// eight sibling tensor-core CTAs are streaming at the same time.  8 lanes share a key (16 bytes of the 128-byte K / V
// row each): a warp-wide load touches 4 rows x 128 contiguous bytes; the 48 lane-groups of the CTA each run an online
// softmax over the keys dealt to them and the partial states are merged through shared memory.
// kBar == 0: called by the whole CTA (__syncthreads); kBar > 0: called by threads 0 .. kThreads - 1 only (named barrier kBar).
// kRel: + the decomposed relative-position bias of a token grid of Sh x 64 (SAM's global-attention blocks), from p.rcat_hi / p.rcat_lo.
// kDropT: attention dropout (p.drop) on the normalised probabilities, as the tensor-core paths apply it.
template <int kThreads, int kBar = 0, bool kRel = false, bool kDropT = false>
__device__ __forceinline__ void attn_tail_rows(const AttnParams& p, float* sm, int head, int b, int row0, int nrows) {
  auto sync = [] {
    if (kBar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(kBar), "n"(kThreads) : "memory");
  };
  constexpr int kGroups = kThreads / 8;                            // key groups of 8 lanes (48 / 20): lane `sub` owns 8 of the 64 dims
  constexpr int kDepth = 5;                                        // keys in flight per group (cp.async ring)
  constexpr uint32_t kStageBytes = kThreads * 16 * 2;              // one 16-byte K piece and one V piece per thread
  float* s_m = sm;                                                 // [kGroups]
  float* s_l = sm + kGroups;                                       // [kGroups]
  float* s_o = sm + 2 * kGroups;                                   // [kGroups][kHD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, grp = warp * 4 + (lane >> 3);          // 16-byte chunk of the row, key group
  const __nv_bfloat16* base = p.qkv + static_cast<int64_t>(b) * p.N * p.ld_qkv + head * kHD + sub * 8;
  const uint32_t gmask = 0xffu << (lane & 24);
  // ring of kDepth stages behind the 16 KB scratch: every thread copies exactly the K / V pieces it consumes itself, so the
  // ring needs no CTA barrier -- it is a register prefetch queue that lives in shared memory (60 KB in flight per CTA).
  // Measured: the trailing row of N = 1025 costs 0.077 ms per launch with this ring and 0.083 ms with the earlier loop of six
  // dependent L2 round trips -- the cost is the ~18 k warp instructions per trailing row (8 lanes per key: unpack, shuffle
  // reduction, two exponentials per key) competing with the co-resident tile CTA, not the load latency.
  const uint32_t ring = smem_u32(sm) + 16384u + static_cast<uint32_t>(tid) * 16u;
  const int nsteps = (p.N + kGroups - 1) / kGroups;
  auto issue = [&](int step) {
    const int j = grp + step * kGroups;
    if (step < nsteps && j < p.N) {
      const __nv_bfloat16* row = base + static_cast<int64_t>(j) * p.ld_qkv;
      const uint32_t dst = ring + static_cast<uint32_t>(step % kDepth) * kStageBytes;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(row + p.d) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + kThreads * 16), "l"(row + 2 * p.d) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int t = 0; t < nrows; ++t) {
    const int q = row0 + t;
#pragma unroll
    for (int s = 0; s < kDepth; ++s) issue(s);
    float qv[8], o[8];
    {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(q) * p.ld_qkv));
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
      qv[0] = a0.x * p.scale_log2; qv[1] = a0.y * p.scale_log2; qv[2] = a1.x * p.scale_log2; qv[3] = a1.y * p.scale_log2;
      qv[4] = a2.x * p.scale_log2; qv[5] = a2.y * p.scale_log2; qv[6] = a3.x * p.scale_log2; qv[7] = a3.y * p.scale_log2;
    }
    float qb[8];                                                     // kRel: the unscaled query (x log2 e) against the table rows
#pragma unroll
    for (int d = 0; d < 8; ++d) qb[d] = kRel ? qv[d] * (1.4426950408889634f / p.scale_log2) : 0.f;
    const int Sh = p.N >> 6, qh = q >> 6, qw = q & 63;
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int step = 0; step < nsteps; ++step) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kDepth - 1) : "memory");   // this step's pieces have landed
      if (grp + step * kGroups < p.N) {   // uniform inside an 8-lane group
        const uint32_t src = ring + static_cast<uint32_t>(step % kDepth) * kStageBytes;
        uint4 ku, vu;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ku.x), "=r"(ku.y), "=r"(ku.z), "=r"(ku.w) : "r"(src) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(vu.x), "=r"(vu.y), "=r"(vu.z), "=r"(vu.w) : "r"(src + kThreads * 16) : "memory");
        const float2 k0 = unpack_bf16x2(ku.x), k1 = unpack_bf16x2(ku.y), k2 = unpack_bf16x2(ku.z), k3 = unpack_bf16x2(ku.w);
        float sdot = qv[0] * k0.x + qv[1] * k0.y + qv[2] * k1.x + qv[3] * k1.y + qv[4] * k2.x + qv[5] * k2.y + qv[6] * k3.x + qv[7] * k3.y;
        if (kRel) {   // rel_h[q, kh] = q . rel_pos_h[qh - kh + Sh - 1], rel_w[q, kw] = q . rel_pos_w[qw - kw + 63]: this lane's 8 dims of both
          const int j = grp + step * kGroups, kh = j >> 6, kw = j & 63;
          const int rows[2] = {qh - kh + Sh - 1, 2 * Sh - 1 + qw - kw + 63};
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p.rcat_hi + static_cast<int64_t>(rows[r]) * kHD + sub * 8));
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p.rcat_lo + static_cast<int64_t>(rows[r]) * kHD + sub * 8));
            const float2 h0 = unpack_bf16x2(hi.x), h1 = unpack_bf16x2(hi.y), h2 = unpack_bf16x2(hi.z), h3 = unpack_bf16x2(hi.w);
            const float2 l0 = unpack_bf16x2(lo.x), l1 = unpack_bf16x2(lo.y), l2 = unpack_bf16x2(lo.z), l3 = unpack_bf16x2(lo.w);
            sdot += qb[0] * (h0.x + l0.x) + qb[1] * (h0.y + l0.y) + qb[2] * (h1.x + l1.x) + qb[3] * (h1.y + l1.y) +
                    qb[4] * (h2.x + l2.x) + qb[5] * (h2.y + l2.y) + qb[6] * (h3.x + l3.x) + qb[7] * (h3.y + l3.y);
          }
        }
        sdot += __shfl_xor_sync(gmask, sdot, 1);   // reduce inside the 8-lane group (groups may skip the last step,
        sdot += __shfl_xor_sync(gmask, sdot, 2);   //  so the mask names only this group)
        sdot += __shfl_xor_sync(gmask, sdot, 4);
        const float m_new = fmaxf(m, sdot);
        const float a = ex2(m - m_new);
        float pj = ex2(sdot - m_new);
        m = m_new;
        l = l * a + pj;                                               // the normaliser keeps every key: dropout follows the softmax
        if (kDropT)
          pj *= drop_factor(p.drop, (static_cast<uint64_t>(b) * p.heads + head) * p.N + q, static_cast<uint32_t>(grp + step * kGroups));
        const float2 v0 = unpack_bf16x2(vu.x), v1 = unpack_bf16x2(vu.y), v2 = unpack_bf16x2(vu.z), v3 = unpack_bf16x2(vu.w);
        o[0] = fmaf(o[0], a, pj * v0.x); o[1] = fmaf(o[1], a, pj * v0.y); o[2] = fmaf(o[2], a, pj * v1.x); o[3] = fmaf(o[3], a, pj * v1.y);
        o[4] = fmaf(o[4], a, pj * v2.x); o[5] = fmaf(o[5], a, pj * v2.y); o[6] = fmaf(o[6], a, pj * v3.x); o[7] = fmaf(o[7], a, pj * v3.y);
      }
      // refill the stage just consumed: issued after the arithmetic above, which needed ku / vu in registers, so the
      // shared-memory reads have completed before the copy can overwrite them
      issue(step + kDepth);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // merge the 48 partial softmax states: global maximum first, then every group rescales its own (l, o) before the sums
    if (sub == 0) s_m[grp] = m;
    sync();
    float mt = -INFINITY;
#pragma unroll 8
    for (int g = 0; g < kGroups; ++g) mt = fmaxf(mt, s_m[g]);
    const float f = (m == -INFINITY) ? 0.f : ex2(m - mt);
    if (sub == 0) s_l[grp] = l * f;
#pragma unroll
    for (int d = 0; d < 8; ++d) s_o[grp * kHD + sub * 8 + d] = o[d] * f;
    sync();
    if (tid < kHD) {
      float lt = 0.f, acc = 0.f;
#pragma unroll 8
      for (int g = 0; g < kGroups; ++g) {
        lt += s_l[g];
        acc += s_o[g * kHD + tid];
      }
      p.out[(static_cast<int64_t>(b) * p.N + q) * p.ld_out + head * kHD + tid] = __float2bfloat16_rn(acc / lt);
      if (tid == 0 && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (mt + log2f(lt)) * 0.69314718055994531f;
    }
    sync();
  }
}

// ---- the same trailing rows with warp-level MMAs (the v5 grid's default).  The CUDA-core routine above spends ~18 k warp
// instructions per trailing row (8 lanes per key: unpack, shuffles, two exponentials per key and lane) -- as many issue slots as a
// whole 128-row tensor-core tile, taken from the MUFU-bound tile CTA that shares the SM.  Here one lane of warp 8 streams the
// head's K / V through the tile CTAs' own TMA ring (same tensor map, same 128 x 64 SWIZZLE_128B tiles, L2 hits on the lines the
// siblings stream) and warps 0-7 each own 16 keys of every 128-key tile:
//   S^T chunk : m16n8k16, A = the trailing query rows (row g of the fragment = trailing row g, rows 8-15 zero), B = K via ldmatrix
//   softmax   : plain online softmax per (warp, row); the accumulator fragment of S IS the A fragment of P (FlashAttention-2's layout identity)
//   O chunk   : A = P (bf16), B = V via ldmatrix.trans
// ~2.6 k warp instructions per (image, head) for up to 8 trailing rows; the 8 per-warp states are merged through shared memory.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// D[rows 0-7] += A[rows 0-7] B  (rows 8-15 of A are zero: their accumulators are neither fed nor kept)
__device__ __forceinline__ void mma_16816_top(float& c0, float& c1, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  [[maybe_unused]] float d2, d3;
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %10, %10};"
               : "+f"(c0), "+f"(c1), "=f"(d2), "=f"(d3) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1), "f"(0.f));
  (void)d2;
  (void)d3;
}
constexpr int kTailEmptyBar = 16;   // bars[16..19]: ring slot consumed (8 warps arrive); the tile CTAs use bars[0..13]

__device__ __forceinline__ void attn_tail_rows_mma(const CUtensorMap* tm, const AttnParams& p, uint8_t* smem, int head, int b, int row0, int nrows) {
  const uint32_t base = smem_u32(smem);
#ifdef VDR_ATTN_TRACE   // timeline of one trailing-row CTA (tools/attn_trace_tail.py): row = key block (15 = CTA-level stamps), warp 0 / producer
#define TAIL_TRACE(row, ev)                                                                                        \
  do {                                                                                                             \
    if (p.trace != nullptr && (threadIdx.x & 31) == 0 && blockIdx.y == 5 && blockIdx.z == 60 && (row) < 16) p.trace[256 + (row) * 16 + (ev)] = attn_gtime(); \
  } while (0)
#else
#define TAIL_TRACE(row, ev) do { } while (0)
#endif
  if (threadIdx.x == 0) TAIL_TRACE(15, 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * kTileBytes);
  // Two ring stages of (K_j, V_j): one barrier pair per key block (both tiles of a block are waited for and released together --
  // the CTA's lifetime is its hand-shakes, and it holds half an SM while it lives)
  uint64_t* full = bars + 1;                                              // [2] stage landed (K and V tile: 32 KB)
  uint64_t* empty = bars + kTailEmptyBar;                                 // [2]
  float* s_m = reinterpret_cast<float*>(smem + 5 * kTileBytes + 256);     // [8 warps][8 rows]
  float* s_l = s_m + 64;
  float* s_o = reinterpret_cast<float*>(smem);                            // the (unused) Q slot: [8 warps][8 rows][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nblk = (p.N + kBKV - 1) / kBKV;
  const int row_base = b * p.N, colK = p.d + head * kHD, colV = 2 * p.d + head * kHD;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 8);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) TAIL_TRACE(15, 1);
  if (warp == 8) {
    // ---- producer: (K_0, V_0), (K_1, V_1), ... through two stages of two ring slots each
    for (int jj = 0; jj < nblk; ++jj) {
      const int st = jj & 1;
      if (jj >= 2) mbar_wait_relaxed(&empty[st], ((jj >> 1) - 1) & 1);
      TAIL_TRACE(jj, 8);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], 2 * kTileBytes);
        tma_load_2d(tm, &full[st], smem + kTileBytes * (1 + 2 * st), colK, row_base + jj * kBKV);
        tma_load_2d(tm, &full[st], smem + kTileBytes * (2 + 2 * st), colV, row_base + jj * kBKV);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // A fragments of the trailing rows: row g of the fragment = trailing row min(g, nrows - 1) (duplicates are never stored)
    uint32_t qa[4][2];
    {
      const __nv_bfloat16* qrow = p.qkv + static_cast<int64_t>(row_base + row0 + min(g, nrows - 1)) * p.ld_qkv + head * kHD + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        qa[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qrow + 16 * ks));
        qa[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 8));
      }
    }
    float oc[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) oc[nt][0] = oc[nt][1] = 0.f;
    float m = -INFINITY, l = 0.f;
    const int mi = lane >> 3, r8 = lane & 7;
    for (int j = 0; j < nblk; ++j) {
      const int key0 = j * kBKV + warp * 16;
      const bool active = key0 < p.N;                                     // warp-uniform
      const int st = j & 1;
      if (warp == 0) TAIL_TRACE(j, 0);
      mbar_wait_relaxed(&full[st], (j >> 1) & 1);
      if (warp == 0) TAIL_TRACE(j, 1);
      uint32_t pa0 = 0u, pa2 = 0u;
      if (active) {
        const uint32_t kt = base + kTileBytes * (1 + 2 * st);
        float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        const int r = warp * 16 + (mi >> 1) * 8 + r8;                     // matrices 0/1: keys 0-7, matrices 2/3: keys 8-15
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b00, b01, b10, b11;
          ldsm_x4(kt + r * 128 + (((2 * ks + (mi & 1)) ^ (r & 7)) << 4), b00, b01, b10, b11);
          mma_16816_top(s[0][0], s[0][1], qa[ks][0], qa[ks][1], b00, b01);
          mma_16816_top(s[1][0], s[1][1], qa[ks][0], qa[ks][1], b10, b11);
        }
        float x[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) x[nt][e] = (key0 + 8 * nt + 2 * t + e < p.N) ? s[nt][e] * p.scale_log2 : -INFINITY;
        float mx = fmaxf(fmaxf(x[0][0], x[0][1]), fmaxf(x[1][0], x[1][1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));             // the chunk holds at least one key: finite
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2(m - m_new);
        m = m_new;
        const float p00 = ex2(x[0][0] - m), p01 = ex2(x[0][1] - m), p10 = ex2(x[1][0] - m), p11 = ex2(x[1][1] - m);
        l = l * alpha + ((p00 + p01) + (p10 + p11));
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          oc[nt][0] *= alpha;
          oc[nt][1] *= alpha;
        }
        pa0 = cvt_bf16x2(p00, p01);
        pa2 = cvt_bf16x2(p10, p11);
      }
      if (warp == 0) TAIL_TRACE(j, 2);
      if (active) {
        const uint32_t vt = base + kTileBytes * (2 + 2 * st);
        const int r = warp * 16 + (mi & 1) * 8 + r8;                      // matrices 0/2: keys 0-7, matrices 1/3: keys 8-15
#pragma unroll
        for (int np = 0; np < 4; ++np) {                                  // 16 output dims per ldmatrix
          uint32_t v0, v1, v2, v3;
          ldsm_x4_trans(vt + r * 128 + (((2 * np + (mi >> 1)) ^ (r & 7)) << 4), v0, v1, v2, v3);
          mma_16816_top(oc[2 * np][0], oc[2 * np][1], pa0, pa2, v0, v1);
          mma_16816_top(oc[2 * np + 1][0], oc[2 * np + 1][1], pa0, pa2, v2, v3);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (warp == 0) TAIL_TRACE(j, 3);
    }
    if (warp == 0) TAIL_TRACE(15, 2);
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (t == 0) {
      s_m[warp * 8 + g] = m;
      s_l[warp * 8 + g] = l;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      *reinterpret_cast<float2*>(s_o + (warp * 8 + g) * kHD + 8 * nt + 2 * t) = make_float2(oc[nt][0], oc[nt][1]);
  }
  __syncthreads();
  if (threadIdx.x == 0) TAIL_TRACE(15, 3);
  // merge the eight per-warp states: thread (row, dim)
  for (int i = tid; i < nrows * kHD; i += kAttnThreads) {
    const int rr = i >> 6, dd = i & 63;
    float mt = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) mt = fmaxf(mt, s_m[w * 8 + rr]);
    float lt = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float f = ex2(s_m[w * 8 + rr] - mt);                          // a warp without keys: m = -inf -> 0
      lt = fmaf(s_l[w * 8 + rr], f, lt);
      acc = fmaf(s_o[(w * 8 + rr) * kHD + dd], f, acc);
    }
    const int q = row0 + rr;
    p.out[(static_cast<int64_t>(b) * p.N + q) * p.ld_out + head * kHD + dd] = __float2bfloat16_rn(acc / lt);
    if (dd == 0 && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (mt + log2f(lt)) * 0.69314718055994531f;
  }
  if (threadIdx.x == 0) TAIL_TRACE(15, 4);
#undef TAIL_TRACE
}

// Pipeline per 128-key block j (no CTA-wide barrier in the loop):
//   issuers: wait S_j consumed -> prefetch K_{j+2}, issue S_{j+1};  wait P_j ready -> issue O += P_j V_j
//   softmax: wait S_j -> registers -> signal "consumed" -> max (exchanged between the two half-row threads) ->
//            exp2 / sum -> wait O_{j-1} -> (rare) rescale O -> write P_j -> signal "ready"
// so the tensor pipe computes S_{j+1} and O_{j-1} while the exponentials of block j are evaluated.
// kBias: the additive decomposed relative-position bias of SAM's global-attention blocks.  A 128-key block is two rows of
// the Sh x 64 token grid, so each softmax thread (one query row, 64 score columns) sees exactly one grid row per block:
// its bias is one scalar rel_h[q, kh] per block (prefetched from global memory a block ahead) plus the 64 rel_w[q, kw]
// of its query row, which are the same for every block and sit in shared memory as fp16 (16 KB per CTA; |rel_w| of a
// few units -> 1e-3 absolute in the exponent, below the bf16 rounding of P).
__device__ __forceinline__ unsigned long long attn_gtime_always() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ unsigned int g_attn_sm_ticket[1024];   // experiment (VDR_ATTN_DBG >= 100): alternate start delay per SM slot

// kFused (with kBias): the bias terms are not read from a table in HBM but computed by the CTA itself before its first key
// block -- T = Q [R_hi ; R_lo]^T on the tensor cores into the (still unused) TMEM columns: q . rel_pos_h[i] for the 64-row
// windows of rel_pos_h that belong to the tile's two grid rows (columns 0..63 / 64..127) and q . rel_pos_w[i] for all 127
// offsets (columns 128..255); the R tiles borrow the K / V ring.  Each softmax thread keeps its row's 64 rel_h terms in a local
// array (one load per key block, a block ahead) and scatters its half of the rel_w terms into the fp16 tile in shared memory.
template <bool kBias, bool kDrop = false, bool kFused = false>
__global__ void __launch_bounds__(kAttnThreads, 2)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmRhi,
                      const __grid_constant__ CUtensorMap tmRlo, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  // layout: Q | ring0..3 | barriers | max exchange [2 buffers][2 halves][128] | sum exchange [2 halves][128]
  const uint32_t sQ = base, sRing = base + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * kTileBytes);
  uint64_t* bar_q = bars;            // Q landed
  uint64_t* bar_kv = bars + 1;       // [4] ring slot landed
  uint64_t* bar_s = bars + 5;        // S_j = Q K_j^T complete            (tcgen05.commit)
  uint64_t* bar_o = bars + 6;        // O += P_j V_j complete             (tcgen05.commit)
  uint64_t* bar_sfree = bars + 7;    // S_j is in registers               (8 softmax warps arrive)
  uint64_t* bar_pready = bars + 8;   // P_j is in TMEM, O rescaled        (8 softmax warps arrive)
  uint64_t* bar_tail = bars + 9;     // trailing keys' K / V rows are in shared memory (warp 10 arrives)
  uint64_t* bar_r = bars + 10;       // kFused: the rel-pos table tiles landed (in the K / V ring)
  uint64_t* bar_t = bars + 11;       // kFused: T = Q R^T complete
  uint64_t* bar_tread = bars + 12;   // kFused: T is in registers / shared memory (8 softmax warps arrive): its columns become S, O, P
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_max = reinterpret_cast<float*>(smem + 5 * kTileBytes + 256);   // [2][2][128]
  float* s_sum = s_max + 2 * 2 * 128;                                      // [2][128]
  uint4* s_tail = reinterpret_cast<uint4*>(s_sum + 2 * 128);               // [tail_keys][K row (8 x 16 B) | V row (8 x 16 B)]
  float* s_dot = reinterpret_cast<float*>(s_tail + 8 * 16);                // [tail_keys][128] q . k of every query row (unscaled)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBQ, head = blockIdx.y, b = blockIdx.z;
  if (static_cast<int>(blockIdx.x) >= p.q_tiles) {   // the extra CTA of this (image, head): trailing query rows on the CUDA cores
    if (p.dbg == 5) attn_tail_rows<kAttnThreads>(p, reinterpret_cast<float*>(smem), head, b, p.N - p.tail_rows, p.tail_rows);   // A/B: the CUDA-core routine
    else if (p.dbg != 1) attn_tail_rows_mma(&tmQKV, p, smem, head, b, p.N - p.tail_rows, p.tail_rows);
    return;
  }
  if (p.dbg == 4) return;                             // timing experiments: only the trailing-row CTAs work
  if (p.dbg >= 100 && tid == 0) {                     // experiment: every other CTA of an SM starts p.dbg ns late (de-phases the co-resident pair)
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (atomicAdd(&g_attn_sm_ticket[smid & 1023], 1u) & 1u) __nanosleep(p.dbg);
  }
  const int row_base = b * p.N;                       // first token row of this image in the qkv matrix
  const int colQ = head * kHD, colK = p.d + head * kHD, colV = 2 * p.d + head * kHD;
  // Key blocks: full 128-key blocks on the tensor cores; a ragged last block either runs as a narrow MMA block
  // (16-column granularity) or, when it holds only a few keys (<= 8, e.g. the 1025th token of a 32x32-patch image
  // + CLS), is folded into the epilogue on the CUDA cores instead of costing every CTA one more pipeline round trip.
  const int nkv_all = (p.N + kBKV - 1) / kBKV;
  const int last_keys = p.N - (nkv_all - 1) * kBKV;              // keys in the last block (1..128)
  const int tail_keys = (last_keys <= 8 && nkv_all > 1 && !p.no_key_fold) ? last_keys : 0;
  const int nkv = tail_keys ? nkv_all - 1 : nkv_all;             // blocks that go through the MMA pipeline
  const int valid_last = tail_keys ? kBKV : last_keys;           // keys in the last MMA block
  const int ntail = (valid_last + 15) & ~15;                     // ... rounded to the MMA's 16-column granularity

  auto issue_tile = [&](int t) {   // ring tile t: even = K block t/2 (slots 0/2), odd = V block t/2 (slots 1/3)
    const int slot = t & 3;
#if defined(VDR_X_NOLOAD)
    mbar_arrive(&bar_kv[slot]);        // experiment: no K / V traffic at all (the MMAs read whatever the slot holds)
#elif defined(VDR_X_SAMELOAD)
    mbar_arrive_expect_tx(&bar_kv[slot], kTileBytes);   // experiment: every CTA streams the same two tiles (L2 hits only)
    tma_load_2d(&tmQKV, &bar_kv[slot], smem + kTileBytes * (1 + slot), (t & 1) ? 2 * p.d : p.d, 0);
#else
    mbar_arrive_expect_tx(&bar_kv[slot], kTileBytes);
    tma_load_2d(&tmQKV, &bar_kv[slot], smem + kTileBytes * (1 + slot), (t & 1) ? colV : colK, row_base + (t >> 1) * kBKV);
#endif
  };
  if (tid == kSoftmaxWarps * 32) {
    // one thread initialises the barriers and immediately starts the first loads (Q, K0, V0, K1, V1), so that their
    // latency overlaps the TMEM allocation and the CTA-wide barrier below
    if (base & 1023u) { printf("vdr: attention smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_sfree, kSoftmaxWarps);
    mbar_init(bar_pready, kSoftmaxWarps);
    mbar_init(bar_tail, 1);
    mbar_init(bar_r, 1);
    mbar_init(bar_t, 1);
    mbar_init(bar_tread, kSoftmaxWarps);
    reinterpret_cast<volatile uint32_t*>(bars)[32] = 0u;   // "a row sum went out of range" (maximum-free blocks), byte 128 of the barrier block
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_q, kTileBytes);
    tma_load_2d(&tmQKV, bar_q, smem, colQ, row_base + q0);
    if (kFused) {
      // ring slot 0: rel_pos_h rows [qh, qh + 64) for the tile's first grid row (hi | lo), slot 1: the same for qh + 1,
      // slots 2 / 3: rel_pos_w (127 rows -> 128) hi / lo.  rel_h[q, kh] = q . rel_pos_h[qh - kh + Sh - 1] = T[Sh - 1 - kh].
      const int qh = q0 >> 6, rw0 = 2 * (p.N >> 6) - 1;
      mbar_arrive_expect_tx(bar_r, 4 * kTileBytes);
      tma_load_2d(&tmRhi, bar_r, smem + kTileBytes, 0, qh);
      tma_load_2d(&tmRlo, bar_r, smem + kTileBytes + 8192, 0, qh);
      tma_load_2d(&tmRhi, bar_r, smem + 2 * kTileBytes, 0, qh + 1);
      tma_load_2d(&tmRlo, bar_r, smem + 2 * kTileBytes + 8192, 0, qh + 1);
      tma_load_2d(&tmRhi, bar_r, smem + 3 * kTileBytes, 0, rw0);
      tma_load_2d(&tmRhi, bar_r, smem + 3 * kTileBytes + 8192, 0, rw0 + 64);
      tma_load_2d(&tmRlo, bar_r, smem + 4 * kTileBytes, 0, rw0);
      tma_load_2d(&tmRlo, bar_r, smem + 4 * kTileBytes + 8192, 0, rw0 + 64);
    } else {
      issue_tile(0);
      issue_tile(1);
      if (nkv > 1) {
        issue_tile(2);
        issue_tile(3);
      }
    }
  }
  if (warp == 0) tmem_alloc<kAttnTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128, tmem_P = tmem_base + 192;

  if (warp >= kSoftmaxWarps) {
    // =============================================================== issuer warpgroup
    // (32 is all the CTA's register pool leaves: 8 x 104 + 4 x 32 warps = the 960 warp-registers of a launch at 80 per thread;
    //  setmaxnreg.inc of the softmax warpgroups never returns if the issuer warps keep more -- tried with 48)
    reg_dec<32>();
    // Two issuing threads so that the two dependent MMA chains of a block (S_{j+1} = Q K^T: 4 steps, O += P V:
    // 8 steps) are dispatched concurrently.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // B (= V) is MN-major
    if (warp == kSoftmaxWarps) {
      // ---- K tiles + S = Q K^T.  The whole warp runs the loop and waits; the tcgen05 / TMA instructions sit under elect.sync
      // (exactly one lane, known to the compiler: operands go to uniform registers directly -- under `lane == 0` every
      // tcgen05.mma was wrapped in an ELECT / BRA.U.ANY waterfall and cost ~0.1 us of issue time, 0.5 us per S block and
      // 0.85 us per PV block in the round-2 traces)
      auto wait_k = [&](int j) { mbar_wait_relaxed(&bar_kv[(2 * j) & 3], ((2 * j) >> 2) & 1); };
      auto issue_s = [&](int j) {    // (K_j has landed: wait_k(j) precedes the wait for the S columns)
        const int t = 2 * j;
        tc_fence_after();
        const uint64_t dq = umma_desc_kmajor_sw128(sQ);
        const uint64_t dk = umma_desc_kmajor_sw128(sRing + (t & 3) * kTileBytes);
        // the last key block only computes the (16-column granular) part of S that has keys behind it
        const uint32_t idesc = (j == nkv - 1) ? umma_idesc_bf16(128, ntail) : idesc_s;
        if (elect_one()) {
#pragma unroll
#if !defined(VDR_X_NOMMA)
          for (int k = 0; k < kHD / 16; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc, k != 0);
#endif
          if (kFused) {   // + rel_w[q, kw] / scale for both grid rows of the block: S[:, 0:64] += Qw I^T, S[:, 64:128] += Qw I^T
            const uint64_t dqw = umma_desc_kmajor_sw128(base + kFusedQwOff), did = umma_desc_kmajor_sw128(base + kFusedIdOff);
            constexpr uint32_t idw = umma_idesc_f16(128, 64);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_ss(tmem_S, dqw + 2 * k, did + 2 * k, idw, 1u);
              umma_ss(tmem_S + 64, dqw + 2 * k, did + 2 * k, idw, 1u);
            }
          }
          umma_commit(bar_s);
        }
        __syncwarp();
      };
      mbar_wait_relaxed(bar_q, 0);
      if (kFused) {
        mbar_wait_relaxed(bar_r, 0);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dq = umma_desc_kmajor_sw128(sQ);
          constexpr uint32_t id64 = umma_idesc_bf16(128, 64), id128 = umma_idesc_bf16(128, 128);
#pragma unroll
          for (int part = 0; part < 2; ++part) {        // hi, then lo into the same accumulators
            const uint64_t da = umma_desc_kmajor_sw128(sRing + part * 8192), db = umma_desc_kmajor_sw128(sRing + kTileBytes + part * 8192);
            const uint64_t dw = umma_desc_kmajor_sw128(sRing + (2 + part) * kTileBytes);
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) {
              umma_ss(tmem_base, dq + 2 * k, da + 2 * k, id64, (part | k) != 0);
              umma_ss(tmem_base + 64, dq + 2 * k, db + 2 * k, id64, (part | k) != 0);
              umma_ss(tmem_base + 128, dq + 2 * k, dw + 2 * k, id128, (part | k) != 0);
            }
          }
          umma_commit(bar_t);
        }
        __syncwarp();
        mbar_wait_relaxed(bar_t, 0);                       // the ring is free again: start the K / V stream
        if (elect_one()) {
          issue_tile(0);
          issue_tile(1);
          if (nkv > 1) {
            issue_tile(2);
            issue_tile(3);
          }
        }
        __syncwarp();
        mbar_wait_relaxed(bar_tread, 0);                   // every softmax warp has taken its terms out of T
        tc_fence_after();
      }
      wait_k(0);
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (lane == 0) ISS_TRACE(0);
        if (j + 1 < nkv) wait_k(j + 1);                    // (long since landed: keeps the path from "S_j consumed" to the MMAs short)
        mbar_wait_relaxed(bar_sfree, j & 1);               // S_j is in registers -> K_j's slot and the S columns are free
        tc_fence_after();
        if (lane == 0) ISS_TRACE(1);
        if (j + 1 < nkv) issue_s(j + 1);
        if (j + 2 < nkv && elect_one()) issue_tile(2 * j + 4);   // K_{j+2} into K_j's slot
        __syncwarp();
        if (lane == 0) ISS_TRACE(2);
      }
    } else if (warp == kSoftmaxWarps + 2) {
      // ---- the few trailing keys (folded in by the softmax epilogue): their K and V rows -> shared memory, early
      for (int i = lane; i < tail_keys * 16; i += 32) {
        const int t = i >> 4, c = i & 15;
        const __nv_bfloat16* krow = p.qkv + static_cast<int64_t>(row_base + nkv * kBKV + t) * p.ld_qkv;
        s_tail[i] = __ldg(reinterpret_cast<const uint4*>(krow + (c < 8 ? colK : colV)) + (c & 7));
      }
      __syncwarp();
      // ... and their scores: this warp is otherwise idle, so it evaluates q . k for the 128 query rows while the key blocks
      // stream through the tensor pipe (the softmax threads used to do it in their epilogue, twice per row: 200 of the 750
      // instructions a softmax warp spends per CTA outside the block loop)
      if (tail_keys > 0) {
        mbar_wait_relaxed(bar_q, 0);
        for (int t = 0; t < tail_keys; ++t) {
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            float acc = 0.f;
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {
              uint4 qu;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qu.x), "=r"(qu.y), "=r"(qu.z), "=r"(qu.w)
                           : "r"(sQ + r * 128 + ((c ^ (r & 7)) << 4)));
              const uint4 u = s_tail[t * 16 + c];
              const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
              const float2 q0v = unpack_bf16x2(qu.x), q1v = unpack_bf16x2(qu.y), q2v = unpack_bf16x2(qu.z), q3v = unpack_bf16x2(qu.w);
              acc = fmaf(q0v.x, a0.x, acc); acc = fmaf(q0v.y, a0.y, acc);
              acc = fmaf(q1v.x, a1.x, acc); acc = fmaf(q1v.y, a1.y, acc);
              acc = fmaf(q2v.x, a2.x, acc); acc = fmaf(q2v.y, a2.y, acc);
              acc = fmaf(q3v.x, a3.x, acc); acc = fmaf(q3v.y, a3.y, acc);
            }
            s_dot[t * 128 + r] = acc;
          }
        }
        __syncwarp();
      }
      if (lane == 0) mbar_arrive(bar_tail);
    } else if (warp == kSoftmaxWarps + 1) {
      // ---- V tiles + O += P V (warp-wide loop, instructions under elect.sync as above)
      for (int j = 0; j < nkv; ++j) {
        if (lane == 0) ISS_TRACE(3);
        const int t = 2 * j + 1;
        mbar_wait_relaxed(&bar_kv[t & 3], (t >> 2) & 1);   // V_j (landed long before P_j is ready)
        mbar_wait_relaxed(bar_pready, j & 1);              // P_j is in TMEM (and O rescaled if the maximum moved)
        tc_fence_after();
        if (lane == 0) ISS_TRACE(4);
        const uint64_t dv0 = umma_desc_mnmajor_sw128(sRing + (t & 3) * kTileBytes);
        // A = P from TMEM (16 bf16 = 8 columns per K step); O accumulates in TMEM across all key blocks.
        // Each K step covers 16 kv rows x 128 B of the V tile = 2048 B = 128 descriptor address units.
        const bool full = (j < nkv - 1 || ntail == kBKV);
        if (elect_one()) {
#if defined(VDR_X_NOMMA)
          if (false) {
#else
          if (full) {
#endif
#pragma unroll
            for (int k = 0; k < kBKV / 16; ++k)
              umma_ts(tmem_O, tmem_P + k * 8, dv0 + static_cast<uint64_t>(k * 128), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          } else if (!full) {
            const int ksteps = ntail / 16;
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k)
              umma_ts(tmem_O, tmem_P + k * 8, dv0 + static_cast<uint64_t>(k * 128), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(bar_o);
        }
        __syncwarp();
        if (lane == 0) ISS_TRACE(5);
        if (j + 2 < nkv) {                                 // V_{j+2} into V_j's slot as soon as O_j is complete: this warp has nothing else
          mbar_wait_relaxed(bar_o, j & 1);                 //  to do until P_{j+1} is ready, most of a block period later
          if (elect_one()) issue_tile(2 * j + 5);
          __syncwarp();
        }
      }
    }
  } else {
    // =============================================================== softmax warpgroups
    reg_inc<104>();
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;                              // query row of this thread inside the tile
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_S + lane_sel + half * 64;                // this thread's 64 score columns
    const uint32_t tP = tmem_P + lane_sel + half * 32;                // ... as 32 packed bf16x2 columns of P
    const uint32_t tO = tmem_O + lane_sel + half * 32;                // the 32 output columns this thread rescales / stores
    float m_ref = -INFINITY, l_run = 0.f;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    const unsigned char* bw_row = smem + kAttnSmem + row * kBiasPitch;     // kBias: this query row's 64 rel_w terms (fp16)
    const float* rel_row = nullptr;
    float bh_next = 0.f;
    float rh_loc[64];                                                      // kFused: this row's rel_h terms (log2 domain), index Sh - 1 - kh
    const int Sh = p.N >> 6;
    if (kFused) {
      mbar_wait(bar_t, 0);
      tc_fence_after();
      const int qw = row & 63;
      unsigned char* qw_tile = smem + kFusedQwOff;
      const float inv_scale = 1.4426950408889634f / p.scale_log2;
      uint32_t u[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32b_x32(tmem_base + lane_sel + (row & 64) + 32 * c, u);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) rh_loc[32 * c + i] = __uint_as_float(u[i]) * 1.4426950408889634f;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32b_x32(tmem_base + lane_sel + 128 + half * 64 + 32 * c, u);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {                                     // rel_w[q, kw] = q . rel_pos_w[qw - kw + 63] = T_w[qw + 63 - kw]
          const int jj = qw + 63 - (half * 64 + 32 * c + i);
          if (static_cast<unsigned>(jj) < 64u)     // element jj of row `row` of the 128-byte-swizzled K-major tile
            *reinterpret_cast<__half*>(qw_tile + row * 128 + (((jj >> 3) ^ (row & 7)) << 4) + 2 * (jj & 7)) = __float2half_rn(__uint_as_float(u[i]) * inv_scale);
        }
      }
      {   // the identity tile: 64 rows x 8 sixteen-byte chunks, two chunks per softmax thread
        const int t0 = warp * 32 + lane;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int idx = t0 + 256 * k, r = idx >> 3, cch = idx & 7;
          uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
          if ((r >> 3) == cch) {
            const uint32_t one = (r & 1) ? 0x3C000000u : 0x00003C00u;      // fp16 1.0 in the element's half of its word
            const int word = (r & 7) >> 1;
            w0 = word == 0 ? one : 0u; w1 = word == 1 ? one : 0u; w2 = word == 2 ? one : 0u; w3 = word == 3 ? one : 0u;
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + kFusedIdOff + r * 128 + ((cch ^ (r & 7)) << 4)), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      pair_bar_sync(quarter);                                              // both halves of the row are written, both have read T
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tread);
      bh_next = rh_loc[Sh - 1 - half];
    } else if (kBias) {
      rel_row = p.rel + (static_cast<int64_t>(b * p.heads + head) * p.N + q0 + row) * p.rel_pitch;
      const float4* src = reinterpret_cast<const float4*>(rel_row + (p.rel_pitch - 64) + half * 32);   // this thread converts 32 of the 64
      uint2* dst = reinterpret_cast<uint2*>(smem + kAttnSmem + row * kBiasPitch + half * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(src + i);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
      bh_next = __ldg(rel_row + half);                                     // grid row of block 0, this half
      pair_bar_sync(quarter);                                              // the other half-row warp wrote the other 32
    }

    // A ragged last query tile (N = 1025: one valid row): warps whose 32 rows are all beyond N only keep the barriers moving --
    // no TMEM reads, no exponentials, no P (their rows of P / O are never stored and do not touch other rows of the MMAs)
    const bool idle_rows = !kBias && q0 + quarter * 32 >= p.N;
    if (idle_rows) {
      for (int j = 0; j < nkv; ++j) {
        mbar_wait_relaxed(bar_s, j & 1);
        if (j > 0) mbar_wait_relaxed(bar_o, (j - 1) & 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_sfree);
          mbar_arrive(bar_pready);
        }
      }
    }
    for (int j = 0; j < (idle_rows ? 0 : nkv); ++j) {
      ATT_TRACE(0);
      const float bh_cur = bh_next;
      if (kFused) {
        if (j + 1 < nkv) bh_next = rh_loc[Sh - 1 - (2 * (j + 1) + half)];
      } else if (kBias && j + 1 < nkv) {
        bh_next = __ldg(rel_row + 2 * (j + 1) + half);
      }
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      ATT_TRACE(1);
      float alpha = 1.f;
      bool moved = false;
      float lsum;
      float* xmax = s_max + (j & 1) * 256;
      const bool narrow = (j == nkv - 1 && valid_last < kBKV);   // (113..127 valid keys: all eight 16-column groups exist, the last one is masked)
      if (narrow) {
        // ---- last, partial key block: only ntail (multiple of 16) columns exist in S; runtime loop over the 16-column
        //      groups of this thread's half
        const int c_lo = half * 64, c_hi = min(ntail, c_lo + 64);
        float mx = -INFINITY;
        for (int c = c_lo; c < c_hi; c += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_S + lane_sel + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c + i < valid_last) ? __uint_as_float(r[i]) : -INFINITY);
        }
        xmax[half * 128 + row] = mx;
        pair_bar_sync(quarter);
        mx = fmaxf(mx, xmax[(half ^ 1) * 128 + row]);
        const float m_new = fmaxf(m_ref, mx * p.scale_log2);
        moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);
        if (moved) {
          alpha = ex2(m_ref - m_new);
          m_ref = m_new;
        }
        lsum = 0.f;
        if (j > 0) {   // P_{j-1} must have been consumed before its columns are overwritten
          mbar_wait(bar_o, (j - 1) & 1);
          tc_fence_after();
        }
        for (int c = c_lo; c < c_hi; c += 16) {
          uint32_t r[16], w[8];
          tmem_ld_32x32b_x16(tmem_S + lane_sel + c, r);
          tmem_ld_wait();
          uint4 db[2];
          if (kDrop) {
            const uint64_t drow = (static_cast<uint64_t>(b) * p.heads + head) * p.N + q0 + row;
            const uint32_t c8 = static_cast<uint32_t>((j * kBKV + c) >> 3);
            db[0] = drop_bits8(p.drop, drow, c8);
            db[1] = drop_bits8(p.drop, drow, c8 + 1);
          }
          const float dsc = kDrop ? drop_scale(p.drop) : 1.f;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = (c + i < valid_last) ? ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_ref)) : 0.f;
            const float p1 = (c + i + 1 < valid_last) ? ex2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_ref)) : 0.f;
            lsum += p0 + p1;                                  // the normaliser keeps every key: dropout follows the softmax
            if (kDrop) {
              const float k0 = drop_lane16(db[i >> 3], i & 7) >= p.drop.thr16 ? dsc : 0.f;
              const float k1 = drop_lane16(db[i >> 3], (i & 7) + 1) >= p.drop.thr16 ? dsc : 0.f;
              w[i >> 1] = cvt_bf16x2(p0 * k0, p1 * k1);
            } else {
              w[i >> 1] = cvt_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x8(tmem_P + lane_sel + (c >> 1), w);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfree);
        l_run = l_run * alpha + lsum;
        if (j > 0 && moved) {
          const uint64_t alpha2 = pack2(alpha, alpha);
          uint32_t r[32];
          tmem_ld_32x32b_x32(tO, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            unpack2(mul2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), a0, a1);
            r[i] = __float_as_uint(a0);
            r[i + 1] = __float_as_uint(a1);
          }
          tmem_st_32x32b_x32(tO, r);
        }
      } else {
        // this thread's 64 score columns -> registers, then hand the S columns back
        uint32_t sr[2][32];
#if defined(VDR_X_FULLMAX)
        float omx = -INFINITY;   // experiment: the maximum of the OTHER half's 64 scores, reduced by this thread itself (no exchange)
        if (!kBias && !kFused) {
          uint32_t t0[32], t1[32];
          tmem_ld_32x32b_x32(tS + 64 - half * 128, t0);
          tmem_ld_32x32b_x32(tS + 96 - half * 128, t1);
          tmem_ld_32x32b_x32(tS, sr[0]);
          tmem_ld_32x32b_x32(tS + 32, sr[1]);
          tmem_ld_wait();
          float n0 = -INFINITY, n1 = -INFINITY, n2 = -INFINITY, n3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            n0 = max3(n0, __uint_as_float(t0[i]), __uint_as_float(t0[i + 1]));
            n1 = max3(n1, __uint_as_float(t0[i + 2]), __uint_as_float(t0[i + 3]));
            n2 = max3(n2, __uint_as_float(t1[i]), __uint_as_float(t1[i + 1]));
            n3 = max3(n3, __uint_as_float(t1[i + 2]), __uint_as_float(t1[i + 3]));
          }
          omx = fmaxf(fmaxf(n0, n1), fmaxf(n2, n3));
        } else {
          tmem_ld_32x32b_x32(tS, sr[0]);
          tmem_ld_32x32b_x32(tS + 32, sr[1]);
          tmem_ld_wait();
        }
#else
        tmem_ld_32x32b_x32(tS, sr[0]);
        tmem_ld_32x32b_x32(tS + 32, sr[1]);
        tmem_ld_wait();
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfree);
        ATT_TRACE(2);
        if (kBias && !kFused) {   // scores -> log2 domain with the bias added: v = S * scale + (rel_h + rel_w)
          const uint64_t bh2 = pack2(bh_cur, bh_cur);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const uint2 hb = *reinterpret_cast<const uint2*>(bw_row + (c * 32 + i) * 2);
              const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&hb.x));
              const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&hb.y));
              float v0, v1, v2, v3;
              unpack2(fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), scale2, add2(pack2(f0.x, f0.y), bh2)), v0, v1);
              unpack2(fma2(pack2(__uint_as_float(sr[c][i + 2]), __uint_as_float(sr[c][i + 3])), scale2, add2(pack2(f1.x, f1.y), bh2)), v2, v3);
              sr[c][i] = __float_as_uint(v0); sr[c][i + 1] = __float_as_uint(v1);
              sr[c][i + 2] = __float_as_uint(v2); sr[c][i + 3] = __float_as_uint(v3);
            }
          }
        }
#if defined(VDR_X_NOSOFTMAX)
        if (!kBias && !kFused && !kDrop) {   // experiment: the pipeline skeleton alone (S -> registers -> P without any arithmetic)
          if (j > 0) {
            mbar_wait(bar_o, (j - 1) & 1);
            tc_fence_after();
          }
          tmem_st_32x32b_x32(tP, sr[0]);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pready);
          continue;
        }
#endif
        // Maximum-free blocks (plain kernel): the reference maximum m_ref is the row maximum of block 0; later blocks skip the 44 max
        // instructions, the exchange between the two half-row threads and the vote (0.598 -> 0.560 ms at N = 1024).  P = exp2(s - m_ref)
        // may then exceed 1, which is harmless while nothing overflows: P is bf16 with the fp32 exponent range, O and the row sums are
        // fp32 and are normalised by the same sum at the end.  The epilogue checks every row's sum (finite and < 2^100); a CTA with a
        // row out of range -- a score more than ~88 nats above its row's block-0 maximum -- recomputes its tile exactly on the CUDA
        // cores (attn_tail_rows) before it exits.  A guard per block (redo from the scores in registers) was measured first: keeping the
        // 64 scores live next to the 32 packed P registers spills 16 STL.64 + 32 LDL per block and costs what the maximum did.
        const bool do_max = (j == 0) || (kBias && !kFused);   // (the plain, dropout and fused instantiations run maximum-free blocks)
        if (do_max) {
        // block maximum of this half row (four independent chains), exchanged with the other half-row thread
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = max3(mx0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
          mx1 = max3(mx1, __uint_as_float(sr[0][i + 2]), __uint_as_float(sr[0][i + 3]));
          mx2 = max3(mx2, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
          mx3 = max3(mx3, __uint_as_float(sr[1][i + 2]), __uint_as_float(sr[1][i + 3]));
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        // kFused: the scores already hold rel_w / scale (added by the MMA); rel_h of this half's grid row is one scalar per row --
        // the two half-row threads have DIFFERENT rel_h terms, so they exchange the maximum with the term added (log2 domain)
        if (kFused) mx = fmaf(mx, p.scale_log2, bh_cur);
#if defined(VDR_X_FULLMAX)
        if (!kBias && !kFused) {
          mx = fmaxf(mx, omx);
        } else
#endif
#if defined(VDR_X_NOEXCH)
        if (kBias)
#endif
        {
        xmax[half * 128 + row] = mx;
        pair_bar_sync(quarter);
        mx = fmaxf(mx, xmax[(half ^ 1) * 128 + row]);
        }
        const float m_new = fmaxf(m_ref, (kBias || kFused) ? mx : mx * p.scale_log2);
        moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);   // warp-uniform (TMEM accesses are warp-wide) and, since both
        if (moved) {                                             //  half-row warps see the same row maxima, CTA-pair-uniform
          alpha = ex2(m_ref - m_new);
          m_ref = m_new;
        }
        }
        const uint64_t negm2 = kFused ? pack2(bh_cur - m_ref, bh_cur - m_ref) : pack2(-m_ref, -m_ref);
        uint64_t lsum2 = 0ull;
        uint32_t pk[32];
        uint4 dbits = make_uint4(0, 0, 0, 0);
        const uint64_t drop_row = (static_cast<uint64_t>(b) * p.heads + head) * p.N + q0 + row;
        const float dsc = kDrop ? drop_scale(p.drop) : 1.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t s2 = pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1]));
            const uint64_t x2 = (kBias && !kFused) ? add2(s2, negm2) : fma2(s2, scale2, negm2);
            float p0, p1;
#if defined(VDR_X_NOMATH)
            if (true) {
              unpack2(x2, p0, p1);
            } else
#endif
            if ((kPolyMask >> ((i >> 1) & 7)) & 1u) {   // a fixed subset of every 8 pairs: FMA-pipe exp2
              exp2_poly2(x2, p0, p1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            lsum2 = add2(lsum2, pack2(p0, p1));                  // the normaliser keeps every key: dropout follows the softmax
            if (kDrop) {
              if ((i & 7) == 0)                                  // eight consecutive key columns per Philox call
                dbits = drop_bits8(p.drop, drop_row, static_cast<uint32_t>((j * kBKV + half * 64 + c * 32 + i) >> 3));
              const float k0 = drop_lane16(dbits, i & 7) >= p.drop.thr16 ? dsc : 0.f;
              const float k1 = drop_lane16(dbits, (i & 7) + 1) >= p.drop.thr16 ? dsc : 0.f;
              pk[c * 16 + (i >> 1)] = cvt_bf16x2(p0 * k0, p1 * k1);
            } else {
              pk[c * 16 + (i >> 1)] = cvt_bf16x2(p0, p1);
            }
          }
        }
        float l0, l1;
        unpack2(lsum2, l0, l1);
        l_run = l_run * alpha + (l0 + l1);
        ATT_TRACE(3);
        if (j > 0) {   // P_{j-1} must have been consumed (and O_{j-1} accumulated) before P / O are touched
          mbar_wait(bar_o, (j - 1) & 1);
          tc_fence_after();
          if (moved) {   // rare after the first blocks: rescale this thread's 32 columns of the TMEM accumulator in place
            const uint64_t alpha2 = pack2(alpha, alpha);
            uint32_t r[32];
            tmem_ld_32x32b_x32(tO, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a0, a1;
              unpack2(mul2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), a0, a1);
              r[i] = __float_as_uint(a0);
              r[i + 1] = __float_as_uint(a1);
            }
            tmem_st_32x32b_x32(tO, r);
          }
        }
        ATT_TRACE(4);
        tmem_st_32x32b_x32(tP, pk);   // P (bf16 pairs): the A operand of O += P V
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pready);
      ATT_TRACE(5);
    }
    mbar_wait(bar_o, (nkv - 1) & 1);
    tc_fence_after();
    bool bad_row = false;
    if (!idle_rows) {
    // ---- epilogue: this thread's 32 columns of O (TMEM) -> registers, fold the few trailing keys (if any),
    //      combine the two half-row sums, normalise, store
    const int q = q0 + row;
    float o[32];
    {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tO, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]);
    }
    if (tail_keys && p.dbg != 2) {
      mbar_wait(bar_tail, 0);
      // both half-row threads read the same score (-> identical m_ref / sums) and fold the key's V into their own 32 columns;
      // the trailing key's probability is added to half 0's sum only
      for (int t = 0; t < tail_keys; ++t) {
        const float sdot = s_dot[t * 128 + row] * p.scale_log2;   // q . k from the prefetch warp
        const float m_new = fmaxf(m_ref, sdot);
        const float a = ex2(m_ref - m_new);
        const float pj = __bfloat162float(__float2bfloat16_rn(ex2(sdot - m_new)));   // same bf16 rounding of P as the MMA path
        m_ref = m_new;
        l_run = l_run * a + (half == 0 ? pj : 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 u = s_tail[t * 16 + 8 + half * 4 + c];
          const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
          o[c * 8 + 0] = fmaf(o[c * 8 + 0], a, pj * a0.x); o[c * 8 + 1] = fmaf(o[c * 8 + 1], a, pj * a0.y);
          o[c * 8 + 2] = fmaf(o[c * 8 + 2], a, pj * a1.x); o[c * 8 + 3] = fmaf(o[c * 8 + 3], a, pj * a1.y);
          o[c * 8 + 4] = fmaf(o[c * 8 + 4], a, pj * a2.x); o[c * 8 + 5] = fmaf(o[c * 8 + 5], a, pj * a2.y);
          o[c * 8 + 6] = fmaf(o[c * 8 + 6], a, pj * a3.x); o[c * 8 + 7] = fmaf(o[c * 8 + 7], a, pj * a3.y);
        }
      }
    }
    s_sum[half * 128 + row] = l_run;
    pair_bar_sync(quarter);
    const float l_tot = l_run + s_sum[(half ^ 1) * 128 + row];
    const float inv = 1.f / l_tot;
    bad_row = q < p.N && !(l_tot < 0x1p100f);          // (also NaN / +inf: an exponential of a maximum-free block overflowed)
    if (q < p.N) {
      __nv_bfloat16* op = p.out + static_cast<int64_t>(row_base + q) * p.ld_out + head * kHD + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 w;
        w.x = cvt_bf16x2(o[i] * inv, o[i + 1] * inv);
        w.y = cvt_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        w.z = cvt_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        w.w = cvt_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(op + i) = w;
      }
      if (half == 0 && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (m_ref + log2f(l_tot)) * 0.69314718055994531f;
    }
    }   // !idle_rows
    if (!kBias || kFused) {
      // Maximum-free blocks: a row sum out of range means P overflowed somewhere in this tile.  The eight softmax warps agree through
      // shared memory and recompute the tile's rows exactly on the CUDA cores (K / V straight from global memory; every TMA load and
      // MMA of this CTA has completed, its shared memory is free).  Never taken for scores within ~88 nats of a row's block-0 maximum.
      volatile uint32_t* s_flag = reinterpret_cast<volatile uint32_t*>(bars) + 32;
      if (__any_sync(0xffffffffu, bad_row) && lane == 0) *s_flag = 1u;
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (*s_flag != 0u) attn_tail_rows<kSoftmaxWarps * 32, 5, kFused, kDrop>(p, reinterpret_cast<float*>(smem), head, b, q0, min(kBQ, p.N - q0));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kAttnTmemCols>(tmem_base);
  }
}

// ================================================================================================ v6
// Four light CTAs per SM instead of two heavy ones.  The v5 profile (profiles/r02_attention.md) shows every pipe about half
// busy -- MUFU 46 %, issue 50 %, tensor 25 % -- because each CTA is a serial latency chain (S ready -> tcgen05.ld -> max ->
// exp2 -> P store -> PV) and two chains per SM cannot cover each other's gaps.  v6 makes the chain as short and as simple as
// possible and runs FOUR of them per SM (four independent phases per scheduler):
//   * 64-key blocks: S = Q K_j^T is 128 x 64 (64 TMEM columns), one softmax THREAD per query row (64 scores in registers, no
//     exchange of maxima between threads, no pair barrier);
//   * P_j (bf16 pairs, 32 columns) is written over the first half of the S columns it was computed from; O (64 columns)
//     accumulates next to it: 128 TMEM columns per CTA = four CTAs per SM;
//   * 5 warps: 0-3 softmax, 4 = one lane that issues the TMA loads (two K and two V slots of 8 KB) and every MMA;
//     96 registers per thread (4 x 160 threads x 96 = the register file), 49.5 KB of shared memory.
// Per block the issuer waits for PV_{j-1} (the S / P columns are free again), issues S_j, waits for the four softmax warps'
// "P_j stored", issues O += P_j V_j.  Overlap comes from the other three CTAs of the SM.

constexpr int kV6Threads = 160;
constexpr int kV6BK = 64;                                  // keys per block
constexpr int kV6KV = kV6BK * kHD * 2;                     // 8 KB: one 64 x 64 bf16 K or V tile
constexpr int kV6Smem = kTileBytes + 4 * kV6KV + 256;      // Q | K0 K1 | V0 V1 | barriers
constexpr int kV6TmemCols = 128;                           // S (f32) / P (bf16 pairs in its first 32 columns): [0,64)   O: [64,128)

__global__ void __launch_bounds__(kV6Threads, 4)
flash_attn_fwd_v6_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sQ = base, sK = base + kTileBytes, sV = sK + 2 * kV6KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBytes + 4 * kV6KV);
  uint64_t* bar_q = bars;          // Q landed (two 64-row boxes)
  uint64_t* bar_k = bars + 1;      // [2] K slot landed
  uint64_t* bar_v = bars + 3;      // [2] V slot landed
  uint64_t* bar_s = bars + 5;      // S_j complete                      (tcgen05.commit)
  uint64_t* bar_o = bars + 6;      // O += P_j V_j complete             (tcgen05.commit)
  uint64_t* bar_p = bars + 7;      // P_j stored, O rescaled            (4 softmax warps arrive)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBQ, head = blockIdx.y, b = blockIdx.z;
  const int row_base = b * p.N;
  const int colQ = head * kHD, colK = p.d + head * kHD, colV = 2 * p.d + head * kHD;
  // Key blocks: full 64-key blocks on the tensor cores; a ragged last block runs as a narrow MMA block (16-column granularity)
  // or, when it holds only a few keys (<= 8: the 1025th token of a 32x32-patch image + CLS), is folded into the epilogue on the
  // CUDA cores instead of costing every CTA one more round trip of the serial chain.
  const int nkv_all = (p.N + kV6BK - 1) / kV6BK;
  const int last_keys = p.N - (nkv_all - 1) * kV6BK;             // keys in the last block (1..64)
  const int tail_keys = (last_keys <= 8 && nkv_all > 1) ? last_keys : 0;
  const int nkv = tail_keys ? nkv_all - 1 : nkv_all;             // blocks that go through the MMA pipeline
  const int valid_last = tail_keys ? kV6BK : last_keys;          // keys in the last MMA block
  const int ntail = (valid_last + 15) & ~15;                     // ... rounded to the MMA's 16-column granularity

  if (tid == 128) {
    if (base & 1023u) { printf("vdr: attention smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1);
    mbar_init(&bar_k[0], 1); mbar_init(&bar_k[1], 1);
    mbar_init(&bar_v[0], 1); mbar_init(&bar_v[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_p, 4);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_q, kTileBytes);
    tma_load_2d(&tmQKV, bar_q, smem, colQ, row_base + q0);
    tma_load_2d(&tmQKV, bar_q, smem + kV6KV, colQ, row_base + q0 + 64);
    const int pre = nkv > 1 ? 2 : 1;
    for (int j = 0; j < pre; ++j) {
      mbar_arrive_expect_tx(&bar_k[j], kV6KV);
      tma_load_2d(&tmQKV, &bar_k[j], smem + kTileBytes + j * kV6KV, colK, row_base + j * kV6BK);
      mbar_arrive_expect_tx(&bar_v[j], kV6KV);
      tma_load_2d(&tmQKV, &bar_v[j], smem + kTileBytes + (2 + j) * kV6KV, colV, row_base + j * kV6BK);
    }
  }
  if (warp == 0) tmem_alloc<kV6TmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;

  if (warp == 4) {
    // =============================================================== loads + MMA issue
    // The whole warp runs the loop and waits on the barriers; the tcgen05 / TMA instructions sit under elect.sync, which the
    // compiler understands as "exactly one lane": their operands go to uniform registers directly.  (Under `lane == 0` it wraps
    // every tcgen05.mma in an ELECT / BRA.U.ANY waterfall of ~10 dependent instructions -- ~0.1 us per MMA in the v5 traces,
    // and the issue latency is on this kernel's critical path twice per block.)
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kV6BK, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHD, 0, 1);   // B (= V) is MN-major
      const uint64_t dq = umma_desc_kmajor_sw128(sQ);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nkv; ++j) {
        const int slot = j & 1, ph = (j >> 1) & 1;
        const bool last = j == nkv - 1;
        if (lane == 0) ISS_TRACE(0);
        mbar_wait(&bar_k[slot], ph);
        if (lane == 0) ISS_TRACE(1);
        if (j > 0) mbar_wait(bar_o, (j - 1) & 1);          // PV_{j-1} complete: the S / P columns and V_{j-1}'s slot are free
        tc_fence_after();
        if (lane == 0) ISS_TRACE(2);
        const uint64_t dk = umma_desc_kmajor_sw128(sK + slot * kV6KV);
        const uint32_t idesc = last ? umma_idesc_bf16(128, ntail) : idesc_s;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc, k != 0);
          umma_commit(bar_s);
          if (j >= 1 && j + 1 < nkv) {                     // V_{j+1} into V_{j-1}'s slot (PV_{j-1} is complete)
            const int s1 = (j + 1) & 1;
            mbar_arrive_expect_tx(&bar_v[s1], kV6KV);
            tma_load_2d(&tmQKV, &bar_v[s1], smem + kTileBytes + (2 + s1) * kV6KV, colV, row_base + (j + 1) * kV6BK);
          }
        }
        __syncwarp();
        if (lane == 0) ISS_TRACE(3);
        mbar_wait(bar_p, j & 1);                           // P_j stored (so S_j has been read), O rescaled
        if (lane == 0) ISS_TRACE(4);
        mbar_wait(&bar_v[slot], ph);
        tc_fence_after();
        if (lane == 0) ISS_TRACE(5);
        const uint64_t dv = umma_desc_mnmajor_sw128(sV + slot * kV6KV);
        const int ksteps = last ? ntail / 16 : kV6BK / 16;
        if (elect_one()) {
          // A = P from TMEM (16 bf16 = 8 columns per K step); each K step covers 16 kv rows x 128 B of the V tile = 128 address units
          if (ksteps == kV6BK / 16) {
#pragma unroll
            for (int k = 0; k < kV6BK / 16; ++k)
              umma_ts(tmem_O, tmem_S + k * 8, dv + static_cast<uint64_t>(k * 128), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          } else {
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k)
              umma_ts(tmem_O, tmem_S + k * 8, dv + static_cast<uint64_t>(k * 128), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(bar_o);
          if (j + 2 < nkv) {                               // K_{j+2} into K_j's slot (S_j has been read: bar_p(j))
            mbar_arrive_expect_tx(&bar_k[slot], kV6KV);
            tma_load_2d(&tmQKV, &bar_k[slot], smem + kTileBytes + slot * kV6KV, colK, row_base + (j + 2) * kV6BK);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =============================================================== softmax: thread = query row
    const int row = warp * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_S + lane_sel, tO = tmem_O + lane_sel;
    float m_ref = -INFINITY, l_run = 0.f;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < nkv; ++j) {
      ATT_TRACE(0);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      ATT_TRACE(1);
      uint32_t sr[2][32];
      tmem_ld_32x32b_x32(tS, sr[0]);
      tmem_ld_32x32b_x32(tS + 32, sr[1]);
      tmem_ld_wait();
      ATT_TRACE(2);
      if (j == nkv - 1 && valid_last < kV6BK) {   // ragged last block: columns >= valid_last are other tokens / were never written
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid_last) sr[c][i] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = max3(mx0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
        mx1 = max3(mx1, __uint_as_float(sr[0][i + 2]), __uint_as_float(sr[0][i + 3]));
        mx2 = max3(mx2, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
        mx3 = max3(mx3, __uint_as_float(sr[1][i + 2]), __uint_as_float(sr[1][i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m_ref, mx * p.scale_log2);
      // lazy rescaling: the reference maximum only moves when exceeded by 2^8; warp-uniform because TMEM accesses are warp-wide
      const bool moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);
      float alpha = 1.f;
      if (moved) {
        alpha = ex2(m_ref - m_new);
        m_ref = m_new;
      }
      const uint64_t negm2 = pack2(-m_ref, -m_ref);
      uint64_t lsum2 = 0ull;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), scale2, negm2);
          float p0, p1;
          if ((kPolyMask >> ((i >> 1) & 7)) & 1u) {   // a fixed subset of every 8 pairs: FMA-pipe exp2
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
      float l0, l1;
      unpack2(lsum2, l0, l1);
      l_run = l_run * alpha + (l0 + l1);
      ATT_TRACE(3);
      if (j > 0 && moved) {   // rare after the first blocks.  PV_{j-1} is complete: S_j was issued behind it
        const uint64_t alpha2 = pack2(alpha, alpha);
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tO + hlf * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            unpack2(mul2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), a0, a1);
            r[i] = __float_as_uint(a0);
            r[i + 1] = __float_as_uint(a1);
          }
          tmem_st_32x32b_x32(tO + hlf * 32, r);
        }
      }
      tmem_st_32x32b_x32(tS, pk);   // P_j over the first 32 of the S columns: the A operand of O += P V
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      ATT_TRACE(4);
    }
    mbar_wait(bar_o, (nkv - 1) & 1);
    tc_fence_after();
    // ---- epilogue: fold the few trailing keys (if any), normalise this row's 64 output columns and store them (128
    //      contiguous bytes per thread)
    const int q = q0 + row;
    const __nv_bfloat16* ktail = p.qkv + static_cast<int64_t>(row_base + nkv * kV6BK) * p.ld_qkv + colK;
    const __nv_bfloat16* vtail = ktail + p.d;                  // colV = colK + d
    __nv_bfloat16* op = p.out + static_cast<int64_t>(row_base + q) * p.ld_out + head * kHD;
    float m_fin = m_ref, l_fin = l_run;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tO + hlf * 32, r);
      tmem_ld_wait();
      float m = m_ref, l = l_run;                              // both halves replay the same (m, l) recurrence: no arrays across the loop
      for (int t = 0; t < tail_keys; ++t) {
        // q . k for this thread's query row: Q row from the (swizzled) shared-memory tile, K / V rows straight from global memory
        // (the same 128 bytes for every thread of the CTA: one L1 line)
        uint4 ku[8], vu[4];
#pragma unroll
        for (int c = 0; c < 8; ++c) ku[c] = __ldg(reinterpret_cast<const uint4*>(ktail + static_cast<int64_t>(t) * p.ld_qkv) + c);
#pragma unroll
        for (int c = 0; c < 4; ++c) vu[c] = __ldg(reinterpret_cast<const uint4*>(vtail + static_cast<int64_t>(t) * p.ld_qkv + hlf * 32) + c);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 qu;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qu.x), "=r"(qu.y), "=r"(qu.z), "=r"(qu.w)
                       : "r"(sQ + row * 128 + ((c ^ (row & 7)) << 4)));
          const float2 q0v = unpack_bf16x2(qu.x), q1v = unpack_bf16x2(qu.y), q2v = unpack_bf16x2(qu.z), q3v = unpack_bf16x2(qu.w);
          const float2 a0 = unpack_bf16x2(ku[c].x), a1 = unpack_bf16x2(ku[c].y), a2 = unpack_bf16x2(ku[c].z), a3 = unpack_bf16x2(ku[c].w);
          d0 = fmaf(q0v.x, a0.x, d0); d1 = fmaf(q0v.y, a0.y, d1);
          d0 = fmaf(q1v.x, a1.x, d0); d1 = fmaf(q1v.y, a1.y, d1);
          d0 = fmaf(q2v.x, a2.x, d0); d1 = fmaf(q2v.y, a2.y, d1);
          d0 = fmaf(q3v.x, a3.x, d0); d1 = fmaf(q3v.y, a3.y, d1);
        }
        const float sdot = (d0 + d1) * p.scale_log2;
        const float m_new = fmaxf(m, sdot);
        const float a = ex2(m - m_new);
        const float pp = __bfloat162float(__float2bfloat16_rn(ex2(sdot - m_new)));   // same bf16 rounding of P as the MMA path
        m = m_new;
        l = l * a + pp;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 v0 = unpack_bf16x2(vu[c].x), v1 = unpack_bf16x2(vu[c].y), v2 = unpack_bf16x2(vu[c].z), v3 = unpack_bf16x2(vu[c].w);
          r[c * 8 + 0] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 0]), a, pp * v0.x));
          r[c * 8 + 1] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 1]), a, pp * v0.y));
          r[c * 8 + 2] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 2]), a, pp * v1.x));
          r[c * 8 + 3] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 3]), a, pp * v1.y));
          r[c * 8 + 4] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 4]), a, pp * v2.x));
          r[c * 8 + 5] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 5]), a, pp * v2.y));
          r[c * 8 + 6] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 6]), a, pp * v3.x));
          r[c * 8 + 7] = __float_as_uint(fmaf(__uint_as_float(r[c * 8 + 7]), a, pp * v3.y));
        }
      }
      m_fin = m;
      l_fin = l;
      const float inv = 1.f / l;
      if (q < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = cvt_bf16x2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
          w.y = cvt_bf16x2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
          w.z = cvt_bf16x2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
          w.w = cvt_bf16x2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + hlf * 32 + i) = w;
        }
      }
    }
    m_ref = m_fin;
    l_run = l_fin;
    if (q < p.N && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (m_ref + log2f(l_run)) * 0.69314718055994531f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kV6TmemCols>(tmem_base);
  }
}

// ================================================================================================ v7: persistent tiles
// The v5 timeline of one CTA (tools/attn_trace.py, B 120 / N 1024 / 12 heads, alone): 17.3 us of life for eight key blocks that
// take 1.5 us each in the steady state -- 0.6 us of set-up (barriers, TMEM allocation), 1.0 us until Q / K_0 have arrived, a slow
// first block, 0.8 us for the last PV to drain, 1.1 us of epilogue and the CTA exit / relaunch.  v7 runs the SAME per-block
// pipeline (same warp roles, same TMEM columns, same softmax arithmetic as the plain v5 instantiation) but one resident CTA walks
// over query tiles w = blockIdx.x, + gridDim.x, ...; every counter of the block pipeline (ring slot, barrier phase) simply keeps
// running across the tile boundary, so the K / V tiles, the Q tile (double-buffered) and S_0 of the next tile are in flight while
// the softmax warps finish the current one; TMEM and the barriers are set up once.
//   O is safe across the boundary without a new barrier: PV_0 of the next tile is issued after "P_0 stored", which every softmax
//   warp signals only after its epilogue has read the previous O out of TMEM.
//   Q buffer (it & 1) is refilled with tile it + 2 after "S_last in registers" of tile it; the trailing-key warp's read of that
//   buffer is ordered before it because the softmax warps wait for its "tail ready" before they signal that block.
// The few trailing query rows (N mod 128 <= 8) run as a second, small kernel (flash_attn_tail_kernel: the mma.sync routine above).
#ifdef VDR_ATTN_TRACE
#define V7_TRACE(ev)                                                                                               \
  do {                                                                                                             \
    if (p.trace != nullptr && (threadIdx.x & 31) == 0 && blockIdx.x == 100 && g >= 16 && g < 32) p.trace[(g - 16) * 16 + (ev)] = attn_gtime(); \
  } while (0)
#else
#define V7_TRACE(ev) do { } while (0)
#endif
constexpr int kV7Smem = 6 * kTileBytes /*Q0, Q1, 4 ring slots*/ + 256 + 3 * 2 * 128 * 4 + 8 * 2 * 128 + 8 * 128 * 4;

__global__ void __launch_bounds__(kAttnThreads, 2)
flash_attn_fwd_v7_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sRing = base + 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);
  uint64_t* bar_q = bars;            // [2] Q buffer landed
  uint64_t* bar_kv = bars + 2;       // [4] ring slot landed
  uint64_t* bar_s = bars + 6;        // S_g complete
  uint64_t* bar_o = bars + 7;        // O += P_g V_g complete
  uint64_t* bar_sfree = bars + 8;    // S_g in registers (8 warps)
  uint64_t* bar_pready = bars + 9;   // P_g in TMEM (8 warps)
  uint64_t* bar_tail = bars + 10;    // trailing keys' rows + scores of tile it ready (warp 10)
  uint64_t* bar_tailfree = bars + 11;   // ... consumed by the epilogue of tile it (8 warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
  float* s_max = reinterpret_cast<float*>(smem + 6 * kTileBytes + 256);   // [2][2][128]
  float* s_sum = s_max + 2 * 2 * 128;                                      // [2][128]
  uint4* s_tail = reinterpret_cast<uint4*>(s_sum + 2 * 128);               // [tail_keys][K row | V row]
  float* s_dot = reinterpret_cast<float*>(s_tail + 8 * 16);                // [tail_keys][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = p.B * p.heads * p.q_tiles;
  const int my_tiles = (static_cast<int>(blockIdx.x) < total) ? (total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  const int nkv_all = (p.N + kBKV - 1) / kBKV;
  const int last_keys = p.N - (nkv_all - 1) * kBKV;
  const int tail_keys = (last_keys <= 8 && nkv_all > 1 && !p.no_key_fold) ? last_keys : 0;
  const int nkv = tail_keys ? nkv_all - 1 : nkv_all;
  const int valid_last = tail_keys ? kBKV : last_keys;
  const int ntail = (valid_last + 15) & ~15;
  const int G_total = my_tiles * nkv;                  // key blocks this CTA walks through

  auto tile_coords = [&](int it, int& b, int& head, int& x) {
    const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
    x = w % p.q_tiles;
    const int bh = w / p.q_tiles;
    head = bh % p.heads;
    b = bh / p.heads;
  };
  auto issue_tile = [&](int T) {   // ring tile T: even = K of global block T/2, odd = its V
    const int G = T >> 1, it = G / nkv, j = G - it * nkv;
    int b, head, x;
    tile_coords(it, b, head, x);
    const int slot = T & 3;
    mbar_arrive_expect_tx(&bar_kv[slot], kTileBytes);
    tma_load_2d(&tmQKV, &bar_kv[slot], smem + kTileBytes * (2 + slot), ((T & 1) ? 2 * p.d : p.d) + head * kHD, b * p.N + j * kBKV);
  };
  auto issue_q = [&](int it) {
    int b, head, x;
    tile_coords(it, b, head, x);
    mbar_arrive_expect_tx(&bar_q[it & 1], kTileBytes);
    tma_load_2d(&tmQKV, &bar_q[it & 1], smem + kTileBytes * (it & 1), head * kHD, b * p.N + x * kBQ);
  };
  if (tid == kSoftmaxWarps * 32) {
    if (base & 1023u) { printf("vdr: attention smem base not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bar_q[0], 1);
    mbar_init(&bar_q[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_sfree, kSoftmaxWarps);
    mbar_init(bar_pready, kSoftmaxWarps);
    mbar_init(bar_tail, 1);
    mbar_init(bar_tailfree, kSoftmaxWarps);
    fence_barrier_init();
    if (my_tiles > 0) issue_q(0);
    for (int T = 0; T < 4 && T < 2 * G_total; ++T) issue_tile(T);
    if (my_tiles > 1) issue_q(1);
  }
  if (warp == 0) tmem_alloc<kAttnTmemCols>(tmem_ptr);
  if (p.dbg >= 100 && tid == 0) {   // experiment: the second CTA of an SM starts p.dbg ns late (the two resident CTAs would otherwise run in lockstep)
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (atomicAdd(&g_attn_sm_ticket[smid & 1023], 1u) & 1u) {
      const unsigned long long t0 = attn_gtime_always();
      while (attn_gtime_always() - t0 < static_cast<unsigned long long>(p.dbg)) __nanosleep(200);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128, tmem_P = tmem_base + 192;

  if (warp >= kSoftmaxWarps) {
    // =============================================================== issuer warpgroup
    reg_dec<32>();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    if (warp == kSoftmaxWarps) {
      // ---- K tiles, Q tiles + S = Q K^T
      auto issue_s = [&](int g, int it, int j) {
        if (j == 0) mbar_wait_relaxed(&bar_q[it & 1], (it >> 1) & 1);
        const int T = 2 * g;
        mbar_wait_relaxed(&bar_kv[T & 3], (T >> 2) & 1);
        tc_fence_after();
        { --g; V7_TRACE(15); ++g; }
        const uint64_t dq = umma_desc_kmajor_sw128(base + (it & 1) * kTileBytes);
        const uint64_t dk = umma_desc_kmajor_sw128(sRing + (T & 3) * kTileBytes);
        const uint32_t idesc = (j == nkv - 1) ? umma_idesc_bf16(128, ntail) : idesc_s;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc, k != 0);
          umma_commit(bar_s);
        }
        __syncwarp();
      };
      if (G_total > 0) issue_s(0, 0, 0);
      int it = 0, j = 0;
      for (int g = 0; g < G_total; ++g) {
        V7_TRACE(8);
        mbar_wait_relaxed(bar_sfree, g & 1);               // S_g in registers: its S columns, K_g's slot (and, at j = nkv - 1, the Q buffer) are free
        tc_fence_after();
        V7_TRACE(9);
        const bool last = (j == nkv - 1);
        if (last && it + 2 < my_tiles && elect_one()) issue_q(it + 2);
        __syncwarp();
        const int jn = last ? 0 : j + 1, itn = last ? it + 1 : it;
        if (g + 1 < G_total) issue_s(g + 1, itn, jn);
        V7_TRACE(10);
        if (g + 2 < G_total && elect_one()) issue_tile(2 * g + 4);
        __syncwarp();
        it = itn;
        j = jn;
      }
    } else if (warp == kSoftmaxWarps + 2) {
      // ---- the few trailing keys of every tile: K / V rows -> shared memory, q . k for the 128 query rows
      if (tail_keys > 0) {
        for (int it = 0; it < my_tiles; ++it) {
          int b, head, x;
          tile_coords(it, b, head, x);
          if (it > 0) mbar_wait_relaxed(bar_tailfree, (it - 1) & 1);
          const int colK = p.d + head * kHD, colV = 2 * p.d + head * kHD;
          for (int i = lane; i < tail_keys * 16; i += 32) {
            const int t = i >> 4, c = i & 15;
            const __nv_bfloat16* krow = p.qkv + static_cast<int64_t>(b * p.N + nkv * kBKV + t) * p.ld_qkv;
            s_tail[i] = __ldg(reinterpret_cast<const uint4*>(krow + (c < 8 ? colK : colV)) + (c & 7));
          }
          __syncwarp();
          mbar_wait_relaxed(&bar_q[it & 1], (it >> 1) & 1);
          const uint32_t sQ = base + (it & 1) * kTileBytes;
          for (int t = 0; t < tail_keys; ++t) {
            for (int i = 0; i < 4; ++i) {
              const int r = lane + 32 * i;
              float acc = 0.f;
#pragma unroll 2
              for (int c = 0; c < 8; ++c) {
                uint4 qu;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qu.x), "=r"(qu.y), "=r"(qu.z), "=r"(qu.w)
                             : "r"(sQ + r * 128 + ((c ^ (r & 7)) << 4)));
                const uint4 u = s_tail[t * 16 + c];
                const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
                const float2 q0v = unpack_bf16x2(qu.x), q1v = unpack_bf16x2(qu.y), q2v = unpack_bf16x2(qu.z), q3v = unpack_bf16x2(qu.w);
                acc = fmaf(q0v.x, a0.x, acc); acc = fmaf(q0v.y, a0.y, acc);
                acc = fmaf(q1v.x, a1.x, acc); acc = fmaf(q1v.y, a1.y, acc);
                acc = fmaf(q2v.x, a2.x, acc); acc = fmaf(q2v.y, a2.y, acc);
                acc = fmaf(q3v.x, a3.x, acc); acc = fmaf(q3v.y, a3.y, acc);
              }
              s_dot[t * 128 + r] = acc;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tail);
        }
      }
    } else if (warp == kSoftmaxWarps + 1) {
      // ---- V tiles + O += P V
      int j = 0;
      for (int g = 0; g < G_total; ++g) {
        V7_TRACE(11);
        mbar_wait_relaxed(bar_pready, g & 1);
        tc_fence_after();
        V7_TRACE(12);
        if (g >= 1 && g + 1 < G_total) {                   // V_{g+1} into V_{g-1}'s slot
          mbar_wait_relaxed(bar_o, (g - 1) & 1);
          if (elect_one()) issue_tile(2 * g + 3);
          __syncwarp();
        }
        const int T = 2 * g + 1;
        mbar_wait_relaxed(&bar_kv[T & 3], (T >> 2) & 1);
        tc_fence_after();
        V7_TRACE(14);
        const uint64_t dv0 = umma_desc_mnmajor_sw128(sRing + (T & 3) * kTileBytes);
        const bool full = (j < nkv - 1 || ntail == kBKV);
        if (elect_one()) {
          if (full) {
#pragma unroll
            for (int k = 0; k < kBKV / 16; ++k)
              umma_ts(tmem_O, tmem_P + k * 8, dv0 + static_cast<uint64_t>(k * 128), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          } else {
            const int ksteps = ntail / 16;
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k)
              umma_ts(tmem_O, tmem_P + k * 8, dv0 + static_cast<uint64_t>(k * 128), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(bar_o);
        }
        __syncwarp();
        V7_TRACE(13);
        j = (j == nkv - 1) ? 0 : j + 1;
      }
    }
  } else {
    // =============================================================== softmax warpgroups
    reg_inc<104>();
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_S + lane_sel + half * 64;
    const uint32_t tP = tmem_P + lane_sel + half * 32;
    const uint32_t tO = tmem_O + lane_sel + half * 32;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    int g = 0;
    for (int it = 0; it < my_tiles; ++it) {
      int b, head, x;
      tile_coords(it, b, head, x);
      const int q0 = x * kBQ;
      const int row_base = b * p.N;
      float m_ref = -INFINITY, l_run = 0.f;
      const bool idle_rows = q0 + quarter * 32 >= p.N;   // ragged last query tile: these 32 rows do not exist
      for (int j = 0; j < nkv; ++j, ++g) {
        if (warp == 1) V7_TRACE(0);
        mbar_wait(bar_s, g & 1);
        tc_fence_after();
        if (warp == 1) V7_TRACE(1);
        if (tail_keys && j == nkv - 1) mbar_wait(bar_tail, it & 1);   // orders the trailing-key warp's Q reads before the Q refill (see above)
        float alpha = 1.f;
        bool moved = false;
        float lsum;
        float* xmax = s_max + (g & 1) * 256;
        const bool narrow = (j == nkv - 1 && valid_last < kBKV);
        if (idle_rows) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sfree);
          if (j > 0) mbar_wait(bar_o, (g - 1) & 1);
        } else if (narrow) {
          const int c_lo = half * 64, c_hi = min(ntail, c_lo + 64);
          float mx = -INFINITY;
          for (int c = c_lo; c < c_hi; c += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_S + lane_sel + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c + i < valid_last) ? __uint_as_float(r[i]) : -INFINITY);
          }
          xmax[half * 128 + row] = mx;
          pair_bar_sync(quarter);
          mx = fmaxf(mx, xmax[(half ^ 1) * 128 + row]);
          const float m_new = fmaxf(m_ref, mx * p.scale_log2);
          moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);
          if (moved) {
            alpha = ex2(m_ref - m_new);
            m_ref = m_new;
          }
          lsum = 0.f;
          if (j > 0) {
            mbar_wait(bar_o, (g - 1) & 1);
            tc_fence_after();
          }
          for (int c = c_lo; c < c_hi; c += 16) {
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
          l_run = l_run * alpha + lsum;
          if (j > 0 && moved) {
            const uint64_t alpha2 = pack2(alpha, alpha);
            uint32_t r[32];
            tmem_ld_32x32b_x32(tO, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a0, a1;
              unpack2(mul2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), a0, a1);
              r[i] = __float_as_uint(a0);
              r[i + 1] = __float_as_uint(a1);
            }
            tmem_st_32x32b_x32(tO, r);
          }
        } else {
          uint32_t sr[2][32];
          tmem_ld_32x32b_x32(tS, sr[0]);
          tmem_ld_32x32b_x32(tS + 32, sr[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sfree);
          if (warp == 1) V7_TRACE(2);
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = max3(mx0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
            mx1 = max3(mx1, __uint_as_float(sr[0][i + 2]), __uint_as_float(sr[0][i + 3]));
            mx2 = max3(mx2, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
            mx3 = max3(mx3, __uint_as_float(sr[1][i + 2]), __uint_as_float(sr[1][i + 3]));
          }
          float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          xmax[half * 128 + row] = mx;
          pair_bar_sync(quarter);
          mx = fmaxf(mx, xmax[(half ^ 1) * 128 + row]);
          const float m_new = fmaxf(m_ref, mx * p.scale_log2);
          moved = __any_sync(0xffffffffu, m_new - m_ref > 8.0f);
          if (moved) {
            alpha = ex2(m_ref - m_new);
            m_ref = m_new;
          }
          const uint64_t negm2 = pack2(-m_ref, -m_ref);
          uint64_t lsum2 = 0ull;
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const uint64_t x2 = fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), scale2, negm2);
              float p0, p1;
              if ((kPolyMask >> ((i >> 1) & 7)) & 1u) {
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
          float l0, l1;
          unpack2(lsum2, l0, l1);
          l_run = l_run * alpha + (l0 + l1);
          if (warp == 1) V7_TRACE(3);
          if (j > 0) {
            mbar_wait(bar_o, (g - 1) & 1);
            tc_fence_after();
            if (moved) {
              const uint64_t alpha2 = pack2(alpha, alpha);
              uint32_t r[32];
              tmem_ld_32x32b_x32(tO, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float a0, a1;
                unpack2(mul2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), alpha2), a0, a1);
                r[i] = __float_as_uint(a0);
                r[i + 1] = __float_as_uint(a1);
              }
              tmem_st_32x32b_x32(tO, r);
            }
          }
          if (warp == 1) V7_TRACE(4);
          tmem_st_32x32b_x32(tP, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pready);
        if (warp == 1) V7_TRACE(5);
      }
      mbar_wait(bar_o, (g - 1) & 1);
      tc_fence_after();
      { --g; if (warp == 1) V7_TRACE(6); ++g; }
      // ---- epilogue of this tile (the next tile's S_0 is already on its way)
      if (!idle_rows) {
        const int q = q0 + row;
        float o[32];
        {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tO, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]);
        }
        tc_fence_before();
        if (tail_keys && p.dbg != 2) {
          for (int t = 0; t < tail_keys; ++t) {
            const float sdot = s_dot[t * 128 + row] * p.scale_log2;
            const float m_new = fmaxf(m_ref, sdot);
            const float a = ex2(m_ref - m_new);
            const float pj = __bfloat162float(__float2bfloat16_rn(ex2(sdot - m_new)));
            m_ref = m_new;
            l_run = l_run * a + (half == 0 ? pj : 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 u = s_tail[t * 16 + 8 + half * 4 + c];
              const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
              o[c * 8 + 0] = fmaf(o[c * 8 + 0], a, pj * a0.x); o[c * 8 + 1] = fmaf(o[c * 8 + 1], a, pj * a0.y);
              o[c * 8 + 2] = fmaf(o[c * 8 + 2], a, pj * a1.x); o[c * 8 + 3] = fmaf(o[c * 8 + 3], a, pj * a1.y);
              o[c * 8 + 4] = fmaf(o[c * 8 + 4], a, pj * a2.x); o[c * 8 + 5] = fmaf(o[c * 8 + 5], a, pj * a2.y);
              o[c * 8 + 6] = fmaf(o[c * 8 + 6], a, pj * a3.x); o[c * 8 + 7] = fmaf(o[c * 8 + 7], a, pj * a3.y);
            }
          }
        }
        s_sum[half * 128 + row] = l_run;
        pair_bar_sync(quarter);
        const float l_tot = l_run + s_sum[(half ^ 1) * 128 + row];
        const float inv = 1.f / l_tot;
        if (q < p.N) {
          __nv_bfloat16* op = p.out + static_cast<int64_t>(row_base + q) * p.ld_out + head * kHD + half * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = cvt_bf16x2(o[i] * inv, o[i + 1] * inv);
            w.y = cvt_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
            w.z = cvt_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
            w.w = cvt_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
            *reinterpret_cast<uint4*>(op + i) = w;
          }
          if (half == 0 && p.lse) p.lse[(static_cast<int64_t>(b) * p.heads + head) * p.N + q] = (m_ref + log2f(l_tot)) * 0.69314718055994531f;
        }
      }
      { --g; if (warp == 1) V7_TRACE(7); ++g; }
      if (tail_keys) {   // the trailing-key scratch may be rewritten for the next tile
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tailfree);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kAttnTmemCols>(tmem_base);
  }
}

// The trailing query rows of every (image, head) for the v7 grid: one CTA each (attn_tail_rows_mma)
__global__ void __launch_bounds__(kAttnThreads, 2)
flash_attn_tail_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  attn_tail_rows_mma(&tmQKV, p, smem, blockIdx.x, blockIdx.y, p.N - p.tail_rows, p.tail_rows);
}
constexpr int kTailSmem = 5 * kTileBytes + 256 + 1024;

}  // namespace vdr

static unsigned long long* g_attn_trace = nullptr;
extern "C" void vdr_debug_set_attn_trace(void* device_buf) { g_attn_trace = static_cast<unsigned long long*>(device_buf); }

static int launch_flash_attn(const char* who, const void* qkv, int64_t ld_qkv, const float* rel, int rel_pitch, void* out, int64_t ld_out,
                             float* lse, int B, int N, int heads, float scale, const vdr_dropout* drop, vdr_stream_t stream,
                             const void* rcat_hi = nullptr, const void* rcat_lo = nullptr) {
  using namespace vdr;
  VDR_CHECK_ARG(qkv && out, VDR_EINVAL, "%s: null pointer", who);
  VDR_CHECK_ARG(B > 0 && N > 0 && heads > 0, VDR_EINVAL, "%s: bad shape B=%d N=%d heads=%d", who, B, N, heads);
  const int d = heads * kHD;
  VDR_CHECK_ARG(ld_qkv >= 3 * d && ld_qkv % 8 == 0 && ld_out >= d && ld_out % 8 == 0, VDR_EALIGN, "%s: ld_qkv (%lld) / ld_out (%lld) too small or not multiples of 8", who, (long long)ld_qkv, (long long)ld_out);
  VDR_CHECK_ARG(aligned16(qkv) && aligned16(out), VDR_EALIGN, "%s: pointers must be 16-byte aligned", who);
  VDR_CHECK_ARG(B <= 65535 && heads <= 65535, VDR_EINVAL, "%s: B and heads must be <= 65535", who);
  // v6 (four light CTAs per SM) is faster than v5 when timed alone (N = 1024: 0.596 vs 0.627 ms) but slower inside the
  // power-capped extraction step (0.904 vs 0.875 ms per layer at N = 1025): v5 stays the default, VDR_ATTN_V6=1 selects v6 for A/B runs
  static const bool want_v6 = getenv("VDR_ATTN_V6") != nullptr;
  static const bool want_v5 = getenv("VDR_ATTN_V5") != nullptr;   // A/B: never pick v6 by shape
  const bool dropout = drop != nullptr && drop->thr16 != 0;
  VDR_CHECK_ARG(!dropout || (rel == nullptr && drop->thr16 < 65536u), VDR_EINVAL, "%s: attention dropout needs thr16 < 65536 and no rel-pos bias", who);
  const bool fused = rcat_hi != nullptr;
  // Short sequences with a few trailing rows (N = 257 = 2 * 128 + 1 of ViT-L/14 @224): a third of the v5 grid would be trailing-row CTAs
  // and the tile CTAs are all prologue; v6's light CTAs take every tile through the tensor cores.  Measured, B 120 / 12 heads alone:
  // N = 257 0.114 (v5) vs 0.104 ms (v6); inside the C4 step 0.201 vs 0.177 ms per launch, 6,140-6,300 vs 6,510-6,520 slices/s.
  // N = 513 is a tie, N = 1025 goes to v5 (above).
  const int tail_pre = N % kBQ;
  const bool short_tail = tail_pre >= 1 && tail_pre <= 8 && N > kBQ && N < 512 && !want_v5;
  const bool v6 = rel == nullptr && (want_v6 || short_tail) && !dropout && !fused;
  CUtensorMap tm, tm_rhi, tm_rlo;
  int rc = make_tmap_2d_bf16(&tm, qkv, (uint64_t)B * N, (uint64_t)3 * d, (uint64_t)ld_qkv, v6 ? kV6BK : 128, kHD);
  if (rc != VDR_OK) return rc;
  tm_rhi = tm;
  tm_rlo = tm;
  if (fused) {   // [rel_pos_h (2 Sh - 1) ; rel_pos_w (127)] x 64, 64-row boxes
    const uint64_t rows = 2ull * (N / 64) - 1 + 127;
    rc = make_tmap_2d_bf16(&tm_rhi, rcat_hi, rows, 64, 64, 64, 64);
    if (rc == VDR_OK) rc = make_tmap_2d_bf16(&tm_rlo, rcat_lo, rows, 64, 64, 64, 64);
    if (rc != VDR_OK) return rc;
  }
  static DeviceFlags configured;
  if (!configured.current()) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBias);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_attn_fwd_v6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV6Smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_attn_fwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemFused);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_attn_fwd_v7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV7Smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_attn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(flash_attn_fwd)");
    configured.current() = true;
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
  p.dbg = getenv("VDR_ATTN_DBG") ? atoi(getenv("VDR_ATTN_DBG")) : 0;
  p.rel = rel;
  p.rel_pitch = rel_pitch;
  p.rcat_hi = static_cast<const __nv_bfloat16*>(rcat_hi);
  p.rcat_lo = static_cast<const __nv_bfloat16*>(rcat_lo);
  p.drop = DropSpec{0ull, 0u, 0u, nullptr};
  if (dropout) p.drop = DropSpec{drop->seed, drop->site, drop->thr16, reinterpret_cast<const unsigned long long*>(drop->seed_offset)};
  // full 128-row query tiles on the tensor cores; a short tail of rows (<= 8) on one extra CUDA-core CTA per (image, head)
  // v5: full 128-row query tiles on the tensor cores; a short tail of rows (<= 8) on one extra CUDA-core CTA per (image, head).
  // v6: its CTAs are a quarter of an SM, so a trailing tile with a single valid row costs less than the CUDA-core CTA did
  // (measured alone, N = 1025: 0.723 ms with the v5 CUDA-core routine on 160 threads, 0.711 as a tile; a one-warp-per-row
  // routine with exact two-pass softmax was tried and was far slower, 0.862 ms): every tile goes through the tensor-core path.
  const int tail_rows = N % kBQ;
  // (with dropout every row and key goes through the two tensor-core softmax paths, the only ones that apply the mask)
  static const bool tail_tile = getenv("VDR_ATTN_TAIL_TILE") != nullptr;   // experiment: the trailing rows as a (mostly idle) ninth tile
  const bool vector_tail = !v6 && !dropout && !tail_tile && tail_rows > 0 && tail_rows <= 8;
  p.no_key_fold = dropout ? 1 : 0;
  p.q_tiles = vector_tail ? N / kBQ : (N + kBQ - 1) / kBQ;
  p.tail_rows = vector_tail ? tail_rows : 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // v7 (persistent tiles; experiment, VDR_ATTN_V7=1): measured slower than the one-tile-per-CTA grid (0.672 vs 0.605 ms at N = 1024)
  static const bool want_v7 = getenv("VDR_ATTN_V7") != nullptr;
  if (!fused && rel == nullptr && !dropout && !v6 && want_v7) {
    const int total = B * heads * p.q_tiles;
    if (total > 0) {
      int dev = 0, sms = 0;
      cudaError_t e = cudaGetDevice(&dev);
      if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(SM count)");
      const int ctas = total < 2 * sms ? total : 2 * sms;
      flash_attn_fwd_v7_kernel<<<ctas, kAttnThreads, kV7Smem, s>>>(tm, p);
      count_launch();
      VDR_CHECK_LAUNCH("flash_attn_fwd_v7_kernel");
    }
    if (vector_tail) {
      flash_attn_tail_kernel<<<dim3(heads, B), kAttnThreads, kTailSmem, s>>>(tm, p);
      count_launch();
      VDR_CHECK_LAUNCH("flash_attn_tail_kernel");
    }
    return VDR_OK;
  }
  dim3 grid(p.q_tiles + (vector_tail ? 1 : 0), heads, B);
  if (fused)
    flash_attn_fwd_kernel<true, false, true><<<grid, kAttnThreads, kAttnSmemFused, s>>>(tm, tm_rhi, tm_rlo, p);
  else if (rel != nullptr)
    flash_attn_fwd_kernel<true><<<grid, kAttnThreads, kAttnSmemBias, s>>>(tm, tm_rhi, tm_rlo, p);
  else if (dropout)
    flash_attn_fwd_kernel<false, true><<<grid, kAttnThreads, kAttnSmem, s>>>(tm, tm_rhi, tm_rlo, p);
  else if (v6)
    flash_attn_fwd_v6_kernel<<<grid, kV6Threads, kV6Smem, s>>>(tm, p);
  else
    flash_attn_fwd_kernel<false><<<grid, kAttnThreads, kAttnSmem, s>>>(tm, tm_rhi, tm_rlo, p);
  count_launch();
  VDR_CHECK_LAUNCH("flash_attn_fwd_kernel");
  return VDR_OK;
}

extern "C" int vdr_flash_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int B,
                                  int N, int heads, float scale, const vdr_dropout* drop, vdr_stream_t stream) {
  return launch_flash_attn("vdr_flash_attn_fwd", qkv, ld_qkv, nullptr, 0, out, ld_out, lse, B, N, heads, scale, drop, stream);
}

extern "C" int vdr_flash_attn_relpos_fwd(const void* qkv, int64_t ld_qkv, const float* rel_log2, void* out, int64_t ld_out, int B, int Sh,
                                         int heads, float scale, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(rel_log2 != nullptr && aligned16(rel_log2), VDR_EINVAL, "vdr_flash_attn_relpos_fwd: rel table must be a 16-byte aligned device pointer");
  VDR_CHECK_ARG(Sh > 0 && Sh % 4 == 0 && Sh <= 1020, VDR_EINVAL, "vdr_flash_attn_relpos_fwd: the token grid must be Sh x 64 with Sh a multiple of 4 (Sh = %d)", Sh);
  return launch_flash_attn("vdr_flash_attn_relpos_fwd", qkv, ld_qkv, rel_log2, Sh + 64, out, ld_out, nullptr, B, Sh * 64, heads, scale, nullptr, stream);
}

// The same with the bias terms computed by the kernel itself from the split tables (no table in HBM, no vdr_relpos_tables
// launch): rcat_hi / rcat_lo = [rel_pos_h (2 Sh - 1, 64) ; rel_pos_w (127, 64)] as bf16 hi + lo parts.
extern "C" int vdr_flash_attn_relpos_fused_fwd(const void* qkv, int64_t ld_qkv, const void* rcat_hi_bf16, const void* rcat_lo_bf16, void* out,
                                               int64_t ld_out, int B, int Sh, int heads, float scale, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(rcat_hi_bf16 != nullptr && rcat_lo_bf16 != nullptr && aligned16(rcat_hi_bf16) && aligned16(rcat_lo_bf16), VDR_EINVAL,
                "vdr_flash_attn_relpos_fused_fwd: the split tables must be 16-byte aligned device pointers");
  VDR_CHECK_ARG(Sh > 0 && Sh % 4 == 0 && Sh <= 64, VDR_EINVAL,
                "vdr_flash_attn_relpos_fused_fwd: the token grid must be Sh x 64 with Sh a multiple of 4, Sh <= 64 (Sh = %d)", Sh);
  return launch_flash_attn("vdr_flash_attn_relpos_fused_fwd", qkv, ld_qkv, nullptr, 0, out, ld_out, nullptr, B, Sh * 64, heads, scale, nullptr, stream,
                           rcat_hi_bf16, rcat_lo_bf16);
}
