// tcgen05 GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T + bias (+ R)).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, accumulators in TMEM)
//   warps 2..9  epilogue       (tcgen05.ld -> bias / GELU / residual -> per-warp TMA stores)
// Two TMEM accumulator buffers (2*BN columns) let the epilogue of tile i overlap the MMAs of
// tile i+1.  Replaces the torch.nn.Linear call sites listed in include/vdr.h.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vdr {

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// trace record (CTA 0 only): slot = tile_iter * 16 + event  (events 8.. = epilogue warp 0, chunks 0 and 1)
//   0 mma: accumulator free   1 mma: first stage landed   2 mma: last MMA issued
//   3 epi: accumulator ready  4 epi: tile stored          5 producer: first TMA of tile issued  6 producer: last TMA issued
// (compiled in only with -DVDR_GEMM_TRACE, tools/gemm_trace.py: the checks sit in the per-chunk epilogue body)
#ifdef VDR_GEMM_TRACE
#define VDR_TRACE(ev, it)                                                                      \
  do {                                                                                         \
    if (p.trace != nullptr && blockIdx.x == 0 && (it) < 64) p.trace[(it) * 16 + (ev)] = gtime(); \
  } while (0)
#else
#define VDR_TRACE(ev, it) do { } while (0)
#endif

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32 + 32;
constexpr int kTmaWarp = 0, kMmaWarp = 1, kFirstEpiWarp = 2, kStatsWarp = kFirstEpiWarp + kEpiWarps;

struct GemmParams {
  const float* bias;
  const void* R;
  void* C;
  int64_t ldr, ldc;
  int M, N, K;
  int epilogue, r_dtype, c_dtype;
  int out_group, out_group_stride, out_offset;
  int res_mod, res_offset;
  // LayerNorm folded into this GEMM (vdr.h): consumer side (ln_stats: [ln_slots][M] (sum, sumsq) of A's rows -> per-row rstd / mean,
  // ln_colsum: column sums of the folded weight) and producer side (stats_out: [N/64][M] (sum, sumsq) of the rows written)
  const float2* ln_stats; const float* ln_colsum; float2* stats_out;
  int ln_slots; float ln_eps;
  // A operand as an im2col VIEW of (images, H, W) bf16 pictures (patch embedding, 16-pixel patches): tmA is then a 5-D tensor
  // map (ix, px, py, iy, image), SWIZZLE_32B, and a K block of a 128-patch tile lands as four [128 patches][16 ix] sub-tiles with
  // 32-byte rows, one per pixel row iy = one per MMA K step (a SWIZZLE_128B box with a 32-byte inner dimension faults on sm_100a)
  int a_im2col, ic_patch, ic_gw, ic_np, ic_chan, ic_kb_per_chan;
  int tma_out;                 // 1: bf16 output tiles leave through TMA stores (tmC); 2: ... and the bf16 residual arrives through TMA loads (tmR)
  int64_t out_rows;            // rows of the output matrix (TMA clipping bound)
  unsigned long long* trace;   // debug: per-tile timestamps of CTA 0 (nullptr = off)
  int dbg;                     // debug: 1 = skip MMAs, 2 = skip TMA loads (timing experiments only; results are garbage)
  DropSpec drop;               // generic instantiations, residual epilogue: C = R + dropout(A W^T + bias)  (thr16 == 0: off)
};

// kCtas = 2: a CTA pair (cluster of two SMs, cta_group::2) computes a 256 x BN tile: each CTA stages its own 128 rows
// of A and HALF of the BN rows of W per K block (32 KB instead of 48 KB at BN = 256), the leader CTA issues one
// M = 256 tcgen05.mma that reads both halves, and each CTA's TMEM receives its 128 accumulator rows.  This halves
// the B-operand shared-memory traffic, which bounds the single-CTA kernel (TMA writes + UMMA reads ~ 96 KB per
// 512-cycle K block against 128 B/clk of shared-memory bandwidth).
template <int BN, int kCtas, bool kRes>
struct GemmCfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = (BN / kCtas) * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  // kRes = the instantiation that can take a TMA-loaded bf16 residual (2 KB more per epilogue warp): the stage count is what
  // fits 227 KB next to the epilogue tiles -- the bias / GELU instantiation of the CTA-pair kernel keeps a sixth stage.
  static constexpr int kStages = (kCtas == 2) ? (kRes ? 5 : 6) : ((BN == 256) ? 4 : (BN == 128 ? (kRes ? 5 : 6) : (kRes ? 7 : 8)));
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  // per epilogue warp: a 2 KB store tile (+ a 2 KB TMA-loaded residual tile; the single-CTA 128 x 256 configuration has no
  // room for it next to four 48 KB stages and keeps the register-staged residual path)
  static constexpr bool kResTma = kRes && !(BN == 256 && kCtas == 1);
  static constexpr int kEpiBufBytes = kResTma ? 4096 : 2048;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiBufBytes + 1024 /*align slack*/ + 512 /*barriers*/ + 2 * BN * 4 /*bias, two tiles*/ + 2 * BN * 4 /*folded-LayerNorm column sums*/ +
                                    2 * BM * 8 /*folded-LayerNorm row coefficients*/;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// kFold >= 0 fixes the epilogue at compile time -- bit 0: ln_stats consumer, bit 1: stats_out producer, bits 2..3: epilogue
// selector -- so that the per-chunk epilogue body carries neither branches nor dead code for the other roles (3 % on the K = 768
// shapes); -1 reads everything from the parameters.
template <int BN, int kCtas, bool kRes, int kFold>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using Cfg = GemmCfg<BN, kCtas, kRes>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  // layout (base is 1024-byte aligned): stages | epilogue tiles (swizzle patterns need 512-byte alignment) | barriers | bias
  constexpr int kAuxOff = kStages * Cfg::kStageBytes + kEpiWarps * Cfg::kEpiBufBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kAuxOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;                       // [kEpiWarps] residual chunk landed (one per epilogue warp)
  uint64_t* ln_full = res_bar + kEpiWarps;                  // [2] row coefficients of a tile staged by the statistics warp
  uint64_t* ln_empty = ln_full + 2;                         // [2] ... and read by all epilogue warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ln_empty + 2);
  // per-tile epilogue constants, staged one tile ahead by the auxiliary warp (two buffers):
  float* bias_smem = reinterpret_cast<float*>(smem + kAuxOff + 512);  // [2][BN] bias of the tile's columns
  float* csum_smem = bias_smem + 2 * BN;                              // [2][BN] folded-LayerNorm column sums
  float2* ln_ab = reinterpret_cast<float2*>(csum_smem + 2 * BN);      // [2][BM] folded-LayerNorm row coefficients (rstd, -mean * rstd)
  const uint32_t stage_base = base + kStages * Cfg::kStageBytes;                          // [kEpiWarps][store tile 2 KB (| residual tile 2 KB)]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kTileM = BM * kCtas;                        // rows of one (pair) tile
  const int m_tiles = (p.M + kTileM - 1) / kTileM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;
  const uint32_t cta_rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const int tile0 = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;   // both CTAs of a pair walk the same tiles

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (p.tma_out >= 1) tma_prefetch_desc(&tmC);
    if (p.tma_out >= 2) tma_prefetch_desc(&tmR);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);                // pair: the leader's arrive.expect_tx covers the bytes of BOTH CTAs
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps * kCtas);   // pair: the epilogue warps of BOTH CTAs release the leader
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(&res_bar[w], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ln_full[a], 1);
      mbar_init(&ln_empty[a], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kFirstEpiWarp) {
    if constexpr (kCtas == 2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_ptr);
    else tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all();   // barrier inits visible to the peer before any remote arrive / TMA credit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    // (warp-wide loop, the TMA / mbarrier instructions under elect.sync: see the MMA issuer below)
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          if (elect_one()) {
          if constexpr (kCtas == 2) {
            // both CTAs load their own A rows and their half of W; all bytes are credited to the LEADER's barrier
            const uint32_t lead_bar = mapa_cluster(smem_u32(&full_bar[stage]), 0);
            // (the peer's bytes may be credited before the leader's expect_tx of that phase: the transaction count
            //  simply goes negative first; the phase cannot complete before the leader's own arrival)
            if (p.dbg == 2) {
              if (cta_rank == 0) mbar_arrive(&full_bar[stage]);
            } else {
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * 2);
              if (p.a_im2col) {
                const int m0 = m_blk * kTileM + static_cast<int>(cta_rank) * BM, img = m0 / p.ic_np, py0 = (m0 % p.ic_np) / p.ic_gw;
                const int c = kb / p.ic_kb_per_chan, iy0 = (kb % p.ic_kb_per_chan) * (BK / p.ic_patch);
                tma_load_5d_2sm(&tmA, lead_bar, sa, 0, 0, py0, iy0, img * p.ic_chan + (p.ic_chan > 1 ? c : 0));
              } else {
                tma_load_2d_2sm(&tmA, lead_bar, sa, kb * BK, m_blk * kTileM + static_cast<int>(cta_rank) * BM);
              }
              tma_load_2d_2sm(&tmW, lead_bar, sa + Cfg::kStageBytesA, kb * BK, n_blk * BN + static_cast<int>(cta_rank) * (BN / 2));
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (p.a_im2col) {
              const int m0 = m_blk * BM, img = m0 / p.ic_np, py0 = (m0 % p.ic_np) / p.ic_gw;
              const int c = kb / p.ic_kb_per_chan, iy0 = (kb % p.ic_kb_per_chan) * (BK / p.ic_patch);
              tma_load_5d(&tmA, &full_bar[stage], sa, 0, 0, py0, iy0, img * p.ic_chan + (p.ic_chan > 1 ? c : 0));
            } else {
              tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m_blk * BM);
            }
            tma_load_2d(&tmW, &full_bar[stage], sa + Cfg::kStageBytesA, kb * BK, n_blk * BN);
          }
          }
          __syncwarp();
          if (kb == 0 && lane == 0) VDR_TRACE(5, it);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (lane == 0) VDR_TRACE(6, it);
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the tiles and waits on the barriers; the tcgen05 instructions sit under elect.sync.  The compiler
    // knows that branch has exactly one active lane and feeds the instructions from uniform registers directly; under
    // `lane == 0` it wrapped EVERY tcgen05.mma / commit in an ELECT / BRA.U.ANY waterfall (~19 dependent instructions per MMA,
    // found in the SASS while profiling the attention kernel): the issue loop was then about as long as the 128 cycles the
    // pair's tensor cores need per 256 x 256 x 16 step, i.e. the tensor pipe waited for its issuer.
    if (cta_rank == 0) {   // pair: only the leader CTA issues MMAs
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        if (lane == 0) VDR_TRACE(0, it);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && lane == 0) VDR_TRACE(1, it);
          const uint32_t sa = base + stage * Cfg::kStageBytes;
          // A: +32 bytes per K step inside the 128-byte swizzle row (encoded address units of 16 B); in im2col mode the K
          // steps are the four [128][16] sub-tiles (32-byte rows, SWIZZLE_32B), 4 KB = 256 address units apart
          const uint64_t da = p.a_im2col ? umma_desc_kmajor_sw32(sa) : umma_desc_kmajor_sw128(sa);
          const uint64_t da_step = p.a_im2col ? 256u : 2u;
          const uint64_t db = umma_desc_kmajor_sw128(sa + Cfg::kStageBytesA);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (p.dbg == 1) break;
              if constexpr (kCtas == 2)
                umma_ss_2sm(d_tmem, da + da_step * k, db + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
              else
                umma_ss(d_tmem, da + da_step * k, db + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // smem slot free once these MMAs have read it (pair: in both CTAs)
            if constexpr (kCtas == 2) umma_commit_2sm(&empty_bar[stage], 3);
            else umma_commit(&empty_bar[stage]);
            // accumulator complete (pair: wakes the epilogue warps of both CTAs)
            if (kb == k_blocks - 1) {
              if constexpr (kCtas == 2) umma_commit_2sm(&tmem_full[acc], 3);
              else umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (lane == 0) VDR_TRACE(2, it);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp == kStatsWarp) {
    // ------------------------------------------------------------------ auxiliary warp: epilogue constants, one tile ahead
    // Stages what the epilogue warps need besides the accumulator -- the bias of the tile's columns and, for a folded LayerNorm,
    // the column sums and each row's (rstd, -mean * rstd) from the (sum, sumsq) slots -- in shared memory, so that no global-load
    // latency sits between two tiles of the epilogue warps.
    const float inv_k = 1.0f / static_cast<float>(p.K);
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int buf = it & 1;
      mbar_wait(&ln_empty[buf], ((it >> 1) & 1) ^ 1u);
      const int m0 = (tile / n_tiles) * kTileM + static_cast<int>(cta_rank) * BM, n0 = (tile % n_tiles) * BN;
      for (int i = lane; i < BN; i += 32) bias_smem[buf * BN + i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
      if (kFold >= 0 ? (kFold & 1) != 0 : p.ln_stats != nullptr) {
        for (int i = lane; i < BN; i += 32) csum_smem[buf * BN + i] = (n0 + i < p.N) ? __ldg(p.ln_colsum + n0 + i) : 0.f;
        float su[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
        for (int sl = 0; sl < p.ln_slots; ++sl) {
          const float2* st = p.ln_stats + static_cast<int64_t>(sl) * p.M;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = m0 + lane + 32 * j;
            const float2 t = m < p.M ? __ldg(st + m) : make_float2(0.f, 0.f);
            su[j] += t.x;
            sq[j] += t.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float mean = su[j] * inv_k;
          const float rstd = rsqrtf(fmaxf(sq[j] * inv_k - mean * mean, 0.f) + p.ln_eps);
          ln_ab[buf * BM + lane + 32 * j] = make_float2(rstd, -mean * rstd);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ln_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - kFirstEpiWarp;
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int epi_sel = kFold >= 0 ? (kFold >> 2) : p.epilogue;
    const bool ln_in = kFold >= 0 ? (kFold & 1) != 0 : p.ln_stats != nullptr;
    const bool st_out = kFold >= 0 ? (kFold & 2) != 0 : p.stats_out != nullptr;
    // (the specialised instantiations are only dispatched for bf16 outputs through TMA stores with N a multiple of BN, and --
    //  kRes -- for a bf16 residual through TMA loads: see launch_gemm)
    const bool res_bf16 = (kFold >= 0 && kRes) ? true : (epi_sel == VDR_EPI_BIAS_RESIDUAL && p.r_dtype == VDR_DTYPE_BF16);
    // Per-warp staging tile (32 rows x 64 B, 16-byte chunks XOR-swizzled): accumulators arrive with
    // thread == row, but global memory wants consecutive lanes on consecutive addresses.  Going through
    // this tile turns 32 scattered 16-byte accesses per instruction into 8 rows x 64 contiguous bytes.
    const uint32_t stg = stage_base + ew * Cfg::kEpiBufBytes, rbuf = stg + 2048;
    auto sw = [](int row, int chunk) -> uint32_t { return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); };   // = TMA SWIZZLE_64B
    const int crow = lane >> 2, cchk = lane & 3;   // coalesced layout: rows crow + 8 i, 16-byte chunk cchk
    const bool tma_store = kFold >= 0 ? true : p.tma_out >= 1;
    const bool tma_res = (kFold >= 0 && kRes) ? Cfg::kResTma : (Cfg::kResTma && res_bf16 && p.tma_out >= 2);
    uint32_t res_phase = 0;
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int m_warp = m_blk * kTileM + static_cast<int>(cta_rank) * BM + quarter * 32;
      const int m = m_warp + lane;
      const bool row_ok = m < p.M;
      auto map_rows = [&](int mm, int64_t& orow, int64_t& rrow_) {
        orow = mm;
        if (p.out_group > 0) orow = static_cast<int64_t>(mm / p.out_group) * p.out_group_stride + p.out_offset + mm % p.out_group;
        rrow_ = orow;
        if (p.res_mod > 0) rrow_ = p.res_offset + mm % p.res_mod;
      };
      int64_t out_row, res_row;
      map_rows(m, out_row, res_row);
      int64_t out_row_w, res_row_w;                 // first row of this warp's 32-row slab (TMA paths: the slab is contiguous)
      map_rows(m_warp, out_row_w, res_row_w);
      int64_t out_row_c[4], res_row_c[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) map_rows(m_warp + crow + 8 * i, out_row_c[i], res_row_c[i]);
      const int nw0 = n_blk * BN + half * kColsPerWarp;      // first column of this warp
      const bool warp_ok = m_warp < p.M;                     // the slab has at least one real row
      // The tile's epilogue constants were staged by the auxiliary warp; start the first residual chunk before waiting for the MMAs.
      const int cbuf = it & 1;
      const float* sbias = bias_smem + cbuf * BN + half * kColsPerWarp;
      const float* scsum = csum_smem + cbuf * BN + half * kColsPerWarp;
      uint64_t st_sum2 = pack2(0.f, 0.f), st_sq2 = pack2(0.f, 0.f);
      __syncwarp();
      const __nv_bfloat16* Rb = static_cast<const __nv_bfloat16*>(p.R);
      __nv_bfloat16* Cb = static_cast<__nv_bfloat16*>(p.C);
      uint4 rcur[4], rnxt[4];
      auto load_res = [&](uint4 (&dst)[4], int c) {   // coalesced layout; only full bf16 chunks (non-TMA residual path)
        const int n = nw0 + c;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = (res_bf16 && !tma_res && n + 32 <= p.N && m_warp + crow + 8 * i < p.M)
                       ? *reinterpret_cast<const uint4*>(Rb + res_row_c[i] * p.ldr + n + cchk * 8) : make_uint4(0, 0, 0, 0);
      };
      auto tma_res_load = [&](int c) {                // 32 rows x 32 columns of R -> rbuf (rows / columns past the edge arrive as zeros)
        if (elect_one()) {        // = lane 0 of the converged warp; elect.sync lets the TMA operands go straight to uniform registers
          mbar_arrive_expect_tx(&res_bar[ew], 2048);
          tma_load_2d_addr(&tmR, &res_bar[ew], rbuf, nw0 + c, static_cast<int>(res_row_w));
        }
      };
      if (tma_res) tma_res_load(0);
      else load_res(rcur, 0);
      mbar_wait(&ln_full[cbuf], (it >> 1) & 1);
      // folded LayerNorm: out = rstd * acc - (mean * rstd) * colsum + bias'
      uint64_t ln_a2 = 0, ln_b2 = 0;
      if (ln_in) {
        const float2 ab = ln_ab[cbuf * BM + quarter * 32 + lane];
        ln_a2 = pack2(ab.x, ab.x);
        ln_b2 = pack2(ab.y, ab.y);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (ew == 0 && lane == 0) VDR_TRACE(3, it);
      // Software pipeline over the 32-column chunks of this warp: the tcgen05.ld of chunk c+1 is issued as soon as
      // the registers of chunk c have been turned into packed outputs; the finished chunk leaves through one TMA
      // store per warp (asynchronous: the only wait is for the staging tile to be reusable one chunk later).
      // (Issuing the tcgen05.ld of chunk c+1 before the arithmetic of chunk c -- two register buffers -- was measured and
      //  does not help: while the MMAs of the next tile run, TMEM reads are served slowly and in order across the eight
      //  warps, so a deeper queue only lengthens each load.)
      uint32_t r[32];
      const uint32_t taddr0 = tmem_base + static_cast<uint32_t>(acc * BN + half * kColsPerWarp) + (static_cast<uint32_t>(quarter * 32) << 16);
      tmem_ld_32x32b_x32(taddr0, r);
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += 32) {
        const int col0 = half * kColsPerWarp + c;
        if (!tma_res && c + 32 < kColsPerWarp) load_res(rnxt, c + 32);   // overlap the next residual chunk with this one
        const int n0 = n_blk * BN + col0;
        const bool fast = kFold >= 0 ? true : (p.c_dtype == VDR_DTYPE_BF16 && n0 + 32 <= p.N);
        uint4 rr[4];
        if (fast && res_bf16) {
          if (tma_res) {   // residual tile landed by TMA: this thread's row, then start the next chunk into the same tile
            mbar_wait(&res_bar[ew], res_phase);
            res_phase ^= 1u;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[g].x), "=r"(rr[g].y), "=r"(rr[g].z), "=r"(rr[g].w) : "r"(rbuf + sw(lane, g)) : "memory");
            // The next TMA load overwrites the tile just read.  The loads above are only ISSUED at this point (nothing consumes
            // rr until after the accumulator wait); without the fence a load delayed behind the main loop's shared-memory traffic
            // can observe the next chunk's bytes (seen as a rare 16-byte corruption of the residual, K = 768, pair tiles).
            fence_proxy_async_smem();   // membar.cta: the reads are performed; generic -> async proxy ordering
            __syncwarp();
            if (c + 32 < kColsPerWarp) tma_res_load(c + 32);
          } else {         // residual: coalesced registers -> staging tile -> this thread's row
            if (tma_store) {
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + sw(crow + 8 * i, cchk)), "r"(rcur[i].x), "r"(rcur[i].y), "r"(rcur[i].z), "r"(rcur[i].w) : "memory");
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[g].x), "=r"(rr[g].y), "=r"(rr[g].z), "=r"(rr[g].w) : "r"(stg + sw(lane, g)) : "memory");
            __syncwarp();
          }
        } else if (tma_res) {   // chunk not on the fast path (cannot happen with tma_out == 2: kept for protocol symmetry)
          mbar_wait(&res_bar[ew], res_phase);
          res_phase ^= 1u;
          if (c + 32 < kColsPerWarp) tma_res_load(c + 32);
        }
        tmem_ld_wait();
        if (ew == 0 && lane == 0 && c <= 32) VDR_TRACE(8 + (c >> 5) * 4, it);
        if (fast) {
          uint4 o[4];
          uint64_t v2[16];   // the chunk's 32 columns as 16 packed fp32 pairs
          if (ln_in) {   // folded LayerNorm: rstd * acc + (-mean * rstd * colsum + bias')
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + c + g * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + c + g * 8 + 4);
              const float4 s0 = *reinterpret_cast<const float4*>(scsum + c + g * 8);
              const float4 s1 = *reinterpret_cast<const float4*>(scsum + c + g * 8 + 4);
              v2[g * 4 + 0] = fma2(ln_a2, pack2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1])), fma2(ln_b2, pack2(s0.x, s0.y), pack2(b0.x, b0.y)));
              v2[g * 4 + 1] = fma2(ln_a2, pack2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])), fma2(ln_b2, pack2(s0.z, s0.w), pack2(b0.z, b0.w)));
              v2[g * 4 + 2] = fma2(ln_a2, pack2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])), fma2(ln_b2, pack2(s1.x, s1.y), pack2(b1.x, b1.y)));
              v2[g * 4 + 3] = fma2(ln_a2, pack2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])), fma2(ln_b2, pack2(s1.z, s1.w), pack2(b1.z, b1.w)));
            }
          } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b0 = *reinterpret_cast<const float4*>(sbias + c + g * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sbias + c + g * 8 + 4);
            // bias add on the packed fp32x2 pipe (one issue slot per two columns)
            v2[g * 4 + 0] = add2(pack2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1])), pack2(b0.x, b0.y));
            v2[g * 4 + 1] = add2(pack2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])), pack2(b0.z, b0.w));
            v2[g * 4 + 2] = add2(pack2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])), pack2(b1.x, b1.y));
            v2[g * 4 + 3] = add2(pack2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])), pack2(b1.z, b1.w));
          }
          }
          if (epi_sel == VDR_EPI_BIAS_GELU) {
            // all 16 pairs step by step: 16 independent dependency chains in flight
#ifdef VDR_GELU_AS
            gelu_fast2_x16(v2);
#else
            gelu_sig2_x16(v2);
#endif
          } else if (epi_sel == VDR_EPI_BIAS_RESIDUAL) {
            if (kFold < 0 && p.drop.thr16 != 0) {
              // dropout1 / dropout2 of nn.TransformerEncoderLayer: the sub-layer output is dropped before the residual add.
              // This thread holds columns n0 .. n0 + 31 of output row m: four Philox calls of eight 16-bit lanes each.
              const float dsc = drop_scale(p.drop);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 bits = drop_bits8(p.drop, static_cast<uint64_t>(m), static_cast<uint32_t>((n0 >> 3) + g));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float k0 = drop_lane16(bits, 2 * k) >= p.drop.thr16 ? dsc : 0.f;
                  const float k1 = drop_lane16(bits, 2 * k + 1) >= p.drop.thr16 ? dsc : 0.f;
                  v2[g * 4 + k] = mul2(v2[g * 4 + k], pack2(k0, k1));
                }
              }
            }
            if (res_bf16) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float2 a0 = unpack_bf16x2(rr[g].x), a1 = unpack_bf16x2(rr[g].y), a2 = unpack_bf16x2(rr[g].z), a3 = unpack_bf16x2(rr[g].w);
                v2[g * 4 + 0] = add2(v2[g * 4 + 0], pack2(a0.x, a0.y));
                v2[g * 4 + 1] = add2(v2[g * 4 + 1], pack2(a1.x, a1.y));
                v2[g * 4 + 2] = add2(v2[g * 4 + 2], pack2(a2.x, a2.y));
                v2[g * 4 + 3] = add2(v2[g * 4 + 3], pack2(a3.x, a3.y));
              }
            } else if (row_ok) {   // f32 residual (position embedding: small, cache-resident): row-layout loads
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float* rp = static_cast<const float*>(p.R) + res_row * p.ldr + n0 + g * 8;
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(rp));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
                v2[g * 4 + 0] = add2(v2[g * 4 + 0], pack2(a0.x, a0.y));
                v2[g * 4 + 1] = add2(v2[g * 4 + 1], pack2(a0.z, a0.w));
                v2[g * 4 + 2] = add2(v2[g * 4 + 2], pack2(a1.x, a1.y));
                v2[g * 4 + 3] = add2(v2[g * 4 + 3], pack2(a1.z, a1.w));
              }
            }
          }
          if (st_out) {   // row statistics of what this GEMM writes (the next GEMM's folded LayerNorm): one slot per 64 columns
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              st_sum2 = add2(st_sum2, v2[i]);
              st_sq2 = fma2(v2[i], v2[i], st_sq2);
            }
            if ((c & 32) != 0) {
              float s0, s1, q0, q1;
              unpack2(st_sum2, s0, s1);
              unpack2(st_sq2, q0, q1);
              if (row_ok) p.stats_out[static_cast<int64_t>((n0 - 32) >> 6) * p.M + m] = make_float2(s0 + s1, q0 + q1);
              st_sum2 = pack2(0.f, 0.f);
              st_sq2 = pack2(0.f, 0.f);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float v[8];
            unpack2(v2[g * 4 + 0], v[0], v[1]);
            unpack2(v2[g * 4 + 1], v[2], v[3]);
            unpack2(v2[g * 4 + 2], v[4], v[5]);
            unpack2(v2[g * 4 + 3], v[6], v[7]);
            o[g] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
          if (c + 32 < kColsPerWarp) tmem_ld_32x32b_x32(taddr0 + static_cast<uint32_t>(c + 32), r);   // next chunk in flight
          if (ew == 0 && lane == 0 && c <= 32) VDR_TRACE(9 + (c >> 5) * 4, it);
          if (tma_store) {
            if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's store has read the staging tile
            __syncwarp();
            if (ew == 0 && lane == 0 && c <= 32) VDR_TRACE(10 + (c >> 5) * 4, it);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + sw(lane, g)), "r"(o[g].x), "r"(o[g].y), "r"(o[g].z), "r"(o[g].w) : "memory");
            fence_proxy_async_smem();                  // generic-proxy writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (warp_ok && elect_one()) {   // the elected lane is lane 0, the thread whose bulk groups tma_store_wait_read tracks
              tma_store_2d(&tmC, stg, n0, static_cast<int>(out_row_w));   // rows >= out_rows and columns >= N are clipped
              tma_store_commit();
            }
            if (ew == 0 && lane == 0 && c <= 32) VDR_TRACE(11 + (c >> 5) * 4, it);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + sw(lane, g)), "r"(o[g].x), "r"(o[g].y), "r"(o[g].z), "r"(o[g].w) : "memory");
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 q;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(stg + sw(crow + 8 * i, cchk)) : "memory");
              if (m_warp + crow + 8 * i < p.M) *reinterpret_cast<uint4*>(Cb + out_row_c[i] * p.ldc + n0 + cchk * 8) = q;
            }
            __syncwarp();
          }
        } else {
          if (row_ok) {   // f32 output or a ragged last column group: direct element-wise path
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = n0 + g * 8;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (n + i < p.N) {
                  float v = __uint_as_float(r[g * 8 + i]) + sbias[c + g * 8 + i];
                  if (epi_sel == VDR_EPI_BIAS_GELU) v = gelu_fast(v);
                  else if (epi_sel == VDR_EPI_BIAS_RESIDUAL)
                    v += (p.r_dtype == VDR_DTYPE_BF16) ? __bfloat162float(Rb[res_row * p.ldr + n + i])
                                                       : static_cast<const float*>(p.R)[res_row * p.ldr + n + i];
                  if (p.c_dtype == VDR_DTYPE_BF16) Cb[out_row * p.ldc + n + i] = __float2bfloat16_rn(v);
                  else static_cast<float*>(p.C)[out_row * p.ldc + n + i] = v;
                }
              }
            }
          }
          if (c + 32 < kColsPerWarp) tmem_ld_32x32b_x32(taddr0 + static_cast<uint32_t>(c + 32), r);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) rcur[i] = rnxt[i];
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the buffer back
      tc_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) VDR_TRACE(4, it);
      if (lane == 0) {
        if constexpr (kCtas == 2) mbar_arrive_cluster(mapa_cluster(smem_u32(&tmem_empty[acc]), 0));   // the leader's barrier
        else mbar_arrive(&tmem_empty[acc]);
        mbar_arrive(&ln_empty[cbuf]);   // bias / column sums / row coefficients of this tile are no longer read
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (tma_store && lane == 0) tma_store_wait_read<0>();   // the staging tile must outlive its last store
  }

  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all();   // the peer may still be credited / read by in-flight pair operations
  else __syncthreads();
  if (warp == kFirstEpiWarp) {
    tc_fence_after();
    if constexpr (kCtas == 2) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, int kCtas, bool kRes, int kFold>
static int launch_gemm_(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC, const CUtensorMap& tmR, const GemmParams& p, int grid,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN, kCtas, kRes>;
  static DeviceFlags configured;
  if (!configured.current()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, kCtas, kRes, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm)");
    configured.current() = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, kCtas, kRes, kFold>, tmA, tmW, tmC, tmR, p);
  count_launch();
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_tcgen05_kernel)");
  VDR_CHECK_LAUNCH("gemm_tcgen05_kernel");
  return VDR_OK;
}

template <int BN, int kCtas>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC, const CUtensorMap& tmR, const GemmParams& p, int grid,
                       cudaStream_t stream) {
  if constexpr (BN == 256 && kCtas == 2) {   // the hot configuration: one instantiation per epilogue / folded-LayerNorm role
    const int fold = (p.ln_stats != nullptr ? 1 : 0) | (p.stats_out != nullptr ? 2 : 0);
    const bool spec = p.tma_out >= 1 && p.c_dtype == VDR_DTYPE_BF16 && p.N % BN == 0 && p.trace == nullptr && p.drop.thr16 == 0;
    if (p.tma_out >= 2) {   // bf16 residual through TMA
      if (spec && fold == 0) return launch_gemm_<BN, kCtas, true, (VDR_EPI_BIAS_RESIDUAL << 2) | 0>(tmA, tmW, tmC, tmR, p, grid, stream);
      if (spec && fold == 2) return launch_gemm_<BN, kCtas, true, (VDR_EPI_BIAS_RESIDUAL << 2) | 2>(tmA, tmW, tmC, tmR, p, grid, stream);
      return launch_gemm_<BN, kCtas, true, -1>(tmA, tmW, tmC, tmR, p, grid, stream);
    }
    if (spec && p.epilogue == VDR_EPI_BIAS && fold == 0) return launch_gemm_<BN, kCtas, false, (VDR_EPI_BIAS << 2) | 0>(tmA, tmW, tmC, tmR, p, grid, stream);
    if (spec && p.epilogue == VDR_EPI_BIAS && fold == 1) return launch_gemm_<BN, kCtas, false, (VDR_EPI_BIAS << 2) | 1>(tmA, tmW, tmC, tmR, p, grid, stream);
    if (spec && p.epilogue == VDR_EPI_BIAS_GELU && fold == 0) return launch_gemm_<BN, kCtas, false, (VDR_EPI_BIAS_GELU << 2) | 0>(tmA, tmW, tmC, tmR, p, grid, stream);
    if (spec && p.epilogue == VDR_EPI_BIAS_GELU && fold == 1) return launch_gemm_<BN, kCtas, false, (VDR_EPI_BIAS_GELU << 2) | 1>(tmA, tmW, tmC, tmR, p, grid, stream);
    return launch_gemm_<BN, kCtas, false, -1>(tmA, tmW, tmC, tmR, p, grid, stream);
  } else {
    if (p.tma_out >= 2) return launch_gemm_<BN, kCtas, true, -1>(tmA, tmW, tmC, tmR, p, grid, stream);
    return launch_gemm_<BN, kCtas, false, -1>(tmA, tmW, tmC, tmR, p, grid, stream);
  }
}

}  // namespace vdr

static unsigned long long* g_trace = nullptr;
// debug hook (not part of the drop-in surface): device buffer of 64*8 u64 receiving CTA 0's per-tile timestamps
extern "C" void vdr_debug_set_gemm_trace(void* device_buf) { g_trace = static_cast<unsigned long long*>(device_buf); }

namespace {
struct Im2colSpec {   // A = im2col view of `images` pictures (H, W) bf16 with `chan` channel planes each (1 = gray, reused for all 3)
  const void* src;
  int images, chan, H, W, patch;
};
}  // namespace

static int gemm_impl(const vdr_gemm_args* a, const Im2colSpec* ic, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(a != nullptr, VDR_EINVAL, "vdr_gemm: null args");
  VDR_CHECK_ARG((a->A || ic) && a->W && a->C, VDR_EINVAL, "vdr_gemm: null A/W/C");
  VDR_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, VDR_EINVAL, "vdr_gemm: non-positive shape M=%d N=%d K=%d", a->M, a->N, a->K);
  VDR_CHECK_ARG((ic || a->lda >= a->K) && a->ldw >= a->K && a->ldc >= a->N, VDR_EINVAL, "vdr_gemm: leading dimension smaller than row length");
  VDR_CHECK_ARG((ic || a->lda % 8 == 0) && a->ldw % 8 == 0 && a->ldc % 8 == 0, VDR_EALIGN, "vdr_gemm: lda/ldw/ldc must be multiples of 8 elements");
  VDR_CHECK_ARG((ic || aligned16(a->A)) && aligned16(a->W) && aligned16(a->C), VDR_EALIGN, "vdr_gemm: A/W/C must be 16-byte aligned");
  VDR_CHECK_ARG(a->bias == nullptr || aligned16(a->bias), VDR_EALIGN, "vdr_gemm: bias must be 16-byte aligned");
  // N itself may be ragged (last column group handled element-wise); rows must still start 16-byte aligned.
  VDR_CHECK_ARG(a->epilogue >= VDR_EPI_BIAS && a->epilogue <= VDR_EPI_BIAS_RESIDUAL, VDR_EINVAL, "vdr_gemm: unknown epilogue %d", a->epilogue);
  VDR_CHECK_ARG(a->c_dtype == VDR_DTYPE_BF16 || a->c_dtype == VDR_DTYPE_F32, VDR_EINVAL, "vdr_gemm: bad c_dtype %d", a->c_dtype);
  if (a->epilogue == VDR_EPI_BIAS_RESIDUAL) {
    VDR_CHECK_ARG(a->R != nullptr, VDR_EINVAL, "vdr_gemm: residual epilogue needs R");
    VDR_CHECK_ARG(a->r_dtype == VDR_DTYPE_BF16 || a->r_dtype == VDR_DTYPE_F32, VDR_EINVAL, "vdr_gemm: bad r_dtype %d", a->r_dtype);
    VDR_CHECK_ARG(aligned16(a->R) && a->ldr % 8 == 0 && a->ldr >= a->N, VDR_EALIGN, "vdr_gemm: R must be 16-byte aligned with ldr %% 8 == 0");
  }
  VDR_CHECK_ARG(a->out_group >= 0 && a->res_mod >= 0, VDR_EINVAL, "vdr_gemm: negative row-remap parameters");
  if (a->ln_stats != nullptr) {
    VDR_CHECK_ARG(a->ln_colsum != nullptr && a->ln_slots > 0 && a->ln_eps > 0.f, VDR_EINVAL, "vdr_gemm: ln_stats needs ln_colsum, ln_slots > 0 and ln_eps > 0");
    VDR_CHECK_ARG(a->c_dtype == VDR_DTYPE_BF16 && a->N % 32 == 0 && !ic, VDR_EINVAL, "vdr_gemm: folded LayerNorm needs a bf16 C and N %% 32 == 0 (N = %d)", a->N);
    VDR_CHECK_ARG((reinterpret_cast<uintptr_t>(a->ln_stats) & 7) == 0 && aligned16(a->ln_colsum), VDR_EALIGN, "vdr_gemm: ln_stats must be 8-byte and ln_colsum 16-byte aligned");
  }
  if (a->stats_out != nullptr) {
    VDR_CHECK_ARG(a->epilogue == VDR_EPI_BIAS_RESIDUAL && a->c_dtype == VDR_DTYPE_BF16 && a->N % 64 == 0 && a->out_group == 0, VDR_EINVAL,
                  "vdr_gemm: stats_out needs the residual epilogue, a bf16 C, N %% 64 == 0 (N = %d) and no row remapping", a->N);
    VDR_CHECK_ARG((reinterpret_cast<uintptr_t>(a->stats_out) & 7) == 0, VDR_EALIGN, "vdr_gemm: stats_out must be 8-byte aligned");
  }

  if (a->drop.thr16 != 0) {
    VDR_CHECK_ARG(a->drop.thr16 < 65536u && a->epilogue == VDR_EPI_BIAS_RESIDUAL && a->c_dtype == VDR_DTYPE_BF16 && a->N % 32 == 0 && a->out_group == 0 &&
                  a->ln_stats == nullptr && a->stats_out == nullptr && !ic, VDR_EINVAL,
                  "vdr_gemm: dropout needs thr16 < 65536, the residual epilogue, a bf16 C, N %% 32 == 0 (N = %d), no row remapping and no folded LayerNorm", a->N);
  }
  const int sms = num_sms();
  const int m_tiles = (a->M + BM - 1) / BM;
  auto tiles = [&](int bn) { return m_tiles * ((a->N + bn - 1) / bn); };
  int bn = 256;
  if (a->N % 256 != 0 && a->N % 128 == 0) bn = 128;
  if (bn == 256 && tiles(256) < sms) bn = 128;
  if (bn == 128 && tiles(128) < sms && a->stats_out == nullptr) bn = 64;   // a statistics slot is 64 columns = one warp's slice at BN = 128
  // CTA pairs (256 x 256 tiles) when there is enough work to fill the 74 pairs
  static const bool force_1cta = getenv("VDR_GEMM_1CTA") != nullptr;
  const int pair_tiles = ((a->M + 2 * BM - 1) / (2 * BM)) * ((a->N + 255) / 256);
  const bool pair = !force_1cta && bn == 256 && pair_tiles >= sms / 2;

  CUtensorMap tmA, tmW;
  int rc;
  if (ic) {
    const int gw = ic->W / ic->patch, gh = ic->H / ic->patch;
    // (ix, px, py, iy, image): the box lands as [iy][py][px][ix] = one [128 patches][16 ix] sub-tile per pixel row
    const uint64_t dims[5] = {(uint64_t)ic->patch, (uint64_t)gw, (uint64_t)gh, (uint64_t)ic->patch, (uint64_t)ic->images * ic->chan};
    const uint64_t strides[4] = {(uint64_t)ic->patch * 2, (uint64_t)ic->patch * ic->W * 2, (uint64_t)ic->W * 2, (uint64_t)ic->H * ic->W * 2};
    const uint32_t box[5] = {(uint32_t)ic->patch, (uint32_t)gw, (uint32_t)(BM / gw), (uint32_t)(BK / ic->patch), 1};
    rc = make_tmap_nd_bf16(&tmA, ic->src, 5, dims, strides, box, 32);
  } else {
    rc = make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, BM, BK);
  }
  if (rc != VDR_OK) return rc;
  rc = make_tmap_2d_bf16(&tmW, a->W, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldw, (uint32_t)(pair ? bn / 2 : bn), BK);
  if (rc != VDR_OK) return rc;

  GemmParams p;
  p.bias = a->bias; p.R = a->R; p.C = a->C; p.ldr = a->ldr; p.ldc = a->ldc;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.epilogue = a->epilogue; p.r_dtype = a->r_dtype; p.c_dtype = a->c_dtype;
  p.out_group = a->out_group; p.out_group_stride = a->out_group_stride; p.out_offset = a->out_offset;
  p.res_mod = a->res_mod; p.res_offset = a->res_offset;
  p.ln_stats = reinterpret_cast<const float2*>(a->ln_stats); p.ln_colsum = a->ln_colsum; p.ln_slots = a->ln_slots; p.ln_eps = a->ln_eps;
  p.stats_out = reinterpret_cast<float2*>(a->stats_out);
  p.trace = g_trace;
  p.dbg = getenv("VDR_GEMM_DBG") ? atoi(getenv("VDR_GEMM_DBG")) : 0;
  p.drop = DropSpec{a->drop.seed, a->drop.site, a->drop.thr16, reinterpret_cast<const unsigned long long*>(a->drop.seed_offset)};
  p.a_im2col = ic ? 1 : 0;
  p.ic_patch = p.ic_gw = p.ic_np = p.ic_chan = p.ic_kb_per_chan = 1;
  if (ic) {
    p.ic_patch = ic->patch;
    p.ic_gw = ic->W / ic->patch;
    p.ic_np = (ic->H / ic->patch) * p.ic_gw;
    p.ic_chan = ic->chan;
    p.ic_kb_per_chan = ic->patch * ic->patch / BK;
  }

  // bf16 outputs leave through per-warp TMA stores of 32 x 32 tiles (SWIZZLE_64B staging) when every warp's 32-row slab
  // maps to 32 consecutive output rows; a bf16 residual without row remapping then arrives the same way.
  static const bool no_tma_out = getenv("VDR_GEMM_NO_TMA_OUT") != nullptr;
  CUtensorMap tmC, tmR;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmR, 0, sizeof(tmR));
  p.tma_out = 0;
  p.out_rows = a->M;
  if (a->out_group > 0) p.out_rows = (int64_t)((a->M + a->out_group - 1) / a->out_group - 1) * a->out_group_stride + a->out_offset + a->out_group;
  const bool slab_ok = a->out_group == 0 || (a->out_group % 32 == 0 && a->M % 32 == 0);
  if (!no_tma_out && a->c_dtype == VDR_DTYPE_BF16 && slab_ok && p.out_rows < 0x7fffffffLL) {
    rc = make_tmap_2d_bf16(&tmC, a->C, (uint64_t)p.out_rows, (uint64_t)a->N, (uint64_t)a->ldc, 32, 32, 64);
    if (rc != VDR_OK) return rc;
    p.tma_out = 1;
    if (a->epilogue == VDR_EPI_BIAS_RESIDUAL && a->r_dtype == VDR_DTYPE_BF16 && a->res_mod == 0 && (pair || bn != 256)) {
      rc = make_tmap_2d_bf16(&tmR, a->R, (uint64_t)p.out_rows, (uint64_t)a->N, (uint64_t)a->ldr, 32, 32, 64);
      if (rc != VDR_OK) return rc;
      p.tma_out = 2;
    }
  }

  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (pair) {
    const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
    return launch_gemm<256, 2>(tmA, tmW, tmC, tmR, p, 2 * pairs, s);
  }
  const int total = tiles(bn);
  const int grid = total < sms ? total : sms;
  if (bn == 256) return launch_gemm<256, 1>(tmA, tmW, tmC, tmR, p, grid, s);
  if (bn == 128) return launch_gemm<128, 1>(tmA, tmW, tmC, tmR, p, grid, s);
  return launch_gemm<64, 1>(tmA, tmW, tmC, tmR, p, grid, s);
}

extern "C" int vdr_gemm(const vdr_gemm_args* a, vdr_stream_t stream) { return gemm_impl(a, nullptr, stream); }

extern "C" int vdr_patch_embed_supported(int H, int W, int patch) {
  if (patch <= 0 || H <= 0 || W <= 0 || H % patch || W % patch) return 0;
  const int gw = W / patch, np = (H / patch) * gw;
  // one pixel row of a patch = one 16-wide MMA K step (32-byte TMA rows), 128-row tiles made of whole patch rows of one image
  return (patch == 16 && W % 8 == 0 && gw <= 128 && 128 % gw == 0 && np % 128 == 0) ? 1 : 0;
}

static int patch_embed_impl(const char* who, const void* images_bf16, int B, int C, int wchan, int lead, int H, int W, int patch, const void* Wpe_bf16,
                            int64_t ldw, const float* bias, const float* pos, void* X_bf16, int64_t ldx, int d, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(images_bf16 && Wpe_bf16 && pos && X_bf16, VDR_EINVAL, "%s: null pointer", who);
  VDR_CHECK_ARG(B > 0 && (C == 1 || C == 3) && d > 0 && lead >= 0, VDR_EINVAL, "%s: bad shape B=%d C=%d d=%d token_offset=%d", who, B, C, d, lead);
  VDR_CHECK_ARG(vdr_patch_embed_supported(H, W, patch), VDR_EINVAL,
                "%s: %dx%d images with %d-pixel patches do not tile into TMA im2col boxes (use vdr_im2col_* + vdr_gemm)", who, H, W, patch);
  VDR_CHECK_ARG(aligned16(images_bf16), VDR_EALIGN, "%s: images must be 16-byte aligned", who);
  const int np = (H / patch) * (W / patch);
  VDR_CHECK_ARG((int64_t)B * np < 0x7fffffffLL, VDR_EINVAL, "%s: too many patches", who);
  vdr_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.A = nullptr; a.lda = 0;
  a.W = Wpe_bf16; a.ldw = ldw;
  a.bias = bias;
  a.R = pos; a.ldr = d; a.r_dtype = VDR_DTYPE_F32;
  a.C = X_bf16; a.ldc = ldx; a.c_dtype = VDR_DTYPE_BF16;
  a.M = B * np; a.N = d; a.K = wchan * patch * patch;                 // wchan = channel planes of the WEIGHTS (1: channel-summed)
  a.epilogue = VDR_EPI_BIAS_RESIDUAL;
  if (lead > 0) { a.out_group = np; a.out_group_stride = np + lead; a.out_offset = lead; }   // patch tokens behind each image's CLS row
  a.res_mod = np; a.res_offset = lead;                                // + pos_embed[lead + patch index]
  const Im2colSpec ic{images_bf16, B, C, H, W, patch};
  return gemm_impl(&a, &ic, stream);
}

extern "C" int vdr_patch_embed_gemm(const void* images_bf16, int B, int C, int H, int W, int patch, const void* Wpe_bf16,
                                    int64_t ldw, const float* bias, const float* pos, void* X_bf16, int64_t ldx, int d,
                                    vdr_stream_t stream) {
  return patch_embed_impl("vdr_patch_embed_gemm", images_bf16, B, C, 3, 1, H, W, patch, Wpe_bf16, ldw, bias, pos, X_bf16, ldx, d, stream);
}

// Gray pictures against CHANNEL-SUMMED weights: gray2rgb feeds the same picture to the three input channels, so
// sum_c patch . W_c = patch . (W_r + W_g + W_b) -- K = p*p instead of 3*p*p, the picture is read once instead of three times.
// token_offset: rows in front of every image's patch tokens in X and pos (1 = the CLS row of a ViT, 0 = SAM's encoder).
extern "C" int vdr_patch_embed_gemm_gray(const void* images_bf16, int B, int H, int W, int patch, const void* Wsum_bf16,
                                         int64_t ldw, const float* bias, const float* pos, void* X_bf16, int64_t ldx, int d,
                                         int token_offset, vdr_stream_t stream) {
  return patch_embed_impl("vdr_patch_embed_gemm_gray", images_bf16, B, 1, 1, token_offset, H, W, patch, Wsum_bf16, ldw, bias, pos, X_bf16, ldx, d, stream);
}
