// Shared device-side helpers for libvdr (sm_100a only): mbarrier, TMA, tcgen05/TMEM wrappers.
// Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vdr.h"

namespace vdr {

// ----------------------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int  num_sms();
// Encode a 2-D bf16 row-major tensor map (dims: inner = cols, outer = rows), 128-byte (default) or 64-byte swizzle.
int  make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols,
                       uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols, int swizzle_bytes = 128);

// Rank-N (<= 5) bf16 tensor map: dims / box innermost first, strides_bytes for dims 1..rank-1.
int  make_tmap_nd_bf16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int swizzle_bytes = 128);

#define VDR_CHECK_ARG(cond, code, ...)                 \
  do {                                                 \
    if (!(cond)) {                                     \
      ::vdr::set_error(__VA_ARGS__);                   \
      return (code);                                   \
    }                                                  \
  } while (0)

#define VDR_CHECK_LAUNCH(what)                                         \
  do {                                                                 \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) return ::vdr::cuda_fail(e__, what);        \
  } while (0)

// sam_window_tc.cu: 14 x 14 windows with decomposed rel-pos bias on tcgen05
int  launch_attn_win14_tc(const void* qkv, int64_t ld_qkv, const float* qkv_bias, const void* rcat_hi, const void* rcat_lo, void* out,
                          int64_t ld_out, int B, int gh, int gw, int heads, float scale, cudaStream_t s);

// One flag per CUDA device: cudaFuncSetAttribute & co. are per-device state, a process may drive several devices.
struct DeviceFlags {
  bool f[64] = {};
  bool& current() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    return f[d];
  }
};

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ----------------------------------------------------------------------------- device helpers
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
#ifndef VDR_WAIT_LIMIT
#define VDR_WAIT_LIMIT (1u << 27)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > VDR_WAIT_LIMIT) {
      printf("vdr: mbarrier wait timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x,
             (void*)bar, parity);
      __trap();
    }
  }
}

// ---- TMA --------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared -> global tile store (bulk async group); the box is clipped at the tensor's edges
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest kPending bulk groups of this thread have finished READING their shared-memory source
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void tma_load_2d_addr(const CUtensorMap* map, uint64_t* bar, uint32_t dst_smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 5-D tile load (the im2col view of the patch embedding: ix, iy, px, py, image)
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> f32, single CTA.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns, registers -> TMEM (thread i of the warp writes row lane base + i).
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
// 16-column variants for ragged tails
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (bf16, row = lane, two K elements per 32-bit column) comes from TMEM.
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- CTA pairs (cta_group::2): two SMs of a cluster cooperate on one 256-row MMA --------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // plain form (as CUTLASS' ClusterBarrier::arrive): an explicit .release.cluster here fences every outstanding
  // memory operation of the thread and serialised the pair pipeline to one TMA round trip per stage.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the mbarrier at cluster address `bar`
// (the leader CTA's "full" barrier, which the single MMA-issuing thread waits on).
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256: each CTA supplies its 128 rows of A and its half of B's rows.
__device__ __forceinline__ void umma_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of this thread's prior MMAs arrives on the barrier at the same offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- UMMA descriptors ----------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B),
// 8-row groups 1024 B apart (dense tile as written by a SWIZZLE_128B TMA box of 64 columns).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused here: 1)
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// K-major operand with 32-byte rows (16 bf16 = one MMA K step per row), 32-byte swizzle, 8-row groups 256 B apart:
// what a SWIZZLE_32B TMA box with a 16-element inner dimension writes (the im2col view of the patch embedding).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
// Instruction descriptor, kind::f16: bf16 A/B, f32 D.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt (1 = bf16)
//   [15] A major (0 = K)   [16] B major (0 = K, 1 = MN)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- misc -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(b);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// packed fp32x2 helpers (FFMA2 / FADD2 / FMUL2 on sm_100): two fp32 lanes per issue slot
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 + approx-unit error ~1e-7): one MUFU.RCP, one MUFU.EX2
// and a handful of FMAs, branch-free, no slow-path calls -- the GEMM epilogue applies it to every element of
// the MLP hidden layer, where erff() (two-branch polynomial) made the epilogue the bottleneck.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(ax * ax * -1.4426950408889634f);
  const float r = fmaf(-p * t, e, 1.0f);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, erf_fast(x * 0.70710678118654752f), hx);
}

// Exact-erf GELU of two values on the packed fp32x2 pipes.  erf by Abramowitz-Stegun 7.1.28,
//   erf(z) = 1 - r,  r = 1 / (1 + a1 z + ... + a6 z^6)^16  (z >= 0, |error| <= 3e-7; ~1e-6 as evaluated in fp32),
// and  gelu(x) = x/2 (1 + erf(x / sqrt 2)) = relu(x) - |x|/2 * r  (both signs; no copysign, no 1 - r):
// two FMUL, eight packed FFMA2/FMUL2 for the polynomial and its 16th power, two MUFU.RCP, two FMNMX and two more
// packed ops per PAIR of elements, branch-free, no slow-path calls.
__device__ __forceinline__ uint64_t gelu_fast2(uint64_t x2) {
  float x0, x1;
  unpack2(x2, x0, x1);
  const uint64_t z2 = pack2(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
  uint64_t acc = fma2(z2, pack2(0.0000430638f, 0.0000430638f), pack2(0.0002765672f, 0.0002765672f));
  acc = fma2(acc, z2, pack2(0.0001520143f, 0.0001520143f));
  acc = fma2(acc, z2, pack2(0.0092705272f, 0.0092705272f));
  acc = fma2(acc, z2, pack2(0.0422820123f, 0.0422820123f));
  acc = fma2(acc, z2, pack2(0.0705230784f, 0.0705230784f));
  acc = fma2(acc, z2, pack2(1.0f, 1.0f));
  acc = mul2(acc, acc);
  acc = mul2(acc, acc);
  acc = mul2(acc, acc);
  acc = mul2(acc, acc);
  float p0, p1;
  unpack2(acc, p0, p1);
  const uint64_t r2 = pack2(rcp_approx(p0), rcp_approx(p1));                              // 1 - erf(|x| / sqrt 2)
  const uint64_t relu2 = pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  const uint64_t nhax2 = mul2(z2, pack2(-0.70710678118654752f, -0.70710678118654752f));   // -|x| / 2
  return fma2(nhax2, r2, relu2);
}

// The same GELU over the 16 packed pairs of an epilogue chunk, written step by step so that the 16 dependency chains
// are interleaved.  Packed f32x2 on purpose: on B200 a packed FFMA2 / FMUL2 occupies the FMA pipe for ~4 cycles per
// scheduler against ~1.2 for a scalar FFMA (tools/microbench/pipe_rates.cu), so scalar code has more raw throughput --
// but it also executes ~1.6x the instructions, and the extraction step runs at the 1 kW power cap, where fewer
// instructions win (measured in the pipeline: 0.61 ms packed vs 0.65 ms scalar for the fc1 GEMM).
__device__ __forceinline__ void gelu_fast2_x16(uint64_t (&x2)[16]) {
  uint64_t z2[16], acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0, x1;
    unpack2(x2[i], x0, x1);
    z2[i] = pack2(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = fma2(z2[i], pack2(0.0000430638f, 0.0000430638f), pack2(0.0002765672f, 0.0002765672f));
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = fma2(acc[i], z2[i], pack2(0.0001520143f, 0.0001520143f));
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = fma2(acc[i], z2[i], pack2(0.0092705272f, 0.0092705272f));
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = fma2(acc[i], z2[i], pack2(0.0422820123f, 0.0422820123f));
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = fma2(acc[i], z2[i], pack2(0.0705230784f, 0.0705230784f));
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = fma2(acc[i], z2[i], pack2(1.0f, 1.0f));
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = mul2(acc[i], acc[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float p0, p1;
    unpack2(acc[i], p0, p1);
    acc[i] = pack2(rcp_approx(p0), rcp_approx(p1));                     // 1 - erf(|x| / sqrt 2)
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0, x1;
    unpack2(x2[i], x0, x1);
    const uint64_t relu2 = pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
    x2[i] = fma2(mul2(z2[i], pack2(-0.70710678118654752f, -0.70710678118654752f)), acc[i], relu2);   // relu(x) - |x|/2 * r
  }
}

// GELU as x * sigmoid(2 u(x)), u an odd quintic fitted (minimax over |x| <= 8) to atanh(erf(x / sqrt 2)):
//   gelu(x) = x / (1 + 2^(x * (k0 + k1 x^2 + k2 x^4))),  k = -2 log2(e) * (0.79750528, 0.03700802, -0.00035190)
// |error| <= 2.6e-5 against the erf form for every x (0.3 % of a bf16 ulp at 1.0; the classic cubic "tanh GELU" is 4.7e-4),
// correct limits (x -> +inf: x, x -> -inf: -0).  x^2 is clamped at 64 inside the polynomial, which otherwise changes sign at
// |x| = 10.9.  Per PAIR of elements: five packed FFMA2 / FMUL2 / FADD2, two FMNMX and four MUFU (EX2, RCP) against twelve packed
// operations, two FMUL, two FMNMX and two MUFU for the Abramowitz-Stegun form above.  Only used where the result is rounded to
// bf16 (oracle tolerance in tests/test_gpu_kernels.py); f32 outputs keep gelu_fast.
__device__ __forceinline__ void gelu_sig2_x16(uint64_t (&x2)[16]) {
  uint64_t q2[16], t2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a, b;
    unpack2(mul2(x2[i], x2[i]), a, b);
    q2[i] = pack2(fminf(a, 64.f), fminf(b, 64.f));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) t2[i] = fma2(q2[i], pack2(0.0010153757175430655f, 0.0010153757175430655f), pack2(-0.10678257048130035f, -0.10678257048130035f));
#pragma unroll
  for (int i = 0; i < 16; ++i) t2[i] = fma2(t2[i], q2[i], pack2(-2.3011138439178467f, -2.3011138439178467f));
#pragma unroll
  for (int i = 0; i < 16; ++i) t2[i] = mul2(t2[i], x2[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a, b;
    unpack2(t2[i], a, b);
    t2[i] = add2(pack2(ex2_approx(a), ex2_approx(b)), pack2(1.0f, 1.0f));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a, b;
    unpack2(t2[i], a, b);
    x2[i] = mul2(x2[i], pack2(rcp_approx(a), rcp_approx(b)));
  }
}

// ---- dropout (train mode of the point-cloud classifier: nn.TransformerEncoderLayer(dropout=p), MLPLayer, models_archs.py:51,58,135,187-199)
// Counter-based, so the backward kernels REGENERATE the masks instead of storing them.  Element (row r, column c) of dropout site
// `site` is kept iff the 16-bit lane (c & 7) of Philox4x32-10(counter = (c >> 3, r_lo, r_hi, site), key = seed) is >= thr16;
// kept values are multiplied by 65536 / (65536 - thr16), i.e. p = thr16 / 65536 (0.1 -> 6554, 0.5 -> 32768).  thr16 == 0 = no dropout.
struct DropSpec {
  unsigned long long seed;
  uint32_t site;
  uint32_t thr16;
  const unsigned long long* seed_offset;   // device scalar added to the seed (nullptr = 0): see vdr_dropout
};
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// the eight 16-bit lanes of columns [c8 * 8, c8 * 8 + 8) of row `row`
__device__ __forceinline__ uint4 drop_bits8(const DropSpec& d, uint64_t row, uint32_t c8) {
  const unsigned long long seed = d.seed + (d.seed_offset != nullptr ? __ldg(d.seed_offset) : 0ull);
  return philox4x32_10(make_uint4(c8, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), d.site),
                       make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
}
__device__ __forceinline__ uint32_t drop_lane16(const uint4& bits, int lane8) {
  const uint32_t w = lane8 < 2 ? bits.x : lane8 < 4 ? bits.y : lane8 < 6 ? bits.z : bits.w;
  return (lane8 & 1) ? (w >> 16) : (w & 0xffffu);
}
__device__ __forceinline__ float drop_scale(const DropSpec& d) { return 65536.f / static_cast<float>(65536u - d.thr16); }
// multiplier (0 or scale) of a single element; for vectors and other low-rate sites
__device__ __forceinline__ float drop_factor(const DropSpec& d, uint64_t row, uint32_t col) {
  if (d.thr16 == 0) return 1.f;
  const uint4 bits = drop_bits8(d, row, col >> 3);
  return drop_lane16(bits, static_cast<int>(col & 7)) >= d.thr16 ? drop_scale(d) : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__
}  // namespace vdr
