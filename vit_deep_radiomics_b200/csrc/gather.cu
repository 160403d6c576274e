// Tumour-mask gathers.
//   G1  vdr_mask_gather : stable stream compaction of in-mask ViT tokens into a point cloud
//       (warp ballot + popc ranks inside a tile, tile counts -> exclusive scan -> scatter),
//       rows copied warp-wide with 16-byte accesses, optional fp64 3-D sinusoidal PE epilogue.
//   G2  vdr_voxel_bbox / vdr_voxel_gather : bounding box of the mask, then every voxel inside it.
// Index contracts are in SURVEY.md Appendix A1/A2/A6 and restated in oracle/gather_np.py.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vdr {

constexpr int kTile = 2048;      // candidates per tile
constexpr int kGThreads = 256;   // 8 warps
constexpr int kIters = kTile / kGThreads;
constexpr int kKeep = 16;        // tiles per block whose ballots stay in shared memory between the count and the rank phase

struct G1Geom {
  const uint8_t* mask;
  const int32_t* row_map;
  const int32_t* col_map;
  int S, h, w;
  int64_t mask_slice_stride, mask_row_stride, mask_col_stride;   // bytes between slices / rows / columns of the pixel mask
  int64_t feat_slice_rows, feat_row_pitch, feat_row0;  // token row = k*slice_rows + row0 + a*pitch + b
  uint32_t total;                                       // S*h*w < 2^31: all candidate arithmetic is 32-bit
};

// candidate n = a*(w*S) + b*S + k  (slice fastest)  ->  resized-mask value
__device__ __forceinline__ bool g1_pred(const G1Geom& g, uint32_t n) {
  const uint32_t q = n / static_cast<uint32_t>(g.S), k = n - q * static_cast<uint32_t>(g.S);
  const uint32_t a = q / static_cast<uint32_t>(g.w), b = q - a * static_cast<uint32_t>(g.w);
  const int64_t off = static_cast<int64_t>(k) * g.mask_slice_stride + __ldg(g.row_map + a) * g.mask_row_stride + __ldg(g.col_map + b) * g.mask_col_stride;
  return __ldg(g.mask + off) != 0;
}

struct G1Pe {
  double scale;          // 0 = no positional encoding
  const double* div;     // [D/6] divisors 10000^(6 i / D), host-computed in f64
  double* table;         // [(h + w + S)][2*(D/6)] f64: scale * sin|cos(coord / div) per distinct coordinate (workspace)
  double w_orig, h_orig, res0, res1, res2, noise0, noise1, noise2, mean_x, mean_y, mean_z;
};

// The encoding of a token depends on its three grid indices separately (xi in [0,h), yi in [0,w), zi in [0,S) under
// the reference's meshgrid quirk), so the h + w + S distinct coordinate rows are evaluated once in f64 (same operation
// order as train_models.py:166-176 and :34-44) and the emit phase only adds table entries: (h+w+S) * D/3 sin/cos
// instead of n_sel * D.
__device__ __forceinline__ double g1_pe_entry(const G1Pe& pe, int S, int h, int w, int npair2, int64_t idx) {
  const int e = static_cast<int>(idx / npair2), jj = static_cast<int>(idx % npair2);
  double v;
  if (e < h) v = __dadd_rn(__dsub_rn(__dmul_rn(__dmul_rn(static_cast<double>(e) / static_cast<double>(w), pe.w_orig), pe.res0), pe.mean_x), pe.noise0);
  else if (e < h + w) v = __dadd_rn(__dsub_rn(__dmul_rn(__dmul_rn(static_cast<double>(e - h) / static_cast<double>(h), pe.h_orig), pe.res1), pe.mean_y), pe.noise1);
  else v = __dadd_rn(__dsub_rn(__dmul_rn(static_cast<double>(e - h - w), pe.res2), pe.mean_z), pe.noise2);
  const double arg = v / pe.div[jj >> 1];
  const double enc = (jj & 1) ? cos(arg) : sin(arg);
  return __dmul_rn(enc, pe.scale);
}

struct G1Args {
  G1Geom g;
  const void* feat;
  int64_t ld_feat;
  int D;
  float* out_tok;              // table base: row r of this call lands at table row (*row_offset + r)
  int32_t* out_src;
  int32_t* out_count;          // this call's count
  int64_t cap;                 // table capacity in rows
  const int64_t* row_offset;   // device scalar (nullable = 0): first table row of this call -- the rank's / patient's slot of a shared table
  int src_cols;                // 3: (slice,row,col);  4: (patient,slice,row,col)
  int32_t patient;
  G1Pe pe;
  int vec_ok, npair2;
  int32_t* block_totals;       // [gridDim.x] workspace
  int tiles, tiles_per_block;
  unsigned long long* trace;   // diagnostics (vdr_debug_set_gather_trace): %globaltimer of block 0 / the last block at the phase boundaries
};

__device__ __forceinline__ void g1_stamp(const G1Args& A, int slot) {
  if (A.trace != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    A.trace[(blockIdx.x == 0 ? 0 : 8) + slot] = t;
  }
}

// ONE cooperative launch: [PE table | predicate ballots + per-block totals] -> grid sync -> [ranks -> (slice,row,col) rows at the
// block's base] -> grid sync -> [emit: one warp per OUTPUT row over the whole grid].  The mask is read once (ballots of up to kKeep
// tiles per block stay in shared memory), the scan is a sum over <= gridDim.x block totals, and there is no launch gap between the phases.
template <bool FEAT_BF16>
__global__ void __launch_bounds__(kGThreads, 4) g1_fused_kernel(const G1Args A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t s_ballot[kKeep][kIters * 8];
  __shared__ int s_cnt[kIters * 8];
  __shared__ int s_off[kIters * 8 + 1];
  __shared__ uint16_t s_sel[kTile];
  __shared__ int s_red[2][kGThreads / 32];
  const G1Geom& g = A.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  g1_stamp(A, 0);
  // ---- phase 0: positional-encoding table (grid-strided; a few thousand f64 sin/cos)
  if (A.pe.scale != 0. && A.npair2 > 0) {
    const int64_t entries = static_cast<int64_t>(g.h + g.w + g.S) * A.npair2;
    for (int64_t idx = blockIdx.x * (int64_t)kGThreads + tid; idx < entries; idx += (int64_t)gridDim.x * kGThreads)
      A.pe.table[idx] = g1_pe_entry(A.pe, g.S, g.h, g.w, A.npair2, idx);
  }

  g1_stamp(A, 1);
  // ---- phase 1: predicate -> ballots, block total
  const int t0 = blockIdx.x * A.tiles_per_block;
  const int t1 = min(t0 + A.tiles_per_block, A.tiles);
  int warp_cnt = 0;
  for (int t = t0; t < t1; ++t) {
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const uint32_t n = static_cast<uint32_t>(t) * kTile + it * kGThreads + tid;
      const bool p = (n < g.total) && g1_pred(g, n);
      const uint32_t bal = __ballot_sync(0xffffffffu, p);
      warp_cnt += __popc(bal);
      if (t - t0 < kKeep && lane == 0) s_ballot[t - t0][it * 8 + warp] = bal;
    }
  }
  if (lane == 0) s_red[0][warp] = warp_cnt;
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
#pragma unroll
    for (int i = 0; i < kGThreads / 32; ++i) tot += s_red[0][i];
    A.block_totals[blockIdx.x] = tot;
  }
  g1_stamp(A, 2);
  grid.sync();
  g1_stamp(A, 3);

  // ---- phase 2: exclusive base of this block = sum of the totals of the blocks before it; ranks; (slice,row,col) rows
  int lt = 0, all = 0;
  for (int j = tid; j < static_cast<int>(gridDim.x); j += kGThreads) {
    const int v = A.block_totals[j];
    all += v;
    if (j < static_cast<int>(blockIdx.x)) lt += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
    all += __shfl_xor_sync(0xffffffffu, all, o);
  }
  __syncthreads();   // s_red[0] has been consumed
  if (lane == 0) { s_red[0][warp] = lt; s_red[1][warp] = all; }
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int i = 0; i < kGThreads / 32; ++i) { base += s_red[0][i]; total += s_red[1][i]; }
  if (blockIdx.x == 0 && tid == 0) *A.out_count = total;
  const int64_t row_off = A.row_offset ? *A.row_offset : 0;
  const int sc = A.src_cols;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int t = t0; t < t1; ++t) {
    uint32_t ballots[kIters];
    uint32_t mine = 0;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      uint32_t bal;
      if (t - t0 < kKeep) bal = s_ballot[t - t0][it * 8 + warp];
      else {
        const uint32_t n = static_cast<uint32_t>(t) * kTile + it * kGThreads + tid;
        bal = __ballot_sync(0xffffffffu, (n < g.total) && g1_pred(g, n));
      }
      ballots[it] = bal;
      mine |= ((bal >> lane) & 1u) << it;
      if (lane == 0) s_cnt[it * 8 + warp] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 64 (iteration, warp) counts, two per lane
      const int v0 = s_cnt[lane * 2], v1 = s_cnt[lane * 2 + 1];
      const int sum = v0 + v1;
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      s_off[lane * 2] = incl - sum;
      s_off[lane * 2 + 1] = incl - sum + v0;
      if (lane == 31) s_off[kIters * 8] = incl;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kIters; ++it)
      if ((mine >> it) & 1u) s_sel[s_off[it * 8 + warp] + __popc(ballots[it] & lt_mask)] = static_cast<uint16_t>(it * kGThreads + tid);
    __syncthreads();
    const int tile_n = s_off[kIters * 8];
    for (int j = tid; j < tile_n; j += kGThreads) {
      const int64_t row = row_off + base + j;
      if (row >= A.cap) break;
      const uint32_t n = static_cast<uint32_t>(t) * kTile + s_sel[j];
      const uint32_t q = n / static_cast<uint32_t>(g.S), k = n - q * static_cast<uint32_t>(g.S);
      const uint32_t a = q / static_cast<uint32_t>(g.w), b = q - a * static_cast<uint32_t>(g.w);
      int32_t* o = A.out_src + row * sc;
      if (sc == 4) *o++ = A.patient;
      o[0] = static_cast<int32_t>(k);
      o[1] = static_cast<int32_t>(a);
      o[2] = static_cast<int32_t>(b);
    }
    base += tile_n;
    __syncthreads();   // s_cnt / s_off / s_sel are rewritten by the next tile
  }
  g1_stamp(A, 4);
  grid.sync();
  g1_stamp(A, 5);

  // ---- phase 3: emit.  One warp per output row over the whole grid: copies the descriptor row and adds the positional
  // encoding from the per-coordinate f64 table (row in, row out, three cache-resident table rows).
  int64_t count = total;
  if (row_off + count > A.cap) count = A.cap > row_off ? A.cap - row_off : 0;
  const int D = A.D, third = D / 3, two_third = (2 * D) / 3, npair2 = A.npair2;
  const int64_t ld_feat = A.ld_feat;
  const double* table = A.pe.table;     // written in phase 0 of this launch: plain (coherent) loads, not the read-only path
  const bool pe_on = A.pe.scale != 0.;
  for (int64_t r = blockIdx.x * (int64_t)(kGThreads / 32) + warp; r < count; r += (int64_t)gridDim.x * (kGThreads / 32)) {
    const int32_t* sp = A.out_src + (row_off + r) * sc + (sc - 3);
    const int k = sp[0], a = sp[1], b = sp[2];
    // reference meshgrid(indexing='xy') quirk: with q = n / S = a*w + b:  xi = q % h, yi = q / h, zi = k
    const uint32_t q = static_cast<uint32_t>(a) * static_cast<uint32_t>(g.w) + static_cast<uint32_t>(b);
    const uint32_t yi = q / static_cast<uint32_t>(g.h), xi = q - yi * static_cast<uint32_t>(g.h);
    const double* tx = table + static_cast<int64_t>(xi) * npair2;
    const double* ty = table + static_cast<int64_t>(g.h + yi) * npair2;
    const double* tz = table + static_cast<int64_t>(g.h + g.w + k) * npair2;
    const int64_t src_row = static_cast<int64_t>(k) * g.feat_slice_rows + g.feat_row0 + a * g.feat_row_pitch + b;
    float* orow = A.out_tok + (row_off + r) * D;
    auto pe_add = [&](float f, int col) -> float {
      int jj;
      const double* t;
      if (col < third) { jj = col; t = tx; }
      else if (col < two_third) { jj = col - third; t = ty; }
      else { jj = col - two_third; t = tz; }
      if (jj < npair2) return static_cast<float>(__dadd_rn(static_cast<double>(f), t[jj]));
      return f;
    };
    if (A.vec_ok) {
      // D % 24 == 0 (every 8-column chunk lies inside one axis block and all of its columns are encoded): the chunk's
      // eight table entries are four 16-byte loads issued together with the descriptor loads -- one memory round trip.
      const bool pe_vec = pe_on && D % 24 == 0;
      for (int c0 = lane * 8; c0 < D; c0 += 256) {
        float f[8];
        double2 t4[4];
        if (pe_vec) {
          const double* t = (c0 < third) ? tx + c0 : (c0 < two_third) ? ty + (c0 - third) : tz + (c0 - two_third);
#pragma unroll
          for (int i = 0; i < 4; ++i) t4[i] = reinterpret_cast<const double2*>(t)[i];
        }
        if (FEAT_BF16) {
          const uint4 v = __ldcs(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(A.feat) + src_row * ld_feat + c0));
          const float2 p0 = unpack_bf16x2(v.x), p1 = unpack_bf16x2(v.y), p2 = unpack_bf16x2(v.z), p3 = unpack_bf16x2(v.w);
          f[0] = p0.x; f[1] = p0.y; f[2] = p1.x; f[3] = p1.y; f[4] = p2.x; f[5] = p2.y; f[6] = p3.x; f[7] = p3.y;
        } else {
          const float* fp = static_cast<const float*>(A.feat) + src_row * ld_feat + c0;
          const float4 u0 = __ldcs(reinterpret_cast<const float4*>(fp)), u1 = __ldcs(reinterpret_cast<const float4*>(fp + 4));
          f[0] = u0.x; f[1] = u0.y; f[2] = u0.z; f[3] = u0.w; f[4] = u1.x; f[5] = u1.y; f[6] = u1.z; f[7] = u1.w;
        }
        if (pe_vec) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[2 * i] = static_cast<float>(__dadd_rn(static_cast<double>(f[2 * i]), t4[i].x));
            f[2 * i + 1] = static_cast<float>(__dadd_rn(static_cast<double>(f[2 * i + 1]), t4[i].y));
          }
        } else if (pe_on) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = pe_add(f[i], c0 + i);
        }
        __stcs(reinterpret_cast<float4*>(orow + c0), make_float4(f[0], f[1], f[2], f[3]));
        __stcs(reinterpret_cast<float4*>(orow + c0 + 4), make_float4(f[4], f[5], f[6], f[7]));
      }
    } else {  // any D / alignment: one element per lane
      for (int c = lane; c < D; c += 32) {
        float f = FEAT_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(A.feat)[src_row * ld_feat + c])
                            : static_cast<const float*>(A.feat)[src_row * ld_feat + c];
        if (pe_on) f = pe_add(f, c);
        orow[c] = f;
      }
    }
  }
  g1_stamp(A, 6);
}

// Count only (the multi-GPU table needs every patient's row count before any rank emits: SURVEY.md 8e): tile ballots, one
// 64-bit atomic per block into a zeroed device counter.
__global__ void __launch_bounds__(kGThreads) g1_count_kernel(G1Geom g, unsigned long long* __restrict__ out) {
  __shared__ int s_warp[kGThreads / 32];
  int cnt = 0;
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const uint32_t n = blockIdx.x * static_cast<uint32_t>(kTile) + it * kGThreads + threadIdx.x;
    const bool p = (n < g.total) && g1_pred(g, n);
    cnt += __popc(__ballot_sync(0xffffffffu, p));
  }
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;  // every lane holds the warp total
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kGThreads / 32; ++i) t += s_warp[i];
    if (t) atomicAdd(out, static_cast<unsigned long long>(t));
  }
}

// offsets[i] = base + sum(counts[0..i)), i = 0..n (n + 1 values): the row offsets of n patients' slots in one table.
__global__ void __launch_bounds__(1024) exclusive_scan_i64_kernel(const int64_t* __restrict__ counts, int n, int64_t* __restrict__ offsets) {
  __shared__ int64_t s_warp[32];
  const int per = (n + 1023) / 1024;
  const int begin = min(threadIdx.x * per, n), end = min(begin + per, n);
  int64_t sum = 0;
  for (int i = begin; i < end; ++i) sum += counts[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int64_t w = s_warp[lane];
    int64_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  int64_t run = s_warp[warp] + incl - sum;
  for (int i = begin; i < end; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
  if (threadIdx.x == 1023) offsets[n] = run;
}

// ----------------------------------------------------------------------------- G2
__global__ void bbox_init_kernel(int32_t* bbox) {
  if (threadIdx.x < 6) bbox[threadIdx.x] = (threadIdx.x & 1) ? -1 : 0x7fffffff;
}

// kRowCol = false: the reference's xy-meshgrid convention (xi = (n/S) % H, yi = n/(H*S));  true: (col, row) of voxel n,
// i.e. the bounding box of the union mask over slices that generate_features starts from (tfds_dense_descriptor.py:257-260).
template <bool kRowCol>
__global__ void __launch_bounds__(256) voxel_bbox_kernel(const uint8_t* __restrict__ mask, int H, int W, int S,
                                                         int64_t total, int32_t* __restrict__ bbox) {
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
  // 16 mask bytes per thread per step
  const int64_t nvec = total >> 4;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvec + 1; v += (int64_t)gridDim.x * blockDim.x) {
    uint32_t wds[4] = {0, 0, 0, 0};
    int cnt = 16;
    if (v < nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(mask) + v);
      wds[0] = u.x; wds[1] = u.y; wds[2] = u.z; wds[3] = u.w;
    } else {  // tail
      cnt = static_cast<int>(total - (nvec << 4));
      for (int i = 0; i < cnt; ++i) wds[i >> 2] |= static_cast<uint32_t>(mask[(nvec << 4) + i]) << ((i & 3) * 8);
    }
    if ((wds[0] | wds[1] | wds[2] | wds[3]) == 0) continue;
    for (int i = 0; i < cnt; ++i) {
      if ((wds[i >> 2] >> ((i & 3) * 8)) & 0xffu) {
        const int64_t n = (v << 4) + i;
        const int zi = static_cast<int>(n % S);
        const int64_t q = n / S;
        const int xi = kRowCol ? static_cast<int>(q % W) : static_cast<int>(q % H);
        const int yi = kRowCol ? static_cast<int>(q / W) : static_cast<int>(q / H);
        lo[0] = min(lo[0], xi); hi[0] = max(hi[0], xi);
        lo[1] = min(lo[1], yi); hi[1] = max(hi[1], yi);
        lo[2] = min(lo[2], zi); hi[2] = max(hi[2], zi);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0 && hi[a] >= 0) {
      atomicMin(bbox + 2 * a, lo[a]);
      atomicMax(bbox + 2 * a + 1, hi[a]);
    }
  }
}

// Voxels inside the box, ascending flat index n = (yi*H + xi)*S + zi: a box in (yi, xi, zi) is
// lexicographically ordered exactly like n, so output row j maps to its voxel in closed form.
__global__ void __launch_bounds__(256)
voxel_gather_kernel(const float* __restrict__ img, const uint8_t* __restrict__ mask, int H, int S,
                    const int32_t* __restrict__ bbox, int32_t* __restrict__ out_flat, float* __restrict__ out_raw,
                    uint8_t* __restrict__ out_mask, int32_t* __restrict__ out_count, int cap) {
  const int x0 = bbox[0], x1 = bbox[1], y0 = bbox[2], y1 = bbox[3], z0 = bbox[4], z1 = bbox[5];
  const int64_t ex = (x1 >= x0) ? (x1 - x0 + 1) : 0, ey = (y1 >= y0) ? (y1 - y0 + 1) : 0, ez = (z1 >= z0) ? (z1 - z0 + 1) : 0;
  const int64_t count = ex * ey * ez;
  if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = static_cast<int32_t>(count);
  const int64_t lim = count < cap ? count : cap;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < lim; j += (int64_t)gridDim.x * blockDim.x) {
    const int zi = z0 + static_cast<int>(j % ez);
    const int64_t t = j / ez;
    const int xi = x0 + static_cast<int>(t % ex);
    const int yi = y0 + static_cast<int>(t / ex);
    const int64_t n = (static_cast<int64_t>(yi) * H + xi) * S + zi;
    out_flat[j] = static_cast<int32_t>(n);
    out_raw[j] = __ldg(img + n);
    out_mask[j] = __ldg(mask + n);
  }
}

}  // namespace vdr

static unsigned long long* g_g1_trace = nullptr;   // diagnostics only

extern "C" int vdr_debug_set_gather_trace(void* dev_u64x16) {
  g_g1_trace = static_cast<unsigned long long*>(dev_u64x16);
  return VDR_OK;
}

constexpr int kMaxCoopBlocks = 2048;   // upper bound of the fused kernel's grid (148 SMs x <= 8 blocks); sizes the block-totals scratch

static size_t g1_scan_bytes() { return (size_t)kMaxCoopBlocks * sizeof(int32_t); }

extern "C" size_t vdr_mask_gather_workspace_bytes(int S, int h, int w, int D) {
  return g1_scan_bytes() + (size_t)(h + w + S) * (size_t)(2 * (D / 6)) * sizeof(double);
}

// co-resident blocks of the fused kernel on the current device (cooperative launch), cached per device
template <bool BF16>
static int g1_coop_blocks() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vdr::g1_fused_kernel<BF16>, vdr::kGThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int n = per_sm * vdr::num_sms();
    cached[dev] = n > kMaxCoopBlocks ? kMaxCoopBlocks : n;
  }
  return cached[dev];
}

static int g1_check_geometry(const char* who, const uint8_t* mask, int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                             const int32_t* row_map, const int32_t* col_map, int S, int h, int w) {
  VDR_CHECK_ARG(mask && row_map && col_map, VDR_EINVAL, "%s: null pointer", who);
  VDR_CHECK_ARG(S > 0 && h > 0 && w > 0, VDR_EINVAL, "%s: bad shape", who);
  VDR_CHECK_ARG(mask_slice_stride >= 0 && mask_row_stride >= 0 && mask_col_stride >= 1, VDR_EINVAL, "%s: bad strides", who);
  VDR_CHECK_ARG((int64_t)S * h * w < 0x7fffffffLL, VDR_EINVAL, "%s: too many candidates", who);
  return VDR_OK;
}

extern "C" int vdr_mask_gather_table(const void* feat, int feat_dtype, int64_t ld_feat, int64_t feat_slice_rows,
                                     int64_t feat_row_pitch, int64_t feat_row0, const uint8_t* mask,
                                     int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                                     const int32_t* row_map, const int32_t* col_map, int S, int h, int w, int D,
                                     float* table_tok, int32_t* table_src, int src_cols, int32_t patient, const int64_t* row_offset,
                                     int64_t table_cap, int32_t* out_count, double pe_scale,
                                     const double* pe_div, const double* coef_host, void* workspace,
                                     size_t workspace_bytes, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(feat && table_tok && table_src && out_count && workspace, VDR_EINVAL, "vdr_mask_gather: null pointer");
  if (int rc = g1_check_geometry("vdr_mask_gather", mask, mask_slice_stride, mask_row_stride, mask_col_stride, row_map, col_map, S, h, w)) return rc;
  VDR_CHECK_ARG(D > 0 && table_cap >= 0 && (src_cols == 3 || src_cols == 4), VDR_EINVAL, "vdr_mask_gather: bad shape");
  VDR_CHECK_ARG(feat_slice_rows >= 0 && feat_row_pitch >= w && feat_row0 >= 0, VDR_EINVAL, "vdr_mask_gather: bad strides");
  VDR_CHECK_ARG(ld_feat >= D, VDR_EINVAL, "vdr_mask_gather: ld_feat (%lld) smaller than D (%d)", (long long)ld_feat, D);
  // 16-byte vector path when the rows allow it; otherwise an element-wise path (any D, e.g. the reference's D = 12 tests)
  const bool vec_ok = D % 8 == 0 && ld_feat % 8 == 0 && aligned16(feat) && aligned16(table_tok);
  VDR_CHECK_ARG(feat_dtype == VDR_DTYPE_BF16 || feat_dtype == VDR_DTYPE_F32, VDR_EINVAL, "vdr_mask_gather: bad feat_dtype");
  VDR_CHECK_ARG(workspace_bytes >= vdr_mask_gather_workspace_bytes(S, h, w, D), VDR_EWORKSPACE, "vdr_mask_gather: workspace too small (%zu < %zu)", workspace_bytes, vdr_mask_gather_workspace_bytes(S, h, w, D));
  VDR_CHECK_ARG(aligned16(workspace), VDR_EALIGN, "vdr_mask_gather: workspace must be 16-byte aligned");
  VDR_CHECK_ARG(pe_scale == 0.0 || (pe_div && coef_host), VDR_EINVAL, "vdr_mask_gather: positional encoding needs pe_div and coef_host");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  G1Args A{};
  A.g = G1Geom{mask, row_map, col_map, S, h, w, mask_slice_stride, mask_row_stride, mask_col_stride, feat_slice_rows, feat_row_pitch, feat_row0,
               (uint32_t)((int64_t)S * h * w)};
  A.feat = feat; A.ld_feat = ld_feat; A.D = D;
  A.out_tok = table_tok; A.out_src = table_src; A.out_count = out_count; A.cap = table_cap;
  A.row_offset = row_offset; A.src_cols = src_cols; A.patient = patient;
  A.pe.scale = pe_scale;
  A.pe.div = pe_div;
  A.pe.table = reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + g1_scan_bytes());
  A.npair2 = 2 * (D / 6);
  A.vec_ok = vec_ok ? 1 : 0;
  if (pe_scale != 0.0) {
    A.pe.w_orig = coef_host[0]; A.pe.h_orig = coef_host[1];
    A.pe.res0 = coef_host[2]; A.pe.res1 = coef_host[3]; A.pe.res2 = coef_host[4];
    A.pe.noise0 = coef_host[5]; A.pe.noise1 = coef_host[6]; A.pe.noise2 = coef_host[7];
    A.pe.mean_x = coef_host[8]; A.pe.mean_y = coef_host[9]; A.pe.mean_z = coef_host[10];
  }
  A.block_totals = static_cast<int32_t*>(workspace);
  A.trace = g_g1_trace;
  A.tiles = (int)(((int64_t)A.g.total + kTile - 1) / kTile);
  const bool bf16 = feat_dtype == VDR_DTYPE_BF16;
  const int max_blocks = bf16 ? g1_coop_blocks<true>() : g1_coop_blocks<false>();
  // enough warps for the emit phase (one per candidate row, eight per block), at least one block per tile, all co-resident
  int64_t want = ((int64_t)A.g.total + 7) / 8;
  if (want < A.tiles) want = A.tiles;
  int grid = (int)(want < max_blocks ? want : max_blocks);
  if (grid < 1) grid = 1;
  A.tiles_per_block = (A.tiles + grid - 1) / grid;
  void* params[] = {&A};
  const void* fn = bf16 ? reinterpret_cast<const void*>(g1_fused_kernel<true>) : reinterpret_cast<const void*>(g1_fused_kernel<false>);
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kGThreads), params, 0, s);
  if (e != cudaSuccess) return cuda_fail(e, "g1_fused_kernel (cooperative launch)");
  count_launch();
  VDR_CHECK_LAUNCH("g1_fused_kernel");
  return VDR_OK;
}

extern "C" int vdr_mask_gather(const void* feat, int feat_dtype, int64_t ld_feat, int64_t feat_slice_rows,
                               int64_t feat_row_pitch, int64_t feat_row0, const uint8_t* mask,
                               int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                               const int32_t* row_map, const int32_t* col_map, int S, int h, int w, int D,
                               float* out_tok, int32_t* out_src, int32_t* out_count, int cap, double pe_scale,
                               const double* pe_div, const double* coef_host, void* workspace,
                               size_t workspace_bytes, vdr_stream_t stream) {
  VDR_CHECK_ARG(cap >= 0, VDR_EINVAL, "vdr_mask_gather: bad shape");
  return vdr_mask_gather_table(feat, feat_dtype, ld_feat, feat_slice_rows, feat_row_pitch, feat_row0, mask, mask_slice_stride, mask_row_stride,
                               mask_col_stride, row_map, col_map, S, h, w, D, out_tok, out_src, 3, 0, nullptr, cap, out_count, pe_scale, pe_div,
                               coef_host, workspace, workspace_bytes, stream);
}

extern "C" int vdr_mask_count(const uint8_t* mask, int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                              const int32_t* row_map, const int32_t* col_map, int S, int h, int w, int64_t* out_count, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(out_count, VDR_EINVAL, "vdr_mask_count: null pointer");
  if (int rc = g1_check_geometry("vdr_mask_count", mask, mask_slice_stride, mask_row_stride, mask_col_stride, row_map, col_map, S, h, w)) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  G1Geom g{mask, row_map, col_map, S, h, w, mask_slice_stride, mask_row_stride, mask_col_stride, 0, 0, 0, (uint32_t)((int64_t)S * h * w)};
  const int tiles = (int)(((int64_t)g.total + kTile - 1) / kTile);
  cudaError_t e = cudaMemsetAsync(out_count, 0, sizeof(int64_t), s);
  if (e != cudaSuccess) return cuda_fail(e, "vdr_mask_count: memset");
  g1_count_kernel<<<tiles, kGThreads, 0, s>>>(g, reinterpret_cast<unsigned long long*>(out_count));
  count_launch();
  VDR_CHECK_LAUNCH("g1_count_kernel");
  return VDR_OK;
}

extern "C" int vdr_exclusive_scan_i64(const int64_t* counts, int n, int64_t* offsets, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(counts && offsets && n >= 0, VDR_EINVAL, "vdr_exclusive_scan_i64: bad arguments");
  exclusive_scan_i64_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(counts, n, offsets);
  count_launch();
  VDR_CHECK_LAUNCH("exclusive_scan_i64_kernel");
  return VDR_OK;
}

extern "C" int vdr_voxel_bbox(const uint8_t* mask, int H, int W, int S, int32_t* bbox, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(mask && bbox, VDR_EINVAL, "vdr_voxel_bbox: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0, VDR_EINVAL, "vdr_voxel_bbox: bad shape");
  VDR_CHECK_ARG((int64_t)H * W * S < 0x7fffffffLL, VDR_EINVAL, "vdr_voxel_bbox: volume too large for int32 flat indices");
  VDR_CHECK_ARG(aligned16(mask), VDR_EALIGN, "vdr_voxel_bbox: mask must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)H * W * S;
  bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
  int64_t blocks = ((total >> 4) + 1 + 255) / 256;
  const int64_t capb = (int64_t)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  voxel_bbox_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(mask, H, W, S, total, bbox);
  count_launch(2);
  VDR_CHECK_LAUNCH("voxel_bbox_kernel");
  return VDR_OK;
}

extern "C" int vdr_mask_bbox(const uint8_t* mask, int H, int W, int S, int32_t* bbox, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(mask && bbox, VDR_EINVAL, "vdr_mask_bbox: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0, VDR_EINVAL, "vdr_mask_bbox: bad shape");
  VDR_CHECK_ARG((int64_t)H * W * S < 0x7fffffffLL, VDR_EINVAL, "vdr_mask_bbox: volume too large for int32 flat indices");
  VDR_CHECK_ARG(aligned16(mask), VDR_EALIGN, "vdr_mask_bbox: mask must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)H * W * S;
  bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
  int64_t blocks = ((total >> 4) + 1 + 255) / 256;
  const int64_t capb = (int64_t)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  voxel_bbox_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(mask, H, W, S, total, bbox);
  count_launch(2);
  VDR_CHECK_LAUNCH("voxel_bbox_kernel");
  return VDR_OK;
}

extern "C" int vdr_voxel_gather(const float* img, const uint8_t* mask, int H, int W, int S, const int32_t* bbox,
                                int32_t* out_flat, float* out_raw, uint8_t* out_mask, int32_t* out_count, int cap,
                                vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(img && mask && bbox && out_flat && out_raw && out_mask && out_count, VDR_EINVAL, "vdr_voxel_gather: null pointer");
  VDR_CHECK_ARG(H > 0 && W > 0 && S > 0 && cap >= 0, VDR_EINVAL, "vdr_voxel_gather: bad shape");
  VDR_CHECK_ARG((int64_t)H * W * S < 0x7fffffffLL, VDR_EINVAL, "vdr_voxel_gather: volume too large for int32 flat indices");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = ((int64_t)cap + 255) / 256;
  const int64_t capb = (int64_t)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  if (blocks < 1) blocks = 1;
  voxel_gather_kernel<<<(unsigned)blocks, 256, 0, s>>>(img, mask, H, S, bbox, out_flat, out_raw, out_mask, out_count, cap);
  count_launch();
  VDR_CHECK_LAUNCH("voxel_gather_kernel");
  return VDR_OK;
}
