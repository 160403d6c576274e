/*
 * vdr.h -- C ABI of libvdr.so, the B200 (sm_100a) kernels behind the vit-deep-radiomics hot path.
 *
 * The reference (larosi/vit-deep-radiomics) has no FFI: every GPU op is a stock PyTorch or
 * third-party module call made from Python.  Each entry point below
 * therefore cites the reference *call site* it replaces (file:line under /root/reference/src).
 * The Python drop-in modules in vit_deep_radiomics_b200/ bind these with ctypes (see
 * INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory (incl. workspaces) and keeps it alive until `stream` reaches the op;
 *   - the library never allocates device memory, never synchronises, never changes the device;
 *   - return 0 on success, a negative VDR_E* for argument errors, a positive cudaError_t otherwise;
 *     vdr_last_error_string() gives a thread-local description of the last non-zero return;
 *   - matrices are row-major; "ld*" are leading dimensions in ELEMENTS;
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = float.
 */
#ifndef VDR_H_
#define VDR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vdr_stream_t;

#define VDR_OK            0
#define VDR_EINVAL       -1   /* bad shape / null pointer / unsupported value */
#define VDR_EALIGN       -2   /* pointer or leading dimension not 16-byte aligned */
#define VDR_EWORKSPACE   -3   /* workspace too small */
#define VDR_EUNSUPPORTED -4   /* valid request this build does not implement */
#define VDR_EDRIVER      -5   /* could not obtain a driver entry point (TMA descriptor encode) */

#define VDR_DTYPE_BF16 0
#define VDR_DTYPE_F32  1

/* epilogue selectors for vdr_gemm */
#define VDR_EPI_BIAS          0   /* C = A W^T + b                                         */
#define VDR_EPI_BIAS_GELU     1   /* C = gelu_erf(A W^T + b)                                */
#define VDR_EPI_BIAS_RESIDUAL 2   /* C = A W^T + b + R                                      */

int         vdr_version(void);
const char* vdr_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t    vdr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Dropout of the classifier's train mode (nn.TransformerEncoderLayer(dropout=0.1 | 0.5), MLPLayer(0.1): models_archs.py:51,58,135,
 * 187-199; active under model.train(), train_models.py:652).  Counter based: element (row r, column c) of dropout site `site` is
 * KEPT iff the 16-bit lane (c & 7) of Philox4x32-10(counter = (c >> 3, r_lo, r_hi, site), key = (seed_lo, seed_hi)) is >= thr16,
 * and kept values are multiplied by 65536 / (65536 - thr16): p = thr16 / 65536.  The backward kernels regenerate the masks from
 * the same (seed + *seed_offset, site), nothing is stored.  thr16 == 0 (or a NULL pointer) = no dropout, bit-identical to the p = 0 path.
 * RNG streams cannot match PyTorch's: parity is statistical (keep rate, 1 / (1 - p) scale) plus exact gradients against fp32
 * autograd with the mask exported by vdr_dropout_mask. */
typedef struct {
  uint64_t seed;
  uint32_t site;
  uint32_t thr16;
  const uint64_t* seed_offset;   /* optional DEVICE scalar added to `seed` by the kernels (NULL = 0): a training step captured in a
                                    CUDA graph bumps it on the device, so every replay draws fresh masks from a fixed launch */
} vdr_dropout;
/* out = x * mask / (1 - p): bf16 (rows, cols) with row pitches ldx / ldo (elements), cols % 8 == 0.  The backward of a dropout
 * that the forward applied inside a GEMM epilogue. */
int vdr_dropout_apply(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int cols, const vdr_dropout* drop,
                      vdr_stream_t stream);
/* out[r * cols + c] = 1 if element (r, c) of the site is kept: the mask as every kernel of this library computes it. */
int vdr_dropout_mask(uint8_t* out, int64_t rows, int64_t cols, const vdr_dropout* drop, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores: C[M,N] = epi(A[M,K] * W[N,K]^T + bias[N] (+ R)).
 * A, W bf16 (K contiguous), fp32 accumulation in TMEM, bias f32 (may be NULL), C bf16 or f32.
 * Replaces: torch.nn.Linear inside the backbone (tfds_dense_descriptor.py:123, K3 in SURVEY 2.4)
 * and inside nn.TransformerEncoderLayer / MLPLayer (models_archs.py:130-137,146,193-199).
 *
 * Row remapping (used to write patch tokens behind the CLS token and to add pos-embed):
 *   out_row(m) = (out_group > 0) ? (m / out_group) * out_group_stride + out_offset + m % out_group : m
 *   res_row(m) = (res_mod   > 0) ? res_offset + m % res_mod : out_row(m)
 * Requirements: lda/ldw/ldc/ldr % 8 == 0, 16-byte aligned pointers.  N is arbitrary (a ragged last
 *   column group is stored element-wise) and K is arbitrary
 *   (the K tail is zero-filled by TMA), e.g. K = 588 for 14x14 patches with lda = ldw = 592.
 */
typedef struct {
  const void* A;  int64_t lda;
  const void* W;  int64_t ldw;
  const float* bias;
  const void* R;  int64_t ldr;  int r_dtype;      /* residual (bf16 or f32), only for EPI_BIAS_RESIDUAL */
  void* C;        int64_t ldc;  int c_dtype;
  int M, N, K;
  int epilogue;
  int out_group, out_group_stride, out_offset;    /* 0,0,0 = identity */
  int res_mod, res_offset;                        /* 0,0 = same row as the output */
  /* LayerNorm folded into the GEMMs either side of it (all NULL / 0 = off), the ViT blocks' `x -> LN -> Linear` pairs:
   *   LN(x) W^T + b  =  rstd_m * (x W'^T - mean_m * colsum_n) + b'_n   with W' = gamma (.) W, colsum_n = sum_k W'[n,k],
   *   b' = b + W beta (vdr_fold_layernorm).  The GEMM then reads the raw residual stream x and normalises in its epilogue.
   * stats_out (producer, EPI_BIAS_RESIDUAL only, N % 64 == 0, no row remapping, bf16 C): per output row and per 64-column
   *   slot, (sum, sum of squares) of the values written: float pairs laid out [N/64][M].  Deterministic (no atomics).
   * ln_stats (consumer): such a table for the rows of A, `ln_slots` slots (summed in slot order); the row mean / rstd over
   *   K elements with `ln_eps`; W must be the folded W', bias the folded b', ln_colsum (N) f32.  bf16 C, N % 32 == 0. */
  const float* ln_stats;  int ln_slots;  float ln_eps;
  const float* ln_colsum;
  float* stats_out;
  /* EPI_BIAS_RESIDUAL only: C = R + dropout(A W^T + bias), element (m, n) of the site = output element (m, n) -- dropout1 /
   * dropout2 of nn.TransformerEncoderLayer (the sub-layer output is dropped BEFORE the residual add).  bf16 C, N % 32 == 0,
   * no row remapping, no folded LayerNorm / statistics (the classifier's shapes).  thr16 == 0 = off. */
  vdr_dropout drop;
} vdr_gemm_args;

int vdr_gemm(const vdr_gemm_args* args, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Patch extraction (im2col) fused with the reference's host-side image preparation.
 * Replaces: prepare_image (tfds_dense_descriptor.py:30-48: gray2rgb + HWC->NCHW + float32 cast)
 * and the data movement of the patch-embedding conv (K1).  Output A[(b,py,px), (c,iy,ix)] bf16,
 * i.e. the K-major operand of vdr_gemm with W = conv weight viewed as (d, 3*p*p).
 *   src: f32, element strides (in elements) for batch/channel/row/col; channel stride 0 = gray
 *   image replicated to 3 channels (gray2rgb, :41).
 */
int vdr_im2col_patches(const float* src, int64_t sb, int64_t sc, int64_t sy, int64_t sx,
                       int B, int H, int W, int patch, void* A_bf16, vdr_stream_t stream);

/* Batched slice staging for a whole volume: (H, W, S) f32 (slice index fastest, the np.dstack layout of
 * get_voxels, tfds_dense_descriptor.py:353-362) -> (S, ch, cw) bf16 slices of the crop window
 * [y0, y0+ch) x [x0, x0+cw) (crop_image, :265), transposed through shared memory so reads and writes are
 * both coalesced; then im2col of those gray slices with the 3 channels replicated (gray2rgb, :41). */
int vdr_volume_to_slices(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw,
                         void* slices_bf16, vdr_stream_t stream);
int vdr_im2col_gray_bf16(const void* slices_bf16, int B, int H, int W, int patch, void* A_bf16,
                         vdr_stream_t stream);
/* The same staging with prepare_image's resize (tfds_dense_descriptor.py:40-44, skimage.transform.resize of a float image)
 * for crop windows whose size differs from the backbone input: optional Gaussian anti-aliasing (only when an axis shrinks:
 * sigma = (in/out - 1)/2, truncate 4, boundary 'mirror') + order-1 resampling at (o + 0.5) * in/out - 0.5 with mirrored
 * indices (scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True), which is what skimage >= 0.19 calls), f32 arithmetic,
 * bf16 slices (S, OH, OW) out.  workspace: vdr_volume_to_slices_resized_workspace_bytes (0 when nothing shrinks). */
size_t vdr_volume_to_slices_resized_workspace_bytes(int S, int ch, int cw, int OH, int OW);
int vdr_volume_to_slices_resized(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw, int OH, int OW,
                                 void* slices_bf16, void* workspace, size_t workspace_bytes, vdr_stream_t stream);

/* The same staging (plain crop when OH x OW == ch x cw, else the resize above) into a CELL-PADDED layout: every cell_in x cell_in patch of
 * the (OH, OW) slice lands in the top-left corner of a cell_out x cell_out cell of a (S, OH / cell_in * cell_out, OW / cell_in * cell_out)
 * bf16 image whose other pixels the caller has zeroed once.  A 14-pixel patch grid (ViT-L/14, DINOv2; reference
 * src/tfds_dense_descriptor.py:70-88,110-139) becomes a 16-pixel one that vdr_patch_embed_gemm's TMA im2col view addresses; the patch
 * weights get zero columns at the pad positions. */
int vdr_volume_to_slices_cells(const float* vol, int H, int W, int S, int y0, int x0, int ch, int cw, int OH, int OW,
                               int cell_in, int cell_out, void* slices_bf16, void* workspace, size_t workspace_bytes,
                               vdr_stream_t stream);

/* Patch embedding as ONE TMA-fed im2col GEMM (K1: Conv2d(3, d, p, p) + position embedding of the backbone the reference
 * calls at tfds_dense_descriptor.py:123): the A operand is never materialised -- a 5-D tensor map (ix, iy, px, py, image)
 * over the bf16 pictures delivers each 128-patch x 64-k tile of the im2col matrix straight into the swizzled shared-memory
 * layout tcgen05.mma reads.  images (B*C, H, W) bf16, C = 1 (gray picture reused for the 3 input channels, gray2rgb :41) or 3;
 * Wpe (d, 3*p*p) bf16 K-major with k = (c, iy, ix) = the conv weight as stored in the checkpoint; pos (N, d) f32 with
 * N = patches + 1 (row 0 belongs to the CLS token); writes X[b*N + 1 + patch, :] = patch . Wpe^T + bias + pos[1 + patch] (bf16).
 * vdr_patch_embed_supported: 1 when the geometry tiles into TMA boxes (p == 16, 128 % (W/p) == 0, patches % 128 == 0:
 * 16-pixel patches on 256 / 512 / 1024 / 2048-wide pictures); otherwise use vdr_im2col_* + vdr_gemm. */
int vdr_patch_embed_supported(int H, int W, int patch);
int vdr_patch_embed_gemm(const void* images_bf16, int B, int C, int H, int W, int patch, const void* Wpe_bf16,
                         int64_t ldw, const float* bias, const float* pos, void* X_bf16, int64_t ldx, int d,
                         vdr_stream_t stream);
/* The same for gray pictures (B, H, W) against CHANNEL-SUMMED weights Wsum (d, p*p) = Wpe[:, 0] + Wpe[:, 1] + Wpe[:, 2] (summed in
 * f32, rounded to bf16 once): gray2rgb (:41) feeds one picture to all three input channels, so the convolution is
 * patch . (W_r + W_g + W_b) -- K = p*p, a third of the MMA work and one read of the picture instead of three.
 * token_offset = rows in front of every image's patch tokens in X and pos: 1 (the CLS row of a ViT) or 0 (SAM's image encoder,
 * `model.image_encoder` of load_medsam, tfds_dense_descriptor.py:91-107). */
int vdr_patch_embed_gemm_gray(const void* images_bf16, int B, int H, int W, int patch, const void* Wsum_bf16,
                              int64_t ldw, const float* bias, const float* pos, void* X_bf16, int64_t ldx, int d,
                              int token_offset, vdr_stream_t stream);

/* CLS rows of the token matrix: X[b*N + 0, :] = cls[:] + pos[0, :]   (f32 params -> bf16 tokens) */
int vdr_write_cls_rows(const float* cls, const float* pos0, void* X_bf16, int B, int N, int d,
                       vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The whole backbone forward as one call: patch embedding -> depth x (LN, QKV, attention, proj + residual, LN, fc1 + GELU,
 * fc2 + residual) -> final LayerNorm, every kernel enqueued on `stream` (no host synchronisation, no allocation).
 * Replaces: `model.image_encoder(img_tensor)` for a batch of slices (tfds_dense_descriptor.py:123; pre-norm ViT with the
 * timm / DINOv2 parameter set: patch_embed, cls_token, pos_embed, blocks[i].{norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2}, norm).
 *   weights   device pointers; GEMM weights bf16 row-major (out_features, in_features), everything else f32.
 *             `blocks` is a HOST array of `depth` entries.  pe_w is (dim, pe_ldw >= 3*patch^2) with k = (c, iy, ix).
 *   images    (B*C, H, W) bf16, C = 1 (gray slices reused for the 3 input channels) or 3
 *   tokens    (B*N, dim) f32 with row pitch ld_out, N = patches + 1; row b*N is image b's CLS token
 *   workspace >= vdr_vit_forward_workspace_bytes(weights, B), 256-byte aligned
 */
typedef struct {
  const float *n1w, *n1b;
  const void* qkv_w;  const float* qkv_b;     /* (3*dim, dim) */
  const void* proj_w; const float* proj_b;    /* (dim, dim) */
  const float *n2w, *n2b;
  const void* fc1_w;  const float* fc1_b;     /* (4*dim, dim) */
  const void* fc2_w;  const float* fc2_b;     /* (dim, 4*dim) */
  /* optional (all six or none): norm1 folded into qkv and norm2 into fc1 by vdr_fold_layernorm.  When present the forward runs
   * no LayerNorm kernel inside the blocks: the residual GEMMs emit row statistics and the qkv / fc1 GEMMs normalise in their epilogue. */
  const void* qkv_wf; const float* qkv_bf; const float* qkv_cs;
  const void* fc1_wf; const float* fc1_bf; const float* fc1_cs;
} vdr_vit_block;

typedef struct {
  int dim, depth, heads, patch, H, W;
  float eps;                                  /* LayerNorm epsilon (<= 0: 1e-6) */
  const void* pe_w; int64_t pe_ldw; const float* pe_b;
  const float* cls;                           /* (dim) */
  const float* pos;                           /* (N, dim) */
  const float *norm_w, *norm_b;
  const vdr_vit_block* blocks;
  const void* pe_w_gray; int64_t pe_gray_ldw; /* optional: channel-summed patch weights (dim, patch^2), used for C = 1 pictures */
} vdr_vit_weights;

size_t vdr_vit_forward_workspace_bytes(const vdr_vit_weights* weights, int B);
int vdr_vit_forward(const vdr_vit_weights* weights, const void* images_bf16, int B, int C, float* tokens, int64_t ld_out,
                    void* workspace, size_t workspace_bytes, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dim, one warp per row, 16-byte vector loads, fp32 statistics.
 * Replaces: Block.norm1/norm2 + final norm of the backbone (K2) and nn.LayerNorm in the
 * classifier (models_archs.py:145 and the post-norms inside nn.TransformerEncoderLayer).
 * x bf16 (rows, d) -> y (bf16 or f32).  d % 8 == 0, d <= 4096.  Optional mean/rstd (f32, rows)
 * are written when non-NULL (saved for the backward pass).
 */
int vdr_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta,
                      void* y, int64_t ldy, int y_dtype, float* mean, float* rstd,
                      int rows, int d, float eps, vdr_stream_t stream);

/* Row statistics in the layout vdr_gemm's `ln_stats` reads, one slot: stats[2*r] = sum_k x[r,k], stats[2*r+1] = sum_k x[r,k]^2. */
int vdr_row_stats(const void* x_bf16, int64_t ldx, int rows, int d, float* stats, vdr_stream_t stream);

/* Fold a LayerNorm (gamma, beta over K) into the Linear (W (N,K) bf16, bias (N) f32 or NULL) that consumes it:
 * Wf = bf16(gamma (.) W), colsum[n] = sum_k float(Wf[n,k]) (of the ROUNDED weights, so the identity holds for what the
 * tensor cores multiply), bias_f[n] = bias[n] + sum_k beta[k] W[n,k].  Weight preparation, run once per checkpoint. */
int vdr_fold_layernorm(const void* W_bf16, int64_t ldw, const float* bias, const float* gamma, const float* beta, int N, int K,
                       void* Wf_bf16, int64_t ldwf, float* bias_f, float* colsum, vdr_stream_t stream);

/* dx, dgamma, dbeta of the LayerNorm above.  dgamma/dbeta are ACCUMULATED (+=) into f32 buffers. */
int vdr_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx,
                      const float* gamma, const float* mean, const float* rstd,
                      void* dx, int64_t lddx, float* dgamma, float* dbeta,
                      int rows, int d, vdr_stream_t stream);

/* CLS concat + LayerNorm in one pass: Y[0] = LN(cls), Y[1+i] = LN(X[i]).
 * Replaces: torch.cat([cls, x]) + self.norm(x)  (models_archs.py:143-145).
 * X f32 (n, d) tokens as produced by the gather; Y bf16 (n+1, d); mean/rstd (n+1) optional. */
int vdr_cls_concat_layernorm_fwd(const float* X, const float* cls, const float* gamma,
                                 const float* beta, void* Y_bf16, float* mean, float* rstd,
                                 int n, int d, float eps, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused flash-style self-attention forward on tcgen05 (head_dim 64).
 * Replaces: Attention.forward of the backbone (K4) and the SDPA inside
 * nn.TransformerEncoderLayer (models_archs.py:146).
 * qkv bf16 (B*N, 3*d) as written by the QKV GEMM: [q(h,64) | k(h,64) | v(h,64)] per token;
 * out bf16 (B*N, d).  lse (B, h, N) f32 optional (log-sum-exp, for the backward pass).
 * drop (NULL = none): attention dropout of nn.MultiheadAttention -- the NORMALISED probabilities are dropped before P V (the row
 * sums keep every key); element (row (b*heads + h)*N + q, column k) of the site.
 * Numerics: exact softmax for any finite input.  The kernel takes the reference maximum of a row from its first
 * 128-key block and runs the later blocks maximum-free; a tile in which a later score exceeds that maximum by more than ~88 nats
 * (exp2 would overflow) is detected from its row sums and recomputed exactly before the kernel returns -- slower for such tiles,
 * never wrong.  The same holds for vdr_flash_attn_relpos_fused_fwd below.
 */
int vdr_flash_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse,
                       int B, int N, int heads, float scale, const vdr_dropout* drop, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G1: tumour-mask gather of ViT descriptors into a per-patient point cloud.
 * Replaces: PETCTDataset3D._get_features (train_models.py:143-182): nearest mask resize (:151),
 * (h,w,S) flatten order (:159-163), boolean row gather (:180) and, optionally, the 3-D
 * sinusoidal positional encoding added as PE/4 (:30-44,178-180; computed in f64).
 *
 *   feat      descriptors, slice-major as the backbone writes them (bf16 or f32): token (k,a,b) is
 *             row  k*feat_slice_rows + feat_row0 + a*feat_row_pitch + b  of a matrix with ld_feat
 *             elements per row -- a dense (S,h,w,D) tensor is (h*w, w, 0); the ROI (r0:,c0:) of the
 *             backbone's token matrix (CLS first, grid gh x gw) is (gh*gw+1, gw, 1+r0*gw+c0),
 *             which is extract_roi (visualization_utils.py:115-125) without a copy
 *   mask      u8 pixel masks: value (k,r,c) at mask[k*mask_slice_stride + r*mask_row_stride + c*mask_col_stride]
 *             ((HM*WM, WM, 1) for slice-major masks, (1, W*S, S) for an (H,W,S) volume mask read in place);
 *             row_map[h], col_map[w] (int32) give the source pixel row/col of each feature-grid
 *             row/col (the order-0 resize index maps of :151)
 *   out_tok   (cap, D) f32 packed tokens in ascending n = a*(w*S) + b*S + k   (stable)
 *   out_src   (cap, 3) int32 (slice k, row a, col b) per token
 *   out_count int32[1] number selected (may exceed cap: rows beyond cap are not written)
 *   pe        if pe_scale != 0: tokens = f32(f64(token) + pe_scale * PE3D(x,y,z)) with
 *             x = (xi/w)*w_orig*res[0] - mean_x + noise[0] ... exactly as :166-176, in f64;
 *             pe_div (device, f64[D/6]) = 10000^(6i/D) as computed by the host (:34);
 *             coef_host (HOST, f64[11]) = {w_orig, h_orig, res0, res1, res2,
 *                                          noise0, noise1, noise2, mean_x, mean_y, mean_z}
 *             The encoding is evaluated once per distinct coordinate ((h + w + S) rows of 2*(D/6) f64 in the
 *             workspace) and added row-wise: bit-identical to evaluating it per token.
 *   workspace >= vdr_mask_gather_workspace_bytes(S,h,w,D), 16-byte aligned
 * One cooperative launch (g1_fused_kernel): the mask is read once, tile ballots stay in shared memory between the count and
 * the rank phase, the scan is a sum over the co-resident blocks' totals, rows are emitted by one warp per OUTPUT row.
 */
size_t vdr_mask_gather_workspace_bytes(int S, int h, int w, int D);
int vdr_mask_gather(const void* feat, int feat_dtype, int64_t ld_feat, int64_t feat_slice_rows,
                    int64_t feat_row_pitch, int64_t feat_row0,
                    const uint8_t* mask, int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                    const int32_t* row_map, const int32_t* col_map,
                    int S, int h, int w, int D,
                    float* out_tok, int32_t* out_src, int32_t* out_count, int cap,
                    double pe_scale, const double* pe_div, const double* coef_host,
                    void* workspace, size_t workspace_bytes, vdr_stream_t stream);

/* The same gather writing into a slot of a SHARED point-cloud table (the table the ranks assemble with one
 * all-gather; replaces the reference's per-patient loop + merge, tfds_dense_descriptor.py:421, merge_dataframe_features.py:12-30):
 * row r of this call lands at table row (*row_offset + r) -- row_offset is a DEVICE scalar (NULL = 0), typically one entry of
 * the exclusive scan over every patient's count (vdr_mask_count -> counts all-gather -> vdr_exclusive_scan_i64), so the kernel
 * writes at the rank's offset of the all-gather buffer with no host round trip in between.  table_src has src_cols int32 per
 * row: 3 = (slice,row,col), 4 = (patient,slice,row,col) with `patient` written by the kernel.  Rows at or beyond table_cap are
 * not written; out_count (int32[1]) receives this call's count. */
int vdr_mask_gather_table(const void* feat, int feat_dtype, int64_t ld_feat, int64_t feat_slice_rows,
                          int64_t feat_row_pitch, int64_t feat_row0,
                          const uint8_t* mask, int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                          const int32_t* row_map, const int32_t* col_map,
                          int S, int h, int w, int D,
                          float* table_tok, int32_t* table_src, int src_cols, int32_t patient, const int64_t* row_offset,
                          int64_t table_cap, int32_t* out_count,
                          double pe_scale, const double* pe_div, const double* coef_host,
                          void* workspace, size_t workspace_bytes, vdr_stream_t stream);
/* Number of tokens vdr_mask_gather would select (same mask / index-map arguments), as a device int64 -- the row count every
 * rank exchanges before any rank emits. */
int vdr_mask_count(const uint8_t* mask, int64_t mask_slice_stride, int64_t mask_row_stride, int64_t mask_col_stride,
                   const int32_t* row_map, const int32_t* col_map, int S, int h, int w, int64_t* out_count, vdr_stream_t stream);
/* Diagnostics: when set to a device buffer of 16 uint64, the gather kernel stamps %globaltimer of its first ([0..6]) and last
 * ([8..14]) block at its phase boundaries (start, table, count, sync, ranks, sync, emit).  NULL (default) switches it off. */
int vdr_debug_set_gather_trace(void* dev_u64x16);
/* offsets[i] = sum(counts[0..i)), i = 0..n (n + 1 values, device): table row offsets of n patients' slots. */
int vdr_exclusive_scan_i64(const int64_t* counts, int n, int64_t* offsets, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * V4 / N2: the offline augmentation of the extraction loop on the device.
 * Replaces: flip_image (tfds_dense_descriptor.py:306-325) and rotate_image (:328-350: scipy.ndimage.rotate(x, angle, axes=(0, 1),
 * reshape=False, mode='nearest'), cubic spline; np.clip(image, 0, 1); mask > 0), applied to whole volumes 12 times per patient
 * (:463-466).  The rotation restates scipy's algorithm operation for operation in IEEE double (12-pixel edge padding, cubic
 * prefilter with its 'reflect' initialisers, unmapped coordinates, clamped taps, 16-term sum): the results are bit-identical to
 * scipy 1.18 -- masks AND float32 images (tests/test_gpu_augment.py; the reference's bool mask is the interpolated value cast to
 * unsigned char, i.e. |t| >= 1, which depends on the last bit of t inside the mask).
 *   src / dst  (H, W, planes) volumes, plane index fastest (np.dstack layout; planes = slices x channels)
 *   src_kind   0: f32 image -> f32 image clipped to [0, 1];  1: bool mask (u8 storage) -> u8 mask;  2: uint8 mask -> u8 mask
 *   flip       0 none, 1 'horizontal' (columns reversed), 2 'vertical' (rows reversed); applied before the rotation as in :463-466
 *   rotate     0: flip only (masks are re-binarised);  1: rotate with xform_host (HOST, f64[6]) = {m00, m01, m10, m11, off0, off1}
 *              = scipy's rot_matrix [[c, s], [-s, c]] (special.cosdg / sindg) and offset = in_center - rot_matrix @ out_center;
 *              z_n_rows = pow(z, H + 24), z_n_cols = pow(z, W + 24) with z = -0x1.126145e9ecd56p-2 (libm pow on the host)
 *   workspace  >= vdr_rotate_workspace_bytes(H, W, planes) (the padded f64 planes), 8-byte aligned; unused when rotate == 0
 */
size_t vdr_rotate_workspace_bytes(int H, int W, int planes);
int vdr_flip_rotate_volume(const void* src, int src_kind, void* dst, int H, int W, int planes, int flip, int rotate,
                           const double* xform_host, double z_n_rows, double z_n_cols, void* workspace, size_t workspace_bytes,
                           vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G2: voxel point cloud.  Replaces: to_pointcloud_df + the caller's mask_box filter
 * (create_pointcloud_dataframe.py:15-31,78).
 *   img f32 (H,W,S), mask u8 (H,W,S).  Pass 1 reduces the index-space bounding box of mask>0
 *   (bbox int32[6] = xi_min,xi_max,yi_min,yi_max,zi_min,zi_max with the reference's xy-meshgrid
 *   index convention xi=(n/S)%H, yi=n/(H*S), zi=n%S; empty mask -> min>max).  Pass 2 compacts
 *   every voxel inside the box: out_flat int32 (flat index n), out_raw f32, out_mask u8,
 *   out_count int32[1].  Order is ascending n (stable).  A box in (yi, xi, zi) is ordered exactly
 *   like n = (yi*H + xi)*S + zi, so pass 2 needs no scan: output row j maps to its voxel in closed form.
 */
int vdr_voxel_bbox(const uint8_t* mask, int H, int W, int S, int32_t* bbox, vdr_stream_t stream);
/* Same reduction in (col, row, slice) order: bbox = col_min, col_max, row_min, row_max, slice_min, slice_max of mask > 0
 * = the bounding box of the union mask over slices that generate_features starts from (tfds_dense_descriptor.py:257-260). */
int vdr_mask_bbox(const uint8_t* mask, int H, int W, int S, int32_t* bbox, vdr_stream_t stream);
int vdr_voxel_gather(const float* img, const uint8_t* mask, int H, int W, int S, const int32_t* bbox,
                     int32_t* out_flat, float* out_raw, uint8_t* out_mask, int32_t* out_count, int cap,
                     vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Training-only kernels of the point-cloud classifier (backward of models_archs.py:141-147 as driven by
 * loss.backward() at train_models.py:683).  dgrad / wgrad reuse vdr_gemm on transposed operands.
 */
/* h = dropout(gelu_erf(z)), dz = dh * mask / (1 - p) * gelu'(z); bf16, n elements (n % 8 == 0).  drop (NULL = none) is the
 * feed-forward block's inner dropout; with it the n elements are a contiguous (n / cols, cols) matrix, cols % 8 == 0. */
int vdr_gelu_fwd(const void* z, void* h, int64_t n, int cols, const vdr_dropout* drop, vdr_stream_t stream);
int vdr_gelu_bwd(const void* dh, const void* z, void* dz, int64_t n, int cols, const vdr_dropout* drop, vdr_stream_t stream);
/* out (cols, rows; pitch ld_out) = in (rows, cols; pitch ld_in)^T, bf16, through a shared-memory tile. */
int vdr_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols,
                       vdr_stream_t stream);
/* out_accum[c] += sum_r in[r][c]  (bias gradients; bf16 in, f32 accumulate). */
int vdr_colsum_bf16(const void* in, int64_t ld, int rows, int cols, float* out_accum, vdr_stream_t stream);
/* Fused flash-style attention backward on tcgen05 (head_dim 64): dqkv (B*N, 3d) bf16 = [dQ | dK | dV] from the packed qkv of
 * the forward, its output O, the upstream gradient dO and the log-sum-exp the forward saved.  One CTA per (128-key block, head,
 * image) keeps dK / dV in TMEM while it walks the query blocks; S, dP, P, dS never leave the SM; dQ is accumulated in an f32
 * scratch with vector reductions and converted at the end.  Replaces: autograd of F.scaled_dot_product_attention inside
 * nn.TransformerEncoderLayer (models_archs.py:146) on the training path (loss.backward(), train_models.py:683).
 * workspace >= vdr_flash_attn_bwd_workspace_bytes(B, N, heads), 16-byte aligned. */
size_t vdr_flash_attn_bwd_workspace_bytes(int B, int N, int heads);
int vdr_flash_attn_bwd(const void* qkv, int64_t ld_qkv, const void* O, const void* dO, int64_t ld_o, const float* lse,
                       void* dqkv, int64_t ld_dqkv, int B, int N, int heads, float scale, const vdr_dropout* drop, void* workspace,
                       size_t workspace_bytes, vdr_stream_t stream);
/* Attention backward pieces of the earlier, unfused path (scores materialised per head; N <= ~16k tokens in this model):
 *   delta[h][i] = sum_c dO[i][h*64+c] * O[i][h*64+c]
 *   P = exp(S*scale - lse) (0 for key columns >= N),  dS = P * (dP - delta) * scale     (S, dP f32; P, dS bf16) */
int vdr_attn_delta(const void* dO, const void* O, int64_t ld, int N, int heads, float* delta, vdr_stream_t stream);
int vdr_attn_p_ds(const float* S, const float* dP, const float* lse, const float* delta, void* P, void* dS,
                  int N, int64_t ldp, float scale, vdr_stream_t stream);
/* Backward of vdr_cls_concat_layernorm_fwd: accumulates dgamma, dbeta and dcls (the point cloud X is data). */
int vdr_cls_concat_layernorm_bwd(const void* dY, const float* X, const float* cls, const float* gamma,
                                 const float* mean, const float* rstd, float* dgamma, float* dbeta, float* dcls,
                                 int n, int d, vdr_stream_t stream);
/* Classification head MLPLayer (models_archs.py:186-200): zc = W1 cls + b1, logits = drop(W2 drop(gelu(zc)) + b2), and its
 * backward (accumulates dW1, db1, dW2, db2; writes dcls = W1^T dzc + dcls_in).  f32 weights.  drop (NULL = none): the layer's two
 * dropouts, hidden units = row 0 of the site, outputs = row 1 (the reference really drops the logits, :198-199). */
int vdr_cls_head_fwd(const void* cls_bf16, const float* W1, const float* b1, const float* W2, const float* b2,
                     float* zc, float* logits, int d, int H1, int C, const vdr_dropout* drop, vdr_stream_t stream);
int vdr_cls_head_bwd(const void* cls_bf16, const float* W1, const float* W2, const float* zc, const float* dlogits,
                     const float* dcls_in, float* dW1, float* db1, float* dW2, float* db2, float* dcls,
                     int d, int H1, int C, const vdr_dropout* drop, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Bimodal PET+CT classifier pieces (TransformerNoduleBimodalClassifier, models_archs.py:38-124).
 * Linear layer on one f32 vector: y = W x + b (W (rows, cols) row-major); backward ACCUMULATES dW, db (optional) and dx. */
int vdr_linear_vec_fwd(const float* W, const float* b, const float* x, float* y, int rows, int cols, vdr_stream_t stream);
int vdr_linear_vec_bwd(const float* W, const float* x, const float* dy, float* dW, float* db, float* dx, int rows, int cols,
                       vdr_stream_t stream);
/* Cross attention of CrossAttentionLayer (:174-183) for the single query row the model keeps (`x_attn[:, 0, :]`, :102-103):
 * q0 (d) f32 = projected CLS query; kv (n, 2d) bf16 = [K | V] projections of the other modality's tokens; head_dim 64.
 * fwd: p (heads, n) f32 = softmax(q0_h K_h^T scale) (saved), o (d) f32 = p V.  bwd: dq0 (d) f32, dkv (n, 2d) bf16
 * (scratch: heads * n floats). */
int vdr_cross_cls_attn_fwd(const float* q0, const void* kv, int64_t ld_kv, int n, int heads, float scale, float* p, float* o,
                           vdr_stream_t stream);
int vdr_cross_cls_attn_bwd(const float* q0, const void* kv, int64_t ld_kv, const float* p, const float* d_o, int n, int heads,
                           float scale, float* dq0, void* dkv, int64_t ld_dkv, float* scratch, vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MedSAM / SAM ViT-B image encoder pieces: what `model.image_encoder(img_tensor)` of
 * sam_model_registry['vit_b'] (tfds_dense_descriptor.py:104,123; third-party segment_anything, un-vendored) needs beyond the
 * plain-ViT kernels above.  All matrices are token-major bf16 (rows = tokens), head_dim 64.
 *
 * vdr_window_rows: window_partition / window_unpartition of the encoder blocks.  to_windows != 0: src (B*H*W, d) ->
 *   dst (B*nwh*nww*ws*ws, d), nwh = ceil(H/ws), rows outside H x W written as zeros (the padding is applied after norm1);
 *   to_windows == 0: the inverse, pad rows dropped.
 * Relative-position tables: rel_pos_h (2*Sh-1, 64) and rel_pos_w (2*Sw-1, 64) f32 (shared by the heads; tables of another
 *   length are interpolated by the caller) are passed concatenated [rel_pos_h ; rel_pos_w] and split as rcat_hi = bf16(R),
 *   rcat_lo = bf16(R - hi): the kernels multiply queries with both parts on the tensor cores (fp32-class accuracy).
 * vdr_relpos_tables: the decomposed terms of add_decomposed_rel_pos for every query of BW images/windows of Sh x Sw tokens:
 *   rel[((bw*heads + h)*N + q)*(Sh+Sw) + j] = out_scale * q . rel_pos_h[qh - j + Sh - 1] (j < Sh) or
 *   out_scale * q . rel_pos_w[qw - (j - Sh) + Sw - 1]; q = the UNSCALED query vector read from qkv.
 * vdr_flash_attn_relpos_fwd: the tcgen05 flash kernel (vdr_flash_attn_fwd) with that bias added to the scores, for token grids
 *   of Sh x 64 with Sh % 4 == 0 (the 64 x 64 grid of SAM's global-attention blocks): rel_log2 = vdr_relpos_tables(..., Sw = 64,
 *   out_scale = log2(e)).  out = softmax(q k^T * scale + rel_h[q, kh] + rel_w[q, kw]) v.
 * vdr_attn_relpos_fwd: out = softmax(q k^T * scale + rel_h[q, kh] + rel_w[q, kw]) v per (image/window, head); qkv
 *   (BW*N, >= 3*heads*64) as the qkv GEMM writes it, N = Sh*Sw < 65536; out (BW*N, heads*64) bf16.  Any extent (the 14 x 14
 *   windows); mma.sync kernel that builds its bias terms on chip from rcat_hi / rcat_lo in its prologue (no table in HBM).
 * vdr_im2col3x3_tokens: A[(b,y,x), (ky*3+kx)*C + c] = X[(b, y+ky-1, x+kx-1), c], zero padded: the A operand of the neck's
 *   3x3 convolution as a GEMM against the weight permuted to (out, ky, kx, in). */
int vdr_window_rows(const void* src_bf16, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int B, int H, int W, int ws, int d,
                    int to_windows, vdr_stream_t stream);
int vdr_relpos_tables(const void* qkv_bf16, int64_t ld_qkv, const void* rcat_hi_bf16, const void* rcat_lo_bf16, float* rel, int BW,
                      int Sh, int Sw, int heads, float out_scale, vdr_stream_t stream);
int vdr_flash_attn_relpos_fwd(const void* qkv, int64_t ld_qkv, const float* rel_log2, void* out, int64_t ld_out, int B, int Sh,
                              int heads, float scale, vdr_stream_t stream);
/* vdr_flash_attn_relpos_fwd without the table: the kernel computes its rows' bias terms itself (Q [R_hi ; R_lo]^T on the tensor cores
 * into TMEM before the first key block).  Token grids of Sh x 64, Sh % 4 == 0, Sh <= 64; rcat_* = the split [rel_pos_h ; rel_pos_w]
 * tables ((2*Sh-1) + 127 rows of 64).  Replaces vdr_relpos_tables + vdr_flash_attn_relpos_fwd for SAM's global-attention blocks. */
int vdr_flash_attn_relpos_fused_fwd(const void* qkv, int64_t ld_qkv, const void* rcat_hi_bf16, const void* rcat_lo_bf16, void* out,
                                    int64_t ld_out, int B, int Sh, int heads, float scale, vdr_stream_t stream);
int vdr_attn_relpos_fwd(const void* qkv_bf16, int64_t ld_qkv, const void* rcat_hi_bf16, const void* rcat_lo_bf16, void* out_bf16,
                        int64_t ld_out, int BW, int Sh, int Sw, int heads, float scale, vdr_stream_t stream);
/* Windowed attention straight on the un-partitioned token rows of B images of gh x gw tokens (window_partition, the attention
 * and window_unpartition of a windowed block in one launch): window (wy, wx) reads its ws x ws tokens in place, pad tokens
 * (positions beyond the image: their normalised input is zero, so q = k = v = the qkv bias) are synthesised from qkv_bias
 * ((3*heads*64) f32, the UNFOLDED bias of the qkv Linear), outputs of pad queries are dropped.  ws*ws <= 208. */
int vdr_attn_relpos_windows_fwd(const void* qkv_bf16, int64_t ld_qkv, const float* qkv_bias, const void* rcat_hi_bf16,
                                const void* rcat_lo_bf16, void* out_bf16, int64_t ld_out, int B, int gh, int gw, int ws, int heads,
                                float scale, vdr_stream_t stream);
int vdr_im2col3x3_tokens(const void* X_bf16, int64_t ldx, void* A_bf16, int64_t lda, int B, int H, int W, int C,
                         vdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SAM's image encoder as one call (the reference's DEFAULT backbone: load_medsam + `model.image_encoder(img_tensor)`,
 * src/tfds_dense_descriptor.py:91-107,123; segment_anything ImageEncoderViT: patch_embed, pos_embed, blocks[i].{norm1, attn.qkv,
 * attn.proj, attn.rel_pos_h / rel_pos_w, norm2, mlp.lin1, mlp.lin2}, neck): every kernel enqueued on `stream`, no host
 * synchronisation, no allocation.
 *   blocks      HOST array of `depth` entries.  The LayerNorms are folded into the GEMMs that consume them: qkv_wf / qkv_bf / qkv_cs
 *               and fc1_wf / fc1_bf / fc1_cs from vdr_fold_layernorm are REQUIRED; qkv_b is the unfolded qkv bias (what a window's
 *               pad token projects to).  rel_hi / rel_lo = the split [rel_pos_h ; rel_pos_w] tables of the block's extent
 *               (window x window, or the whole token grid when window == 0 = a global-attention block).
 *   images_bf16 (B, H, W) gray slices (gray2rgb :41 is folded into the channel-summed patch weights pe_w_gray when the geometry
 *               tiles into TMA im2col boxes, else the slices are expanded in the workspace), or NULL with
 *   im2col_bf16 (B*N, 3*patch^2) the materialised patch matrix (k = (c, iy, ix)) of RGB pictures (vdr_im2col_patches)
 *   descriptors (B*N, out_chans) f32 with row pitch ld_out, N = (H/patch)*(W/patch) tokens in (row, col) order: the
 *               (64, 64, 256) map get_dense_descriptor returns per slice (:124-126), for B slices
 *   workspace   >= vdr_sam_forward_workspace_bytes(weights, B), 256-byte aligned
 */
typedef struct {
  const void* qkv_wf; const float* qkv_bf; const float* qkv_cs;   /* norm1 folded into attn.qkv: (3*dim, dim) bf16, bias', column sums */
  const float* qkv_b;                                              /* the unfolded qkv bias (3*dim) */
  const void* proj_w; const float* proj_b;                         /* (dim, dim) */
  const void* fc1_wf; const float* fc1_bf; const float* fc1_cs;   /* norm2 folded into mlp.lin1: (4*dim, dim) */
  const void* fc2_w;  const float* fc2_b;                          /* (dim, 4*dim) */
  const void* rel_hi; const void* rel_lo;                          /* bf16 hi / lo parts of [rel_pos_h ; rel_pos_w] */
  int window;                                                      /* 0 = global attention */
} vdr_sam_block;

typedef struct {
  int dim, depth, heads, patch, H, W, out_chans;
  float eps;                                  /* LayerNorm epsilon (<= 0: 1e-6) */
  const void* pe_w; int64_t pe_ldw;           /* (dim, 3*patch^2) bf16, k = (c, iy, ix) */
  const void* pe_w_gray; int64_t pe_gray_ldw; /* optional: channel-summed (dim, patch^2) for gray slices (vdr_patch_embed_gemm_gray) */
  const float* pe_b;
  const float* pos;                           /* (N, dim) */
  const void* neck0;                          /* (out_chans, dim) bf16: 1x1 conv */
  const float *neck1_w, *neck1_b;             /* LayerNorm2d */
  const void* neck2;                          /* (out_chans, 9*out_chans) bf16: 3x3 conv weight permuted to (out, ky, kx, in) */
  const float *neck3_w, *neck3_b;
  const vdr_sam_block* blocks;
} vdr_sam_weights;

size_t vdr_sam_forward_workspace_bytes(const vdr_sam_weights* weights, int B);
int vdr_sam_forward(const vdr_sam_weights* weights, const void* images_bf16, const void* im2col_bf16, int B, float* descriptors,
                    int64_t ld_out, void* workspace, size_t workspace_bytes, vdr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VDR_H_ */
