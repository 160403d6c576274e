// vdr_vit_forward: the whole backbone forward (patch embedding -> transformer blocks -> final LayerNorm) as ONE C call
// that enqueues every kernel on the caller's stream.  Replaces the `model.image_encoder(img_tensor)` call of the
// reference (src/tfds_dense_descriptor.py:123) for a batch of slices; no host synchronisation, no allocation.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct VitPlan {
  int N, Np, d, K, ldk;
  bool tma_patch_embed;
  bool folded;   // LayerNorms folded into the qkv / fc1 GEMMs (vdr_vit_block::*_wf present)
  size_t off_x, off_y, off_qkv, off_h, off_a, off_st, total;
};

int make_plan(const vdr_vit_weights* w, int B, VitPlan* pl) {
  using namespace vdr;
  VDR_CHECK_ARG(w != nullptr && B > 0, VDR_EINVAL, "vdr_vit_forward: null weights / non-positive batch");
  VDR_CHECK_ARG(w->dim > 0 && w->dim % 64 == 0 && w->heads * 64 == w->dim && w->depth > 0 && w->patch > 0, VDR_EINVAL,
                "vdr_vit_forward: dim (%d) must be heads (%d) x 64, depth (%d) and patch (%d) positive", w->dim, w->heads, w->depth, w->patch);
  VDR_CHECK_ARG(w->H > 0 && w->W > 0 && w->H % w->patch == 0 && w->W % w->patch == 0, VDR_EINVAL,
                "vdr_vit_forward: image size %dx%d is not a multiple of the patch size %d", w->H, w->W, w->patch);
  pl->d = w->dim;
  pl->Np = (w->H / w->patch) * (w->W / w->patch);
  pl->N = pl->Np + 1;
  pl->K = 3 * w->patch * w->patch;
  pl->ldk = (pl->K + 7) & ~7;
  pl->tma_patch_embed = vdr_patch_embed_supported(w->H, w->W, w->patch) != 0;
  const size_t rows = static_cast<size_t>(B) * pl->N;
  size_t off = 0;
  pl->off_x = off;   off += align256(rows * pl->d * 2);
  pl->off_y = off;   off += align256(rows * pl->d * 2);
  pl->off_qkv = off; off += align256(rows * pl->d * 3 * 2);
  pl->off_h = off;   off += align256(rows * pl->d * 4 * 2);
  pl->off_a = off;
  if (!pl->tma_patch_embed) off += align256(static_cast<size_t>(B) * pl->Np * pl->ldk * 2);
  int nf = 0;
  if (w->blocks != nullptr)
    for (int l = 0; l < w->depth; ++l) {
      const vdr_vit_block& b = w->blocks[l];
      const int have = (b.qkv_wf != nullptr) + (b.qkv_bf != nullptr) + (b.qkv_cs != nullptr) + (b.fc1_wf != nullptr) + (b.fc1_bf != nullptr) + (b.fc1_cs != nullptr);
      VDR_CHECK_ARG(have == 0 || have == 6, VDR_EINVAL, "vdr_vit_forward: block %d has %d of the 6 folded-LayerNorm pointers (all or none)", l, have);
      nf += have == 6;
    }
  VDR_CHECK_ARG(nf == 0 || nf == w->depth, VDR_EINVAL, "vdr_vit_forward: %d of %d blocks carry folded LayerNorm weights (all or none)", nf, w->depth);
  pl->folded = nf > 0;
  pl->off_st = off;
  if (pl->folded) off += align256(rows * (pl->d / 64) * 8);      // row statistics: d/64 slots of (sum, sumsq) per row
  pl->total = off;
  return VDR_OK;
}

}  // namespace

extern "C" size_t vdr_vit_forward_workspace_bytes(const vdr_vit_weights* w, int B) {
  VitPlan pl;
  if (make_plan(w, B, &pl) != VDR_OK) return 0;
  return pl.total;
}

extern "C" int vdr_vit_forward(const vdr_vit_weights* w, const void* images_bf16, int B, int C, float* tokens_out,
                               int64_t ld_out, void* workspace, size_t workspace_bytes, vdr_stream_t stream) {
  using namespace vdr;
  VitPlan pl;
  int rc = make_plan(w, B, &pl);
  if (rc != VDR_OK) return rc;
  VDR_CHECK_ARG(images_bf16 && tokens_out && workspace && w->blocks, VDR_EINVAL, "vdr_vit_forward: null pointer");
  VDR_CHECK_ARG(C == 1 || C == 3, VDR_EINVAL, "vdr_vit_forward: C must be 1 (gray) or 3");
  VDR_CHECK_ARG(workspace_bytes >= pl.total, VDR_EWORKSPACE, "vdr_vit_forward: workspace too small (%zu < %zu)", workspace_bytes, pl.total);
  VDR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VDR_EALIGN, "vdr_vit_forward: workspace must be 256-byte aligned");
  VDR_CHECK_ARG(ld_out >= pl.d && ld_out % 4 == 0, VDR_EALIGN, "vdr_vit_forward: ld_out (%lld) must be >= dim and a multiple of 4", (long long)ld_out);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void *X = ws + pl.off_x, *Y = ws + pl.off_y, *QKV = ws + pl.off_qkv, *Hb = ws + pl.off_h, *A = ws + pl.off_a;
  const int d = pl.d, N = pl.N, Np = pl.Np, M = B * N;
  const float eps = w->eps > 0.f ? w->eps : 1e-6f;

  // ---- patch embedding (+ bias + position embedding), rows written behind each image's CLS row
  if (pl.tma_patch_embed) {
    if (C == 1 && w->pe_w_gray)   // gray slices: channel-summed weights, K = patch^2
      rc = vdr_patch_embed_gemm_gray(images_bf16, B, w->H, w->W, w->patch, w->pe_w_gray, w->pe_gray_ldw, w->pe_b, w->pos, X, d, d, 1, stream);
    else
      rc = vdr_patch_embed_gemm(images_bf16, B, C, w->H, w->W, w->patch, w->pe_w, w->pe_ldw, w->pe_b, w->pos, X, d, d, stream);
    if (rc != VDR_OK) return rc;
  } else {
    VDR_CHECK_ARG(C == 1, VDR_EINVAL, "vdr_vit_forward: RGB input needs a geometry the TMA im2col view supports (or vdr_im2col_patches + vdr_gemm)");
    rc = vdr_im2col_gray_bf16(images_bf16, B, w->H, w->W, w->patch, A, stream);
    if (rc != VDR_OK) return rc;
    vdr_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.lda = pl.ldk; g.W = w->pe_w; g.ldw = w->pe_ldw; g.bias = w->pe_b;
    g.R = w->pos; g.ldr = d; g.r_dtype = VDR_DTYPE_F32;
    g.C = X; g.ldc = d; g.c_dtype = VDR_DTYPE_BF16;
    g.M = B * Np; g.N = d; g.K = pl.K; g.epilogue = VDR_EPI_BIAS_RESIDUAL;
    g.out_group = Np; g.out_group_stride = N; g.out_offset = 1; g.res_mod = Np; g.res_offset = 1;
    rc = vdr_gemm(&g, stream);
    if (rc != VDR_OK) return rc;
  }
  rc = vdr_write_cls_rows(w->cls, w->pos, X, B, N, d, stream);
  if (rc != VDR_OK) return rc;

  // ---- transformer blocks (pre-norm): x += proj(attn(LN1 x)); x += fc2(gelu(fc1(LN2 x)))
  float* ST = reinterpret_cast<float*>(ws + pl.off_st);
  // ln_slots > 0: A is the raw residual stream, normalised in the epilogue from ST; stats = true: emit the statistics of C into ST
  auto gemm = [&](const void* a, int64_t lda, const void* wt, const float* bias, int n, int k, int epi, const void* res, void* c,
                  int ln_slots = 0, const float* colsum = nullptr, bool stats = false) {
    vdr_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = a; g.lda = lda; g.W = wt; g.ldw = k; g.bias = bias;
    g.R = res; g.ldr = d; g.r_dtype = VDR_DTYPE_BF16;
    g.C = c; g.ldc = n; g.c_dtype = VDR_DTYPE_BF16;
    g.M = M; g.N = n; g.K = k; g.epilogue = epi;
    if (ln_slots > 0) { g.ln_stats = ST; g.ln_slots = ln_slots; g.ln_eps = eps; g.ln_colsum = colsum; }
    if (stats) g.stats_out = ST;
    return vdr_gemm(&g, stream);
  };
  const float scale = 1.0f / sqrtf(64.f);
  if (pl.folded) {
    // No LayerNorm kernel inside the blocks: the residual GEMMs (in place on X) leave per-row (sum, sumsq) in ST, the qkv / fc1
    // GEMMs read X itself and apply rstd * (acc - mean * colsum) + bias' in their epilogue.  Layer 0 takes its statistics from a
    // read-only pass over the patch-embedding output.
    const int slots = d / 64;
    if ((rc = vdr_row_stats(X, d, M, d, ST, stream)) != VDR_OK) return rc;
    for (int l = 0; l < w->depth; ++l) {
      const vdr_vit_block& b = w->blocks[l];
      if ((rc = gemm(X, d, b.qkv_wf, b.qkv_bf, 3 * d, d, VDR_EPI_BIAS, nullptr, QKV, l == 0 ? 1 : slots, b.qkv_cs)) != VDR_OK) return rc;
      if ((rc = vdr_flash_attn_fwd(QKV, 3 * d, Y, d, nullptr, B, N, w->heads, scale, nullptr, stream)) != VDR_OK) return rc;
      if ((rc = gemm(Y, d, b.proj_w, b.proj_b, d, d, VDR_EPI_BIAS_RESIDUAL, X, X, 0, nullptr, true)) != VDR_OK) return rc;
      if ((rc = gemm(X, d, b.fc1_wf, b.fc1_bf, 4 * d, d, VDR_EPI_BIAS_GELU, nullptr, Hb, slots, b.fc1_cs)) != VDR_OK) return rc;
      if ((rc = gemm(Hb, 4 * d, b.fc2_w, b.fc2_b, d, 4 * d, VDR_EPI_BIAS_RESIDUAL, X, X, 0, nullptr, l + 1 < w->depth)) != VDR_OK) return rc;
    }
    return vdr_layernorm_fwd(X, d, w->norm_w, w->norm_b, tokens_out, ld_out, VDR_DTYPE_F32, nullptr, nullptr, M, d, eps, stream);
  }
  for (int l = 0; l < w->depth; ++l) {
    const vdr_vit_block& b = w->blocks[l];
    if ((rc = vdr_layernorm_fwd(X, d, b.n1w, b.n1b, Y, d, VDR_DTYPE_BF16, nullptr, nullptr, M, d, eps, stream)) != VDR_OK) return rc;
    if ((rc = gemm(Y, d, b.qkv_w, b.qkv_b, 3 * d, d, VDR_EPI_BIAS, nullptr, QKV)) != VDR_OK) return rc;
    if ((rc = vdr_flash_attn_fwd(QKV, 3 * d, Y, d, nullptr, B, N, w->heads, scale, nullptr, stream)) != VDR_OK) return rc;
    if ((rc = gemm(Y, d, b.proj_w, b.proj_b, d, d, VDR_EPI_BIAS_RESIDUAL, X, X)) != VDR_OK) return rc;
    if ((rc = vdr_layernorm_fwd(X, d, b.n2w, b.n2b, Y, d, VDR_DTYPE_BF16, nullptr, nullptr, M, d, eps, stream)) != VDR_OK) return rc;
    if ((rc = gemm(Y, d, b.fc1_w, b.fc1_b, 4 * d, d, VDR_EPI_BIAS_GELU, nullptr, Hb)) != VDR_OK) return rc;
    if ((rc = gemm(Hb, 4 * d, b.fc2_w, b.fc2_b, d, 4 * d, VDR_EPI_BIAS_RESIDUAL, X, X)) != VDR_OK) return rc;
  }
  return vdr_layernorm_fwd(X, d, w->norm_w, w->norm_b, tokens_out, ld_out, VDR_DTYPE_F32, nullptr, nullptr, M, d, eps, stream);
}
