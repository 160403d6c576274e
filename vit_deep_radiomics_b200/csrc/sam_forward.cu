// vdr_sam_forward: SAM's ViT image encoder (the reference's default backbone, `model.image_encoder(img_tensor)` of load_medsam,
// src/tfds_dense_descriptor.py:91-107,123) as ONE C call that enqueues every kernel on the caller's stream:
//   patch embedding (+ absolute position embedding) -> depth x (windowed or global attention with the decomposed relative-position
//   bias, MLP) -> neck (1x1 conv, LayerNorm2d, 3x3 conv, LayerNorm2d).
// No LayerNorm kernel runs inside the blocks: norm1 / norm2 are folded into the qkv / lin1 GEMMs (vdr_fold_layernorm), the residual
// GEMMs leave the row statistics.  Windowed blocks read their 14 x 14 windows in place (vdr_attn_relpos_windows_fwd), global blocks
// of Sh x 64 token grids use the flash kernel that computes its own bias terms (vdr_flash_attn_relpos_fused_fwd).
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct SamPlan {
  int gh, gw, N, d, oc, K;
  bool tma_patch_embed;
  size_t off_x, off_y, off_qkv, off_h, off_a, off_st, total;
  size_t off_n0, off_n1, off_na;      // neck buffers: aliases of QKV / H (free once the blocks are done) where they fit
};

int make_plan(const vdr_sam_weights* w, int B, bool need_im2col, SamPlan* pl) {
  using namespace vdr;
  VDR_CHECK_ARG(w != nullptr && B > 0, VDR_EINVAL, "vdr_sam_forward: null weights / non-positive batch");
  VDR_CHECK_ARG(w->dim > 0 && w->dim % 64 == 0 && w->heads * 64 == w->dim && w->depth > 0 && w->patch > 0 && w->out_chans > 0 && w->out_chans % 8 == 0,
                VDR_EINVAL, "vdr_sam_forward: dim (%d) must be heads (%d) x 64; depth (%d), patch (%d) positive; out_chans (%d) a multiple of 8", w->dim,
                w->heads, w->depth, w->patch, w->out_chans);
  VDR_CHECK_ARG(w->H > 0 && w->W > 0 && w->H % w->patch == 0 && w->W % w->patch == 0, VDR_EINVAL,
                "vdr_sam_forward: image size %dx%d is not a multiple of the patch size %d", w->H, w->W, w->patch);
  pl->d = w->dim;
  pl->oc = w->out_chans;
  pl->gh = w->H / w->patch;
  pl->gw = w->W / w->patch;
  pl->N = pl->gh * pl->gw;
  pl->K = 3 * w->patch * w->patch;
  pl->tma_patch_embed = vdr_patch_embed_supported(w->H, w->W, w->patch) != 0 && w->pe_w_gray != nullptr;
  const size_t rows = static_cast<size_t>(B) * pl->N;
  size_t off = 0;
  pl->off_x = off;   off += align256(rows * pl->d * 2);
  pl->off_y = off;   off += align256(rows * pl->d * 2);
  pl->off_qkv = off; off += align256(rows * pl->d * 3 * 2);
  pl->off_h = off;   off += align256(rows * pl->d * 4 * 2);
  pl->off_st = off;  off += align256(rows * (pl->d / 64) * 8);       // row statistics: d/64 slots of (sum, sumsq) per row
  pl->off_a = off;
  if (need_im2col) off += align256(rows * pl->K * 2);                // gray slices of a geometry the TMA im2col view cannot address
  // the neck runs when the blocks are done: its buffers share the QKV / H storage where they fit (SAM ViT-B: 256 of 768 channels)
  if (2 * align256(rows * pl->oc * 2) <= align256(rows * pl->d * 3 * 2)) {
    pl->off_n0 = pl->off_qkv;
    pl->off_n1 = pl->off_qkv + align256(rows * pl->oc * 2);
  } else {
    pl->off_n0 = off; off += align256(rows * pl->oc * 2);
    pl->off_n1 = off; off += align256(rows * pl->oc * 2);
  }
  if (9 * pl->oc <= 4 * pl->d) {
    pl->off_na = pl->off_h;
  } else {
    pl->off_na = off; off += align256(rows * 9 * pl->oc * 2);
  }
  pl->total = off;
  return VDR_OK;
}

}  // namespace

extern "C" size_t vdr_sam_forward_workspace_bytes(const vdr_sam_weights* w, int B) {
  SamPlan pl;
  if (w == nullptr) return 0;
  const bool tma = vdr_patch_embed_supported(w->H, w->W, w->patch) != 0 && w->pe_w_gray != nullptr;
  if (make_plan(w, B, !tma, &pl) != VDR_OK) return 0;                 // (sized for gray slices; a caller-made im2col matrix needs no more)
  return pl.total;
}

extern "C" int vdr_sam_forward(const vdr_sam_weights* w, const void* images_bf16, const void* im2col_bf16, int B, float* descriptors,
                               int64_t ld_out, void* workspace, size_t workspace_bytes, vdr_stream_t stream) {
  using namespace vdr;
  VDR_CHECK_ARG(w != nullptr, VDR_EINVAL, "vdr_sam_forward: null weights");
  VDR_CHECK_ARG((images_bf16 != nullptr) != (im2col_bf16 != nullptr), VDR_EINVAL,
                "vdr_sam_forward: pass either gray slices (images_bf16) or a materialised im2col matrix (im2col_bf16)");
  SamPlan pl;
  const bool tma = images_bf16 && vdr_patch_embed_supported(w->H, w->W, w->patch) != 0 && w->pe_w_gray != nullptr;
  int rc = make_plan(w, B, images_bf16 && !tma, &pl);
  if (rc != VDR_OK) return rc;
  VDR_CHECK_ARG(descriptors && workspace && w->blocks && w->pe_w && w->pe_b && w->pos && w->neck0 && w->neck2, VDR_EINVAL, "vdr_sam_forward: null pointer");
  VDR_CHECK_ARG(workspace_bytes >= pl.total, VDR_EWORKSPACE, "vdr_sam_forward: workspace too small (%zu < %zu)", workspace_bytes, pl.total);
  VDR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, VDR_EALIGN, "vdr_sam_forward: workspace must be 256-byte aligned");
  VDR_CHECK_ARG(ld_out >= pl.oc && ld_out % 4 == 0, VDR_EALIGN, "vdr_sam_forward: ld_out (%lld) must be >= out_chans and a multiple of 4", (long long)ld_out);
  for (int l = 0; l < w->depth; ++l) {
    const vdr_sam_block& b = w->blocks[l];
    VDR_CHECK_ARG(b.qkv_wf && b.qkv_bf && b.qkv_cs && b.fc1_wf && b.fc1_bf && b.fc1_cs && b.qkv_b && b.proj_w && b.proj_b && b.fc2_w && b.fc2_b &&
                  b.rel_hi && b.rel_lo, VDR_EINVAL, "vdr_sam_forward: block %d lacks folded-LayerNorm weights (vdr_fold_layernorm), biases or rel-pos tables", l);
    VDR_CHECK_ARG(b.window >= 0 && b.window * b.window <= 208, VDR_EINVAL, "vdr_sam_forward: block %d: windows of %d x %d tokens are not read in place (<= 208 tokens)",
                  l, b.window, b.window);
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void *X = ws + pl.off_x, *Y = ws + pl.off_y, *QKV = ws + pl.off_qkv, *Hb = ws + pl.off_h;
  float* ST = reinterpret_cast<float*>(ws + pl.off_st);
  const int d = pl.d, N = pl.N, oc = pl.oc, M = B * N, gh = pl.gh, gw = pl.gw;
  const float eps = w->eps > 0.f ? w->eps : 1e-6f;

  // ---- patch embedding (Conv2d p x p, stride p) + absolute position embedding in the GEMM epilogue
  if (tma) {
    rc = vdr_patch_embed_gemm_gray(images_bf16, B, w->H, w->W, w->patch, w->pe_w_gray, w->pe_gray_ldw, w->pe_b, w->pos, X, d, d, 0, stream);
    if (rc != VDR_OK) return rc;
  } else {
    const void* A = im2col_bf16;
    if (images_bf16) {
      void* Aw = ws + pl.off_a;
      if ((rc = vdr_im2col_gray_bf16(images_bf16, B, w->H, w->W, w->patch, Aw, stream)) != VDR_OK) return rc;
      A = Aw;
    }
    vdr_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.lda = pl.K; g.W = w->pe_w; g.ldw = w->pe_ldw; g.bias = w->pe_b;
    g.R = w->pos; g.ldr = d; g.r_dtype = VDR_DTYPE_F32;
    g.C = X; g.ldc = d; g.c_dtype = VDR_DTYPE_BF16;
    g.M = M; g.N = d; g.K = pl.K; g.epilogue = VDR_EPI_BIAS_RESIDUAL;
    g.res_mod = N; g.res_offset = 0;
    if ((rc = vdr_gemm(&g, stream)) != VDR_OK) return rc;
  }

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
  const int slots = d / 64;
  const bool fused_global = gw == 64 && gh % 4 == 0 && gh <= 64;
  if ((rc = vdr_row_stats(X, d, M, d, ST, stream)) != VDR_OK) return rc;
  for (int l = 0; l < w->depth; ++l) {
    const vdr_sam_block& b = w->blocks[l];
    if ((rc = gemm(X, d, b.qkv_wf, b.qkv_bf, 3 * d, d, VDR_EPI_BIAS, nullptr, QKV, l == 0 ? 1 : slots, b.qkv_cs)) != VDR_OK) return rc;
    if (b.window > 0)
      rc = vdr_attn_relpos_windows_fwd(QKV, 3 * d, b.qkv_b, b.rel_hi, b.rel_lo, Y, d, B, gh, gw, b.window, w->heads, scale, stream);
    else if (fused_global)
      rc = vdr_flash_attn_relpos_fused_fwd(QKV, 3 * d, b.rel_hi, b.rel_lo, Y, d, B, gh, w->heads, scale, stream);
    else
      rc = vdr_attn_relpos_fwd(QKV, 3 * d, b.rel_hi, b.rel_lo, Y, d, B, gh, gw, w->heads, scale, stream);
    if (rc != VDR_OK) return rc;
    if ((rc = gemm(Y, d, b.proj_w, b.proj_b, d, d, VDR_EPI_BIAS_RESIDUAL, X, X, 0, nullptr, true)) != VDR_OK) return rc;
    if ((rc = gemm(X, d, b.fc1_wf, b.fc1_bf, 4 * d, d, VDR_EPI_BIAS_GELU, nullptr, Hb, slots, b.fc1_cs)) != VDR_OK) return rc;
    if ((rc = gemm(Hb, 4 * d, b.fc2_w, b.fc2_b, d, 4 * d, VDR_EPI_BIAS_RESIDUAL, X, X, 0, nullptr, l + 1 < w->depth)) != VDR_OK) return rc;
  }

  // ---- neck: 1x1 conv (a GEMM), LayerNorm2d = LayerNorm over the channels of each token, 3x3 conv as im2col + GEMM, LayerNorm2d
  void *N0 = ws + pl.off_n0, *N1 = ws + pl.off_n1, *NA = ws + pl.off_na;
  auto neck_gemm = [&](const void* a, int k, const void* wt, void* c) {
    vdr_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = a; g.lda = k; g.W = wt; g.ldw = k;
    g.C = c; g.ldc = oc; g.c_dtype = VDR_DTYPE_BF16;
    g.M = M; g.N = oc; g.K = k; g.epilogue = VDR_EPI_BIAS;
    return vdr_gemm(&g, stream);
  };
  if ((rc = neck_gemm(X, d, w->neck0, N0)) != VDR_OK) return rc;
  if ((rc = vdr_layernorm_fwd(N0, oc, w->neck1_w, w->neck1_b, N1, oc, VDR_DTYPE_BF16, nullptr, nullptr, M, oc, eps, stream)) != VDR_OK) return rc;
  if ((rc = vdr_im2col3x3_tokens(N1, oc, NA, 9 * oc, B, gh, gw, oc, stream)) != VDR_OK) return rc;
  if ((rc = neck_gemm(NA, 9 * oc, w->neck2, N0)) != VDR_OK) return rc;
  return vdr_layernorm_fwd(N0, oc, w->neck3_w, w->neck3_b, descriptors, ld_out, VDR_DTYPE_F32, nullptr, nullptr, M, oc, eps, stream);
}
