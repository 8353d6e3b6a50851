// C ABI of the wav2vec-S encoder forward (include/w2vs.h): argument validation, weight packing,
// and the full-utterance forward schedule.  Host code only -- every kernel lives in k_*.cu.
//
// Forward schedule (w2vs_encode), all on the caller's stream, no host synchronisation:
//   conv0 (+norm+GELU)                                     wav2vec2.py:773-781 (block 0)
//   conv i>=1 as implicit GEMM (+bias) [+LayerNorm] +GELU   wav2vec2.py:773-781 (blocks 1..6)
//   feature LayerNorm -> post_extract_proj GEMM            wav2vec2.py:556-568
//   prep_masks (frame mask, positions, extended key mask)   wav2vec2.py:560-565, wav2vec_S.py:470-476
//   [positional conv]  embed_tokens (zero pads, +pos, [LN], pad to T', append look-ahead copies)
//   per layer: [LN] QKV GEMM, block-masked attention, out_proj GEMM(+residual), [LN],
//              fc1 GEMM(+GELU), fc2 GEMM(+residual), [LN]  wav2vec2.py:921-978
//   finalize (drop copies/padding, [final LN], BTD | TBD)   wav2vec_S.py:425-440, rain :314-330
#include <string.h>
#include "common.cuh"
#include "kernels.h"
#include "layout.h"

using namespace w2vs;

namespace {

inline bool conv_has_norm(const w2vs_config* cfg, int i) {
  return (cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM && i < cfg->layer_norm_num) ||
         (cfg->extractor_mode == W2VS_EXTRACTOR_DEFAULT && i == 0);
}
inline int c_last(const w2vs_config* cfg) { return cfg->conv_dim[cfg->n_conv - 1]; }

template <typename T> inline T* at(void* base, size_t off) {
  return off == kNone ? nullptr : reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(base) + off);
}
template <typename T> inline const T* at(const void* base, size_t off) {
  return off == kNone ? nullptr : reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(base) + off);
}

}  // namespace

extern "C" {

const char* w2vs_status_string(int32_t status) {
  switch (status) {
    case W2VS_OK: return "ok";
    case W2VS_INVALID_VALUE: return "invalid value";
    case W2VS_UNSUPPORTED: return "unsupported configuration";
    case W2VS_WORKSPACE_TOO_SMALL: return "workspace too small";
    case W2VS_CUDA_ERROR: return "CUDA error";
  }
  return "unknown status";
}

const char* w2vs_last_error(void) { return g_last_error; }

int64_t w2vs_launch_count(int32_t reset) {
  const int64_t v = g_launch_count;
  if (reset) g_launch_count = 0;
  return v;
}

int32_t w2vs_num_ref_tensors(const w2vs_config* cfg) {
  if (validate_config(cfg) != W2VS_OK) return -1;
  int n = 0;
  for (int i = 0; i < cfg->n_conv; ++i) n += 1 + (cfg->conv_bias ? 1 : 0) + (conv_has_norm(cfg, i) ? 2 : 0);
  n += 2;
  if (c_last(cfg) != cfg->embed_dim) n += 2;
  n += cfg->pos_type == W2VS_POS_SIN ? 1 : 3;
  n += 16 * cfg->layers;
  n += 2;
  return n;
}

w2vs_status_t w2vs_packed_weights_size(const w2vs_config* cfg, size_t* bytes) {
  W2VS_TRY(validate_config(cfg));
  W2VS_REQUIRE(bytes != nullptr, "bytes is NULL");
  WeightLayout wl;
  make_weight_layout(cfg, &wl);
  *bytes = wl.total;
  return W2VS_OK;
}

w2vs_status_t w2vs_weights_pack(const w2vs_config* cfg, const void* const* d_ref, int32_t n_tensors,
                                void* d_packed, size_t packed_bytes, void* stream) {
  W2VS_TRY(validate_config(cfg));
  W2VS_REQUIRE(d_ref != nullptr && d_packed != nullptr, "NULL pointer");
  W2VS_REQUIRE(n_tensors == w2vs_num_ref_tensors(cfg), "wrong number of reference tensors");
  for (int i = 0; i < n_tensors; ++i) W2VS_REQUIRE(d_ref[i] != nullptr, "NULL reference tensor");
  WeightLayout wl;
  make_weight_layout(cfg, &wl);
  if (packed_bytes < wl.total) { set_error("packed buffer too small: %zu < %zu", packed_bytes, wl.total); return W2VS_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  const int adt = cfg->dtype;
  int t = 0;
  auto next = [&]() { return reinterpret_cast<const float*>(d_ref[t++]); };
  auto copy_f32 = [&](size_t off, int64_t n) { return launch_pack_copy(next(), at<float>(d_packed, off), W2VS_F32, n, st); };
  auto copy_act = [&](size_t off, int64_t n) { return launch_pack_copy(next(), at<void>(d_packed, off), adt, n, st); };

  int c_in = 1;
  for (int i = 0; i < cfg->n_conv; ++i) {
    const int c = cfg->conv_dim[i], k = cfg->conv_kernel[i];
    if (i == 0) W2VS_TRY(copy_f32(wl.conv[i].w, (int64_t)c * k));
    else W2VS_TRY(launch_pack_conv(next(), at<void>(d_packed, wl.conv[i].w), adt, c, c_in, k, st));
    if (cfg->conv_bias) W2VS_TRY(copy_f32(wl.conv[i].bias, c));
    if (conv_has_norm(cfg, i)) {
      W2VS_TRY(copy_f32(wl.conv[i].norm_w, c));
      W2VS_TRY(copy_f32(wl.conv[i].norm_b, c));
    }
    c_in = c;
  }
  const int D = cfg->embed_dim, F = cfg->ffn_dim;
  W2VS_TRY(copy_f32(wl.feat_ln_w, c_in));
  W2VS_TRY(copy_f32(wl.feat_ln_b, c_in));
  if (c_in != D) {
    W2VS_TRY(copy_act(wl.proj_w, (int64_t)D * c_in));
    W2VS_TRY(copy_f32(wl.proj_b, D));
  }
  if (cfg->pos_type == W2VS_POS_SIN) {
    W2VS_TRY(copy_f32(wl.sin_table, (int64_t)cfg->sin_rows * D));
  } else {
    W2VS_TRY(copy_f32(wl.posconv_b, D));
    const float* g = next();
    const float* v = next();
    W2VS_TRY(launch_pack_posconv(g, v, at<float>(d_packed, wl.posconv_w), D, cfg->conv_pos_groups, cfg->conv_pos, st));
    if (posconv_tc(cfg))
      W2VS_TRY(launch_pack_posconv_tc(at<float>(d_packed, wl.posconv_w), at<void>(d_packed, wl.posconv_wg), D,
                                      cfg->conv_pos_groups, cfg->conv_pos, posconv_dgp(cfg), st));
  }
  const size_t as = act_size(cfg);
  for (int n = 0; n < cfg->layers; ++n) {
    const LayerW l = layer_at(wl, n);
    for (int p = 0; p < 3; ++p) {  // q, k, v
      W2VS_TRY(copy_act(l.wqkv + (size_t)p * D * D * as, (int64_t)D * D));
      W2VS_TRY(copy_f32(l.bqkv + (size_t)p * D * 4, D));
    }
    W2VS_TRY(copy_act(l.wo, (int64_t)D * D));
    W2VS_TRY(copy_f32(l.bo, D));
    W2VS_TRY(copy_f32(l.ln1_w, D));
    W2VS_TRY(copy_f32(l.ln1_b, D));
    W2VS_TRY(copy_act(l.w1, (int64_t)F * D));
    W2VS_TRY(copy_f32(l.b1, F));
    if (l.w2s != kNone) W2VS_TRY(launch_pack_slabs(reinterpret_cast<const float*>(d_ref[t]), at<void>(d_packed, l.w2s), D, F, D, st));
    W2VS_TRY(copy_act(l.w2, (int64_t)D * F));
    W2VS_TRY(copy_f32(l.b2, D));
    W2VS_TRY(copy_f32(l.ln2_w, D));
    W2VS_TRY(copy_f32(l.ln2_b, D));
    if (l.wc != kNone) {   // per-CTA operand pieces + LayerNorm fold vectors of the cluster incremental step, from the tensors just packed
      ClusterPackArgs cp{};
      cp.wqkv = at<void>(d_packed, l.wqkv); cp.wo = at<void>(d_packed, l.wo); cp.w1 = at<void>(d_packed, l.w1); cp.w2 = at<void>(d_packed, l.w2);
      cp.ln1_w = at<float>(d_packed, l.ln1_w); cp.ln1_b = at<float>(d_packed, l.ln1_b); cp.bqkv = at<float>(d_packed, l.bqkv);
      cp.ln2_w = at<float>(d_packed, l.ln2_w); cp.ln2_b = at<float>(d_packed, l.ln2_b); cp.b1 = at<float>(d_packed, l.b1);
      cp.dst = at<void>(d_packed, l.wc);
      W2VS_TRY(launch_pack_cluster(cfg, cp, st));
    }
  }
  W2VS_TRY(copy_f32(wl.enc_ln_w, D));
  W2VS_TRY(copy_f32(wl.enc_ln_b, D));
  if (t != n_tensors) { set_error("internal: consumed %d of %d tensors", t, n_tensors); return W2VS_INVALID_VALUE; }
  return W2VS_OK;
}

w2vs_status_t w2vs_geometry_of(const w2vs_config* cfg, int32_t L, int32_t main_ctx, int32_t right_ctx,
                               w2vs_geometry* out) {
  W2VS_TRY(validate_config(cfg));
  W2VS_REQUIRE(out != nullptr, "out is NULL");
  Geometry g;
  W2VS_TRY(make_geometry(cfg, L, main_ctx, right_ctx, &g));
  memset(out, 0, sizeof(*out));
  out->frames = g.T;
  out->frames_pad = g.T2;
  out->n_blocks = g.nb;
  out->tokens = g.M;
  for (int i = 0; i < cfg->n_conv; ++i) { out->conv_len[i] = g.conv_len[i]; out->conv_rows[i] = g.conv_rows[i]; }
  return W2VS_OK;
}

w2vs_status_t w2vs_get_workspace_size(const w2vs_config* cfg, int32_t B, int32_t L, int32_t main_ctx,
                                      int32_t right_ctx, size_t* bytes) {
  W2VS_TRY(validate_config(cfg));
  W2VS_REQUIRE(bytes != nullptr, "bytes is NULL");
  W2VS_REQUIRE(B >= 1, "B");
  Geometry g;
  W2VS_TRY(make_geometry(cfg, L, main_ctx, right_ctx, &g));
  Workspace ws;
  make_workspace(cfg, g, B, &ws);
  *bytes = ws.total;
  return W2VS_OK;
}

w2vs_status_t w2vs_encode(const w2vs_config* cfg, const void* d_packed, const w2vs_encode_args* a,
                          void* d_ws, size_t ws_bytes, void* stream) {
  W2VS_TRY(validate_config(cfg));
  W2VS_REQUIRE(a != nullptr && d_packed != nullptr && d_ws != nullptr, "NULL pointer");
  W2VS_REQUIRE(a->d_wav != nullptr && a->d_out != nullptr, "NULL wav / out");
  W2VS_REQUIRE(a->B >= 1 && a->L >= 1, "B / L");
  W2VS_REQUIRE(a->wav_dtype == W2VS_F32 || a->wav_dtype == W2VS_BF16 || a->wav_dtype == W2VS_I16, "wav_dtype");
  W2VS_REQUIRE(!(a->d_lengths && a->d_sample_pad_mask), "give d_lengths or d_sample_pad_mask, not both");
  W2VS_REQUIRE(!(a->wav_normalize && a->d_sample_pad_mask), "wav_normalize needs d_lengths (or no padding), not a sample mask");
  W2VS_REQUIRE(a->out_layout == W2VS_LAYOUT_BTD || a->out_layout == W2VS_LAYOUT_TBD, "out_layout");
  Geometry g;
  W2VS_TRY(make_geometry(cfg, a->L, a->main_ctx, a->right_ctx, &g));
  const bool has_mask = a->d_lengths || a->d_sample_pad_mask;
  if (has_mask) W2VS_REQUIRE(a->mask_len >= g.T, "mask_len must be >= number of frames");
  W2VS_REQUIRE(a->drop_tail_frames >= 0 && a->drop_tail_frames <= g.T, "drop_tail_frames");
  if (cfg->pos_type == W2VS_POS_SIN)
    W2VS_REQUIRE(cfg->sin_rows >= g.T + 2, "sinusoidal table too short for this utterance (sin_rows < T+2)");
  Workspace ws;
  make_workspace(cfg, g, a->B, &ws);
  if (ws_bytes < ws.total) { set_error("workspace too small: %zu < %zu", ws_bytes, ws.total); return W2VS_WORKSPACE_TOO_SMALL; }
  WeightLayout wl;
  make_weight_layout(cfg, &wl);

  cudaStream_t st = (cudaStream_t)stream;
  const int adt = cfg->dtype;
  const size_t as = act_size(cfg);
  const int B = a->B, n = cfg->n_conv, D = cfg->embed_dim, F = cfg->ffn_dim;
  const int gemm_impl = W2VS_GEMM_AUTO;
  const void* W = d_packed;

  // ---- feature extractor --------------------------------------------------------------------
  void* bufs[2] = {at<void>(d_ws, ws.conv_a), at<void>(d_ws, ws.conv_b)};
  {
    Conv0Args c{};
    c.wav = a->d_wav; c.wav_dtype = a->wav_dtype; c.wav_ld = a->L;
    c.w = at<float>(W, wl.conv[0].w); c.bias = at<float>(W, wl.conv[0].bias);
    c.gamma = at<float>(W, wl.conv[0].norm_w); c.beta = at<float>(W, wl.conv[0].norm_b);
    c.out = bufs[0]; c.out_dtype = adt;
    c.B = B; c.T0 = g.conv_len[0]; c.rows_per_utt = g.conv_rows[0];
    c.C = cfg->conv_dim[0]; c.k = cfg->conv_kernel[0]; c.stride = cfg->conv_stride[0];
    c.norm = !conv_has_norm(cfg, 0) ? CONV0_NORM_NONE
             : (cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM ? CONV0_NORM_LAYER : CONV0_NORM_GROUP);
    c.gn_stats = at<float>(d_ws, ws.gn_stats);
    if (a->wav_normalize) {
      float* wst = at<float>(d_ws, ws.wav_stats);
      W2VS_TRY(launch_wav_stats(a->d_wav, a->wav_dtype, a->L, a->d_lengths, a->L, B, wst, st));
      c.wav_stats = wst; c.wav_lengths = a->d_lengths;
    }
    W2VS_TRY(launch_conv0(c, st));
  }
  for (int i = 1; i < n; ++i) {
    const int cin = cfg->conv_dim[i - 1], cout = cfg->conv_dim[i];
    const int k = cfg->conv_kernel[i], s = cfg->conv_stride[i];
    const bool ln = conv_has_norm(cfg, i);
    void* src = bufs[(i - 1) & 1];
    void* dst = bufs[i & 1];
    GemmArgs ga{};
    ga.A = src; ga.lda = (int64_t)s * cin; ga.a_rows = (int64_t)B * g.conv_rows[i] + 1;
    ga.W = at<void>(W, wl.conv[i].w); ga.bias = at<float>(W, wl.conv[i].bias); ga.residual = nullptr;
    ga.M = B * g.conv_rows[i]; ga.N = cout; ga.K = k * cin; ga.dtype_ab = adt;
    if (ln) {
      // The pre-norm activation is stored in the model dtype, exactly like the reference module chain
      // (Conv1d output dtype -> Fp32LayerNorm upcasts it, layer_norm.py:39-50); LayerNorm + GELU run in place
      // with fp32 statistics.
      ga.C = dst; ga.ldc = cout; ga.dtype_c = adt; ga.flags = 0;
      W2VS_TRY(launch_gemm(gemm_impl, ga, st));
      LayerNormArgs la{};
      la.x = dst; la.in_dtype = adt; la.ldx = cout;
      la.gamma = at<float>(W, wl.conv[i].norm_w); la.beta = at<float>(W, wl.conv[i].norm_b);
      la.out_f32 = nullptr; la.out_act = dst; la.act_dtype = adt; la.ldo = cout;
      la.rows = ga.M; la.N = cout; la.gelu = 1;
      W2VS_TRY(launch_layernorm(la, st));
    } else {
      ga.C = dst; ga.ldc = cout; ga.dtype_c = adt; ga.flags = W2VS_EPI_GELU;
      W2VS_TRY(launch_gemm(gemm_impl, ga, st));
    }
  }
  const int CL = c_last(cfg);
  const int rows_last = g.conv_rows[n - 1];
  void* conv_out = bufs[(n - 1) & 1];
  if (a->d_tap_conv_out) W2VS_TRY(launch_tap_rows(conv_out, adt, rows_last, a->d_tap_conv_out, B, g.T, CL, st));

  // ---- feature LayerNorm + post_extract_proj -----------------------------------------------------
  float* feats = at<float>(d_ws, ws.feats);
  {
    LayerNormArgs la{};
    la.x = conv_out; la.in_dtype = adt; la.ldx = CL;
    la.gamma = at<float>(W, wl.feat_ln_w); la.beta = at<float>(W, wl.feat_ln_b);
    la.rows = B * rows_last; la.N = CL; la.gelu = 0; la.ldo = CL; la.act_dtype = adt;
    if (CL != D) {
      void* normed = bufs[n & 1];
      la.out_f32 = nullptr; la.out_act = normed;
      W2VS_TRY(launch_layernorm(la, st));
      GemmArgs ga{};
      ga.A = normed; ga.lda = CL; ga.a_rows = (int64_t)B * rows_last;
      ga.W = at<void>(W, wl.proj_w); ga.bias = at<float>(W, wl.proj_b); ga.residual = nullptr;
      ga.C = feats; ga.ldc = D; ga.M = B * rows_last; ga.N = D; ga.K = CL;
      ga.dtype_ab = adt; ga.dtype_c = W2VS_F32; ga.flags = 0;
      W2VS_TRY(launch_gemm(gemm_impl, ga, st));
    } else {
      la.out_f32 = feats; la.out_act = nullptr;
      W2VS_TRY(launch_layernorm(la, st));
    }
  }
  if (a->d_tap_post_proj) W2VS_TRY(launch_tap_rows(feats, W2VS_F32, rows_last, a->d_tap_post_proj, B, g.T, D, st));

  // ---- masks, positions, token buffer --------------------------------------------------------------
  uint8_t* frame_pad = at<uint8_t>(d_ws, ws.frame_pad);
  int32_t* pos = at<int32_t>(d_ws, ws.pos);
  uint8_t* keypad = at<uint8_t>(d_ws, ws.keypad);
  {
    PrepArgs p{};
    p.lengths = a->d_lengths; p.sample_mask = a->d_sample_pad_mask; p.mask_len = has_mask ? a->mask_len : 0;
    p.frame_pad = frame_pad; p.pos = pos; p.keypad = keypad; p.pad_blk = at<uint8_t>(d_ws, ws.pad_blk);
    p.B = B; p.T = g.T; p.T2 = g.T2; p.M = g.M; p.main_ctx = g.main_ctx; p.rc = g.rc > 0 ? g.rc : 1;
    W2VS_TRY(launch_prep_masks(p, st));
  }
  float* X = at<float>(d_ws, ws.x);
  void* Xa = at<void>(d_ws, ws.xa);
  {
    const float* posconv = nullptr;
    int posconv_rows = 0;
    if (posconv_tc(cfg)) {
      // grouped Conv1d(D, D, k, groups) + SamePad + GELU (wav2vec2.py:791-804) as one implicit GEMM per group, all in
      // one launch of the tcgen05 kernel: bias and GELU in its epilogue, fp32 result rows b*Tp + t read by embed_tokens below
      const int k = cfg->conv_pos, groups = cfg->conv_pos_groups, Dg = D / groups, Dgp = posconv_dgp(cfg);
      const int Tp = g.T + k;
      const int64_t rows_tot = (int64_t)B * Tp + k;
      bf16* xg = at<bf16>(d_ws, ws.posconv_xg);
      float* out = at<float>(d_ws, ws.posconv_tmp);
      W2VS_TRY(launch_posconv_pack_x(feats, rows_last, frame_pad, xg, B, g.T, D, k, groups, Dgp, st));
      {
        // all groups in one launch: product grp reads the rows_tot frames of its group, Dg rows of W, and writes
        // the column block grp * Dg of the result
        GemmArgs ga{};
        ga.A = xg; ga.lda = Dgp; ga.a_rows = (int64_t)groups * rows_tot;
        ga.W = at<bf16>(W, wl.posconv_wg); ga.bias = at<float>(W, wl.posconv_b); ga.residual = nullptr;
        ga.C = out; ga.ldc = D; ga.M = (B - 1) * Tp + g.T; ga.N = Dg; ga.K = k * Dgp;
        ga.dtype_ab = W2VS_BF16; ga.dtype_c = W2VS_F32; ga.flags = W2VS_EPI_GELU;
        ga.batch = groups; ga.a_batch_rows = rows_tot; ga.w_batch_rows = Dg; ga.c_batch_stride = Dg;
        W2VS_TRY(launch_gemm(W2VS_GEMM_TCGEN05_2CTA, ga, st));
      }
      posconv = out; posconv_rows = Tp;
    } else if (cfg->pos_type == W2VS_POS_CONV) {
      PosConvArgs pc{};
      pc.feats = feats; pc.feat_rows = rows_last; pc.frame_pad = frame_pad;
      pc.w = at<float>(W, wl.posconv_w); pc.bias = at<float>(W, wl.posconv_b);
      pc.out = at<float>(d_ws, ws.posconv_tmp);
      pc.B = B; pc.T = g.T; pc.D = D; pc.k = cfg->conv_pos; pc.groups = cfg->conv_pos_groups;
      W2VS_TRY(launch_posconv(pc, st));
      posconv = pc.out;
    }
    EmbedArgs e{};
    e.feats = feats; e.feat_rows = rows_last; e.frame_pad = frame_pad; e.pos = pos; e.pos_offset = 0;
    e.sin_table = at<float>(W, wl.sin_table); e.posconv = posconv; e.posconv_rows = posconv_rows;
    e.gamma = cfg->layer_norm_first ? nullptr : at<float>(W, wl.enc_ln_w);
    e.beta = cfg->layer_norm_first ? nullptr : at<float>(W, wl.enc_ln_b);
    e.X = X; e.Xa = Xa; e.act_dtype = adt;
    e.B = B; e.T = g.T; e.T2 = g.T2; e.M = g.M; e.main_ctx = g.main_ctx; e.rc = g.rc > 0 ? g.rc : 1; e.D = D;
    W2VS_TRY(launch_embed(e, st));
  }
  const int tokens = B * g.M;
  if (a->d_tap_enc_in) W2VS_TRY(launch_tap_rows(X, W2VS_F32, g.M, a->d_tap_enc_in, B, g.M, D, st));

  // ---- transformer layers ---------------------------------------------------------------------------
  void* qkv = at<void>(d_ws, ws.qkv);
  void* ctx = at<void>(d_ws, ws.ctx);
  void* h = at<void>(d_ws, ws.h);
  auto layer_norm = [&](size_t gw, size_t gb, bool write_f32) {
    LayerNormArgs la{};
    la.x = X; la.in_dtype = W2VS_F32; la.ldx = D;
    la.gamma = at<float>(W, gw); la.beta = at<float>(W, gb);
    la.out_f32 = write_f32 ? X : nullptr; la.out_act = Xa; la.act_dtype = adt; la.ldo = D;
    la.rows = tokens; la.N = D; la.gelu = 0;
    return launch_layernorm(la, st);
  };
  auto gemm = [&](const void* A, int K, size_t w, size_t b, const float* res, void* C, int N, int cdt, int flags) {
    GemmArgs ga{};
    ga.A = A; ga.lda = K; ga.a_rows = tokens; ga.W = at<void>(W, w); ga.bias = at<float>(W, b);
    ga.residual = res; ga.C = C; ga.ldc = N; ga.M = tokens; ga.N = N; ga.K = K;
    ga.dtype_ab = adt; ga.dtype_c = cdt; ga.flags = flags;
    return launch_gemm(gemm_impl, ga, st);
  };
  const bool pre_ln = cfg->layer_norm_first != 0;
  for (int l = 0; l < cfg->layers; ++l) {
    const LayerW lw = layer_at(wl, l);
    if (pre_ln) W2VS_TRY(layer_norm(lw.ln1_w, lw.ln1_b, false));
    W2VS_TRY(gemm(Xa, D, lw.wqkv, lw.bqkv, nullptr, qkv, 3 * D, adt, 0));
    {
      AttnArgs aa{};
      aa.qkv = qkv; aa.keypad = keypad; aa.ctx = ctx; aa.dtype = adt; aa.pad_blk = at<uint8_t>(d_ws, ws.pad_blk);
      aa.B = B; aa.T2 = g.T2; aa.main_ctx = g.main_ctx; aa.rc = g.rc; aa.heads = cfg->heads; aa.D = D;
      W2VS_TRY(launch_attention(0, aa, st));
    }
    W2VS_TRY(gemm(ctx, D, lw.wo, lw.bo, X, X, D, W2VS_F32, 0));
    if (pre_ln) W2VS_TRY(layer_norm(lw.ln2_w, lw.ln2_b, false));
    else W2VS_TRY(layer_norm(lw.ln1_w, lw.ln1_b, true));
    W2VS_TRY(gemm(Xa, D, lw.w1, lw.b1, nullptr, h, F, adt, W2VS_EPI_GELU));
    W2VS_TRY(gemm(h, F, lw.w2, lw.b2, X, X, D, W2VS_F32, 0));
    if (!pre_ln) W2VS_TRY(layer_norm(lw.ln2_w, lw.ln2_b, true));
    if (a->d_tap_layers)
      W2VS_TRY(launch_tap_rows(X, W2VS_F32, g.M, a->d_tap_layers + (size_t)l * tokens * D, B, g.M, D, st));
  }

  // ---- output -----------------------------------------------------------------------------------------
  const int T_out = g.T - a->drop_tail_frames;
  {
    FinalizeArgs f{};
    f.X = X; f.gamma = pre_ln ? at<float>(W, wl.enc_ln_w) : nullptr; f.beta = pre_ln ? at<float>(W, wl.enc_ln_b) : nullptr;
    f.out = a->d_out; f.out_dtype = cfg->io_dtype == W2VS_F16 ? W2VS_F16 : adt; f.B = B; f.T_out = T_out; f.in_rows_per_utt = g.M; f.D = D;
    f.tbd = a->out_layout == W2VS_LAYOUT_TBD;
    W2VS_TRY(launch_finalize(f, st));
  }
  if (a->d_out_pad_mask) W2VS_TRY(launch_copy_mask(frame_pad, g.T, a->d_out_pad_mask, T_out, B, T_out, st));
  (void)as;
  return W2VS_OK;
}

// ---- single operators -----------------------------------------------------------------------------------
w2vs_status_t w2vs_op_gemm(int32_t impl, int32_t dtype_ab, int32_t dtype_c, const void* d_A, int64_t lda,
                           const void* d_W, const float* d_bias, const float* d_residual, void* d_C,
                           int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t flags, void* stream) {
  W2VS_REQUIRE(d_A && d_W && d_C, "NULL pointer");
  W2VS_REQUIRE(M >= 0 && N >= 1 && K >= 1, "M/N/K");
  W2VS_REQUIRE(lda >= 1 && ldc >= N, "lda/ldc");
  GemmArgs g{};
  g.A = d_A; g.lda = lda; g.W = d_W; g.bias = d_bias; g.residual = d_residual; g.C = d_C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.dtype_ab = dtype_ab; g.dtype_c = dtype_c; g.flags = flags;
  // rows of length lda addressable behind A: the last output row reads K elements starting at (M-1)*lda
  g.a_rows = lda >= K ? M : M + (K - 1) / lda;
  return launch_gemm(impl, g, (cudaStream_t)stream);
}

w2vs_status_t w2vs_op_layernorm(int32_t dtype_in, const void* d_x, int64_t ldx, const float* d_gamma,
                                const float* d_beta, float* d_out_f32, int32_t dtype_act, void* d_out_act,
                                int64_t ldo, int32_t rows, int32_t N, int32_t gelu, void* stream) {
  W2VS_REQUIRE(d_x && d_gamma && d_beta && (d_out_f32 || d_out_act), "NULL pointer");
  LayerNormArgs a{};
  a.x = d_x; a.in_dtype = dtype_in; a.ldx = ldx; a.gamma = d_gamma; a.beta = d_beta;
  a.out_f32 = d_out_f32; a.out_act = d_out_act; a.act_dtype = dtype_act; a.ldo = ldo;
  a.rows = rows; a.N = N; a.gelu = gelu;
  return launch_layernorm(a, (cudaStream_t)stream);
}

w2vs_status_t w2vs_op_attention(int32_t impl, int32_t dtype, const void* d_qkv, const uint8_t* d_keypad,
                                void* d_ctx, int32_t B, int32_t T_pad, int32_t main_ctx, int32_t right_ctx,
                                int32_t heads, int32_t D, void* stream) {
  W2VS_REQUIRE(d_qkv && d_keypad && d_ctx, "NULL pointer");
  W2VS_REQUIRE(B >= 1 && T_pad >= 1 && main_ctx >= 1 && right_ctx >= 0 && heads >= 1, "shape");
  AttnArgs a{};
  a.qkv = d_qkv; a.keypad = d_keypad; a.ctx = d_ctx; a.dtype = dtype;
  a.B = B; a.T2 = T_pad; a.main_ctx = main_ctx; a.rc = right_ctx; a.heads = heads; a.D = D;
  return launch_attention(impl, a, (cudaStream_t)stream);
}

w2vs_status_t w2vs_debug_fused_trace(uint64_t* out, int32_t n) {
  W2VS_REQUIRE(out != nullptr && n >= 0, "out / n");
  return debug_read_fused_trace(reinterpret_cast<unsigned long long*>(out), n);
}

w2vs_status_t w2vs_debug_cluster_trace(uint64_t* out, int32_t n) {
  if (out == nullptr) { debug_cluster_trace_enable(n != 0); return W2VS_OK; }      // (NULL, on): switch the stamps on / off
  W2VS_REQUIRE(n >= 0, "n");
  return debug_read_cluster_trace(reinterpret_cast<unsigned long long*>(out), n);
}

w2vs_status_t w2vs_debug_fault_flags(int32_t* flags) {
  W2VS_REQUIRE(flags != nullptr, "flags is NULL");
  int a = 0, b = 0, c = 0, d = 0;
  W2VS_TRY(debug_read_tc2_fault(&a));
  W2VS_TRY(debug_read_attn_tc_fault(&b));
  W2VS_TRY(debug_read_fused_fault(&c));
  W2VS_TRY(debug_read_cluster_fault(&d));
  *flags = (a ? 1 : 0) | (b ? 2 : 0) | (c ? 4 : 0) | (d ? 8 : 0);
  return W2VS_OK;
}

}  // extern "C"

namespace w2vs {

w2vs_status_t launch_gemm(int impl, const GemmArgs& g, cudaStream_t st) {
  W2VS_REQUIRE(g.dtype_ab == W2VS_F32 || g.dtype_ab == W2VS_BF16, "GEMM operand dtype");
  W2VS_REQUIRE(g.dtype_c == W2VS_F32 || g.dtype_c == W2VS_BF16, "GEMM output dtype");
  if (impl == W2VS_GEMM_AUTO)
    impl = g.dtype_ab != W2VS_BF16 ? W2VS_GEMM_SIMT : gemm_skinny_applicable(g) ? W2VS_GEMM_SKINNY : W2VS_GEMM_TCGEN05_2CTA;
  if (impl == W2VS_GEMM_SKINNY) return launch_gemm_skinny(g, st);
  if (impl == W2VS_GEMM_TCGEN05_2CTA) return launch_gemm_tc2(g, st);   // residual: in place (fp32, residual == C)
  if (impl == W2VS_GEMM_SIMT) return launch_gemm_simt(g, st);
  set_error("invalid value: gemm impl %d", impl);
  return W2VS_INVALID_VALUE;
}

w2vs_status_t launch_attention(int impl, const AttnArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(a.dtype == W2VS_F32 || a.dtype == W2VS_BF16, "attention dtype");
  if (impl == 0) impl = a.dtype == W2VS_BF16 ? (a.n_step_q > 0 ? 2 : 3) : 1;
  if (impl == 3) return launch_attention_tc(a, st);
  if (impl == 2) {
    W2VS_REQUIRE(a.dtype == W2VS_BF16, "tensor-core attention takes bf16");
    return launch_attention_mma(a, st);
  }
  if (impl == 1) return launch_attention_simt(a, st);
  set_error("invalid value: attention impl %d", impl);
  return W2VS_INVALID_VALUE;
}

}  // namespace w2vs
