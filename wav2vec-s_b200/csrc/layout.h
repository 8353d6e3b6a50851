// Host-side geometry, packed-weight layout and workspace layout.  Pure integer arithmetic; every
// number here must agree bit-for-bit with the reference's shape arithmetic.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../../include/w2vs.h"

namespace w2vs {

struct Geometry {
  int n_conv;
  int conv_len[W2VS_MAX_CONV];    // valid output frames per layer (wav2vec2.py:725: (t-k)/s+1)
  int conv_rows[W2VS_MAX_CONV];   // allocated rows / utterance: rows[i-1] == stride_i * rows[i]
  int T;                          // frames = conv_len[n_conv-1]
  int T2;                         // frames padded to seq_multiple (pad_to_multiple, wav2vec_S.py:375)
  int nb;                         // T2 / main (floor)   (gen_block_attn_mask, wav2vec_S.py:459)
  int R;                          // nb * rc look-ahead copies
  int M;                          // tokens per utterance = T2 + R
  int main_ctx, rc;
};

w2vs_status_t validate_config(const w2vs_config* cfg);
w2vs_status_t make_geometry(const w2vs_config* cfg, int L, int main_ctx, int rc, Geometry* g);

// ---- packed weights ----------------------------------------------------------------------------
// One blob; every offset is 256-byte aligned.  "act" tensors are stored in cfg.dtype (bf16 or fp32),
// everything else fp32.
struct ConvW {
  size_t w;       // layer 0: fp32 [C0][k0];  layers >= 1: act [C_out][k*C_in] (K-major, tap-major inside K)
  size_t bias;    // fp32 [C_out] or SIZE_MAX
  size_t norm_w, norm_b;  // fp32 [C_out] or SIZE_MAX
};
struct LayerW {
  size_t wqkv, bqkv;      // act [3D][D], fp32 [3D]
  size_t wo, bo;          // act [D][D], fp32 [D]
  size_t ln1_w, ln1_b;    // self_attn_layer_norm
  size_t w1, b1;          // act [F][D], fp32 [F]
  size_t w2, b2;          // act [D][F], fp32 [D]
  size_t w2s;             // bf16 [D/8][F/D][8][D]: fc2 weights in the slab order of the fused incremental step (each
                          // 8-row x D-wide slab contiguous), or SIZE_MAX when that kernel does not apply
  size_t ln2_w, ln2_b;    // final_layer_norm
  size_t wc;              // bf16, per (cluster, CTA): the four operand pieces of the cluster incremental step exactly as
                          // they lie in shared memory (k_stream_cluster.cu), or SIZE_MAX when that kernel does not apply
};
struct WeightLayout {
  ConvW conv[W2VS_MAX_CONV];
  size_t feat_ln_w, feat_ln_b;     // layer_norm [C_last]
  size_t proj_w, proj_b;           // act [D][C_last], fp32 [D]; SIZE_MAX if C_last == D
  size_t sin_table;                // fp32 [sin_rows][D]
  size_t posconv_w, posconv_b;     // fp32 [groups][k][Dg_in][Dg_out] folded weight-norm, fp32 [D]
  size_t posconv_wg;               // bf16 [groups][Dg_out][k * Dgp] tensor-core operand (posconv_tc() only), else SIZE_MAX
  size_t enc_ln_w, enc_ln_b;       // encoder.layer_norm
  size_t layers_begin;             // LayerW table is computed by layer_at()
  size_t layer_stride;
  LayerW layer0;                   // offsets of layer 0; layer n = layer0 + n*layer_stride
  size_t total;
};
static const size_t kNone = (size_t)-1;
void make_weight_layout(const w2vs_config* cfg, WeightLayout* wl);
inline LayerW layer_at(const WeightLayout& wl, int n) {
  LayerW l = wl.layer0;
  size_t d = (size_t)n * wl.layer_stride;
  l.wqkv += d; l.bqkv += d; l.wo += d; l.bo += d; l.ln1_w += d; l.ln1_b += d;
  l.w1 += d; l.b1 += d; l.w2 += d; l.b2 += d; l.ln2_w += d; l.ln2_b += d;
  if (l.w2s != kNone) l.w2s += d;
  if (l.wc != kNone) l.wc += d;
  return l;
}

// ---- workspace (full-utterance forward) --------------------------------------------------------
struct Workspace {
  size_t conv_a, conv_b;       // act ping/pong [B*rows_i + slack][C_i]
  size_t gn_stats;             // fp32 [B][C0][2] + per-CTA partials (GroupNorm mode only)
  size_t wav_stats;            // fp32 [B][2]  per-utterance (mean, rstd) of the waveform front end
  size_t feats;                // fp32 [B*rows_last][D]  (post_extract_proj output)
  size_t frame_pad;            // u8 [B][T]
  size_t pos;                  // i32 [B][T]
  size_t keypad;               // u8 [B][M]
  size_t pad_blk;              // u8 [B][ceil(M/128)]
  size_t x;                    // fp32 [B*M][D] residual stream
  size_t xa;                   // act [B*M][D]
  size_t qkv;                  // act [B*M][3D]
  size_t ctx;                  // act [B*M][D]
  size_t h;                    // act [B*M][F]
  size_t posconv_tmp;          // fp32 [B][T][D], or [B * (T + k)][D] on the tensor-core path  (pos_type=conv only)
  size_t posconv_xg;           // bf16 [groups][B * (T + k) + k][Dgp] group-major, zero-padded conv input (posconv_tc())
  size_t total;
};
void make_workspace(const w2vs_config* cfg, const Geometry& g, int B, Workspace* ws);

inline size_t act_size(const w2vs_config* cfg) { return cfg->dtype == W2VS_BF16 ? 2 : 4; }
// Models the persistent incremental-step kernel (k_stream_fused.cu) can run: they carry the slab-ordered fc2 copy.
inline bool stream_fused_model(const w2vs_config* cfg) {
  const int D = cfg->embed_dim, F = cfg->ffn_dim;
  return cfg->dtype == W2VS_BF16 && D % 128 == 0 && D <= 1024 && F % D == 0 && cfg->heads * 64 == D &&
         cfg->pos_type == W2VS_POS_SIN && cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM;
}

// Models the cluster incremental-step kernel (k_stream_cluster.cu) may run, shape permitting (stream_cluster_layer_bytes)
inline bool stream_cluster_model(const w2vs_config* cfg) {
  return cfg->dtype == W2VS_BF16 && cfg->layer_norm_first != 0 && cfg->heads * 64 == cfg->embed_dim &&
         cfg->pos_type == W2VS_POS_SIN && cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM;
}
size_t stream_cluster_layer_bytes(const w2vs_config* cfg);   // k_stream_cluster.cu

// Positional conv (pos_type = conv) on the tensor cores: bf16 models whose group width is a multiple of 8.
// Each group is an implicit GEMM over a group-major copy of the frames, channels padded to Dgp (multiple of 64).
inline bool posconv_tc(const w2vs_config* cfg) {
  return cfg->pos_type == W2VS_POS_CONV && cfg->dtype == W2VS_BF16 && (cfg->embed_dim / cfg->conv_pos_groups) % 8 == 0;
}
inline int posconv_dgp(const w2vs_config* cfg) { return (cfg->embed_dim / cfg->conv_pos_groups + 63) / 64 * 64; }

}  // namespace w2vs
