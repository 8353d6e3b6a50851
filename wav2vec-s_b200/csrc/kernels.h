// Internal launcher interface between the C ABI (api.cu / stream.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/w2vs.h"

namespace w2vs {

// ---- conv0 ------------------------------------------------------------------------------------
enum { CONV0_NORM_NONE = 0, CONV0_NORM_LAYER = 1, CONV0_NORM_GROUP = 2 };
struct Conv0Args {
  const void* wav; int wav_dtype; int64_t wav_ld;
  const float *w, *bias, *gamma, *beta;
  void* out; int out_dtype;
  int B, T0, rows_per_utt, C, k, stride, norm;
  float* gn_stats;   // [B][C][2] scale/shift + [B][ceil(T0/256)][2][C] partials (GroupNorm mode)
  const float* wav_stats;      // optional [B][2] (mean, rstd): samples are standardised on load (waveform front end)
  const int32_t* wav_lengths;  // optional [B]: samples at or past the length are left as they are
};
w2vs_status_t launch_conv0(const Conv0Args& a, cudaStream_t st);
// per-utterance (mean, 1/sqrt(var + 1e-5)) over the valid samples: F.layer_norm(feats, feats.shape) of the reference
// data pipeline (fairseq/data/audio/raw_audio_dataset.py:60-72)
w2vs_status_t launch_wav_stats(const void* wav, int wav_dtype, int64_t wav_ld, const int32_t* lengths, int L, int B,
                               float* stats, cudaStream_t st);

// ---- row kernels --------------------------------------------------------------------------------
struct LayerNormArgs {
  const void* x; int in_dtype; int64_t ldx;
  const float *gamma, *beta;
  float* out_f32; void* out_act; int act_dtype; int64_t ldo;
  int rows, N, gelu;
};
w2vs_status_t launch_layernorm(const LayerNormArgs& a, cudaStream_t st);

struct PrepArgs {
  const int32_t* lengths; const uint8_t* sample_mask; int mask_len;
  uint8_t* frame_pad; int32_t* pos; uint8_t* keypad;
  uint8_t* pad_blk;                 // [B][ceil(M/128)] any-padding flags per 128 tokens (may be NULL)
  int B, T, T2, M, main_ctx, rc;
};
w2vs_status_t launch_prep_masks(const PrepArgs& a, cudaStream_t st);

struct EmbedArgs {
  const float* feats; int feat_rows;          // [B*feat_rows, D]
  const uint8_t* frame_pad; const int32_t* pos; int pos_offset;
  const float* sin_table; const float* posconv;
  int posconv_rows;                           // rows per utterance of `posconv` (0: T)
  const float *gamma, *beta;                  // encoder.layer_norm when post-LN, else NULL
  float* X; void* Xa; int act_dtype;
  int B, T, T2, M, main_ctx, rc, D;
};
w2vs_status_t launch_embed(const EmbedArgs& a, cudaStream_t st);

struct FinalizeArgs {
  const float* X; const float *gamma, *beta;
  void* out; int out_dtype;
  int B, T_out; int64_t in_rows_per_utt; int D, tbd;
};
w2vs_status_t launch_finalize(const FinalizeArgs& a, cudaStream_t st);

w2vs_status_t launch_tap_rows(const void* src, int src_dtype, int64_t src_rows_per_utt, float* dst, int B,
                              int T, int C, cudaStream_t st);
w2vs_status_t launch_copy_mask(const uint8_t* src, int src_ld, uint8_t* dst, int dst_ld, int B, int n,
                               cudaStream_t st);

// ---- GEMM ---------------------------------------------------------------------------------------
struct GemmArgs {
  const void* A; int64_t lda; int64_t a_rows;   // a_rows: rows of length lda addressable behind A (TMA bound)
  const void* W;                                // [N, K] row-major
  const float* bias; const float* residual;     // residual fp32 [M, ldc] (may alias C)
  void* C; int64_t ldc;
  int M, N, K; int dtype_ab, dtype_c; int flags;
  // batch > 1 (CTA-pair tcgen05 kernel only): `batch` independent products in one launch -- A of product b starts
  // a_batch_rows rows (of lda) after A of b - 1, W w_batch_rows rows after, C c_batch_stride elements after, bias N
  // floats after; no residual.  The groups of the positional conv.
  int batch; int64_t a_batch_rows, w_batch_rows, c_batch_stride;
  // pdl: the predecessor in the stream is a small row kernel (a LayerNorm of an incremental step): launch the tcgen05
  // kernel as its programmatic dependent, so that barrier set-up, TMEM allocation and the cluster handshake overlap it
  int pdl;
};
w2vs_status_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st);
w2vs_status_t launch_gemm_tc2(const GemmArgs& g, cudaStream_t st);   // CTA-pair tcgen05, TMA-store epilogue
w2vs_status_t launch_gemm_skinny(const GemmArgs& g, cudaStream_t st); // M <= 64 weight-streaming kernel (mma.sync)
bool gemm_skinny_applicable(const GemmArgs& g);
w2vs_status_t launch_gemm(int impl, const GemmArgs& g, cudaStream_t st);  // impl: w2vs_gemm_impl_t
w2vs_status_t debug_read_tc2_fault(int* out);
w2vs_status_t debug_read_attn_tc_fault(int* out);

// ---- attention ------------------------------------------------------------------------------------
// Two modes share the kernels:
//   block mode (n_step_q == 0): queries = keys = the M = T2 + nb*rc tokens of `qkv` [B, M, 3D]; visibility
//       from (T2, main_ctx, rc) + key padding bytes.
//   step mode  (n_step_q  > 0): incremental inference -- n_step_q query tokens per stream (rows of `qkv`,
//       [B, n_step_q, 3D]) attend to the first n_step_keys rows of a K/V cache [B, kv_rows, 2D] (K | V per
//       row); everything is visible, no padding.
struct AttnArgs {
  const void* qkv; const uint8_t* keypad; void* ctx; int dtype;
  int B, T2, main_ctx, rc, heads, D;
  int n_step_q, n_step_keys; const void* kv_cache; int64_t kv_rows;
  const uint8_t* pad_blk;           // optional [B][ceil(M/128)] flags from prep_masks (block mode, tcgen05 kernel)
  // step mode, optional: split the cached keys of a (stream, head) over several CTAs (partial softmax state per
  // split, combined by the CTA that finishes last).  step_partials: fp32 [B][heads][q tiles][8][64][66];
  // step_counters: zeroed unsigned [B][heads][q tiles] (the kernel resets them).
  float* step_partials; unsigned* step_counters;
};
constexpr int kAttnStepMaxSplits = 8;
w2vs_status_t launch_attention_simt(const AttnArgs& a, cudaStream_t st);
w2vs_status_t launch_attention_mma(const AttnArgs& a, cudaStream_t st);   // mma.sync flash kernel (also step mode)
w2vs_status_t launch_attention_tc(const AttnArgs& a, cudaStream_t st);    // tcgen05 / TMEM kernel (block mode)
w2vs_status_t launch_attention(int impl, const AttnArgs& a, cudaStream_t st);

// ---- incremental mode data movement (k_stream.cu) ------------------------------------------------------
w2vs_status_t launch_concat_rows(void* dst, int64_t dst_bs_bytes, const void* srcA, int64_t a_bs_bytes, int offA,
                                 int cA, const void* srcB, int64_t b_bs_bytes, int nB, int row_bytes, int B,
                                 cudaStream_t st);
w2vs_status_t launch_concat_wav(float* dst, int64_t dst_bs, const float* srcA, int64_t a_bs, int offA, int cA,
                                const void* src_new, int new_dtype, int64_t new_bs, int n, int B, cudaStream_t st);
w2vs_status_t launch_kv_append(const void* qkv, void* cache, int64_t cache_rows, int row0, int n_tok, int D,
                               int elem_bytes, int B, cudaStream_t st);

// ---- one conv block of the feature extractor on the new rows of a decision step, one launch (k_conv_step.cu):
// [carry rows | new rows] -> Conv1d(C, C, k, s) + bias -> LayerNorm -> GELU -> out; the rows the next step needs again
// are copied to carry_out.  One stream, bf16, C in {512, 64}, k in {2, 3}, s = 2.
struct ConvStepArgs {
  const void* carry; int n_carry; const void* fresh; int n_fresh;
  const void* W; const float *bias, *gamma, *beta;
  void* out; int n_out; void* carry_out;
  int k, s, C;
};
bool conv_step_applicable(const ConvStepArgs& a);
w2vs_status_t launch_conv_step(const ConvStepArgs& a, cudaStream_t st);

// feature LayerNorm + post_extract_proj + append of a step's new frames, one launch (k_conv_step.cu): x [rows][K] bf16
// -> LN(gamma, beta) -> . W[N][K]^T + bias -> out [rows][N] fp32 (rows <= 64, K in {256, 512})
struct FeatProjArgs {
  const void* x; int rows, K; const float *gamma, *beta; const void* W; const float* bias; float* out; int N;
};
bool feat_proj_applicable(const FeatProjArgs& a);
w2vs_status_t launch_feat_proj(const FeatProjArgs& a, cudaStream_t st);

// ---- fused incremental step (k_stream_fused.cu): embed -> all layers -> final LayerNorm in one cooperative kernel ----
struct WeightLayout;
struct StreamFusedArgs {
  const w2vs_config* cfg; const WeightLayout* wl; const void* W;
  int B, ntok, n_main, f0;
  const float* feats; int64_t feat_rows;     // projected frames [B][feat_rows][D]
  float* R; void* q; void* ctx; void* h;     // workspace: residual stream fp32, q / context / FFN hidden (bf16)
  void* kv; int64_t kv_layer_elems, kv_rows; // K/V cache [layers][B][kv_rows][2D]
  float* partials; int max_splits;           // attention split states, room for max_splits x 32 rows x 66 per (stream, head)
  unsigned* counters;                        // zeroed [B * heads]
  void* out;                                 // [n_main][B][D] bf16
  unsigned long long* bar;                   // two zeroed 64-bit words (grid barrier arrivals / departures)
};
bool stream_fused_applicable(const w2vs_config* cfg, int B, int ntok);
w2vs_status_t launch_stream_fused(const StreamFusedArgs& a, cudaStream_t st);
w2vs_status_t debug_read_fused_fault(int* out);
w2vs_status_t debug_read_fused_trace(unsigned long long* out, int n);

// ---- cluster incremental step (k_stream_cluster.cu): the same step as ONE kernel of H clusters x 8 CTAs, two grid
// barriers per layer (one stream, pre-LN bf16 models of the instantiated shapes); takes the arguments above (q / ctx /
// h / partials / counters unused, R = three residual buffers of 32 rows) and the per-CTA weight copy LayerW::wc ----
constexpr int kStreamClusterSize = 4;
size_t stream_cluster_layer_bytes(const w2vs_config* cfg);      // bytes of LayerW::wc per layer, 0 = kernel not available
bool stream_cluster_applicable(const w2vs_config* cfg, int B, int ntok);
w2vs_status_t launch_stream_cluster(const StreamFusedArgs& a, cudaStream_t st);
struct ClusterPackArgs {      // packed tensors of one layer (device pointers into the blob) -> its LayerW::wc region
  const void *wqkv, *wo, *w1, *w2;
  const float *ln1_w, *ln1_b, *bqkv, *ln2_w, *ln2_b, *b1;
  void* dst;
};
w2vs_status_t launch_pack_cluster(const w2vs_config* cfg, const ClusterPackArgs& p, cudaStream_t st);
w2vs_status_t debug_read_cluster_fault(int* out);
w2vs_status_t debug_read_cluster_trace(unsigned long long* out, int n);
void debug_cluster_trace_enable(int on);

// ---- positional conv + weight packing -----------------------------------------------------------------
struct PosConvArgs {
  const float* feats; int feat_rows; const uint8_t* frame_pad;
  const float* w; const float* bias;            // folded weights [groups][k][Dg_in][Dg_out]
  float* out;                                   // [B, T, D]
  int B, T, D, k, groups;
};
w2vs_status_t launch_posconv(const PosConvArgs& a, cudaStream_t st);
// tensor-core path: group-major bf16 copy of the (padding-zeroed) frames, one [rows_tot][Dgp] matrix per group with
// k/2 zero rows in front of every utterance (Tp = T + k rows per utterance), and the matching weight operand
w2vs_status_t launch_posconv_pack_x(const float* feats, int feat_rows, const uint8_t* frame_pad, void* xg, int B,
                                    int T, int D, int k, int groups, int Dgp, cudaStream_t st);
w2vs_status_t launch_pack_posconv_tc(const float* w_folded, void* dst, int D, int groups, int k, int Dgp,
                                     cudaStream_t st);

w2vs_status_t launch_pack_copy(const float* src, void* dst, int dst_dtype, int64_t n, cudaStream_t st);
w2vs_status_t launch_pack_slabs(const float* src, void* dst, int N, int K, int KC, cudaStream_t st);
w2vs_status_t launch_pack_conv(const float* src, void* dst, int dst_dtype, int C_out, int C_in, int k,
                               cudaStream_t st);
w2vs_status_t launch_pack_posconv(const float* g, const float* v, float* dst, int D, int groups, int k,
                                  cudaStream_t st);

}  // namespace w2vs
