/*
 * w2vs.h -- C ABI of the B200-native wav2vec-S streaming-encoder forward path.
 *
 * This is the drop-in boundary for ONE hot path of biaofuxmu/wav2vec-S: the encoder forward
 *   waveform -> 7x strided Conv1d (+GroupNorm/LayerNorm, GELU) -> LayerNorm -> Linear ->
 *   positional embedding -> block/chunk-masked Transformer encoder
 * in full-utterance mode (w2vs_encode) and chunk-by-chunk incremental mode (w2vs_stream_*).
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   fairseq/fairseq/models/wav2vec/wav2vec2.py:544-603,667-669   Wav2Vec2Model.forward / extract_features
 *   fairseq/fairseq/models/wav2vec/wav2vec2.py:773-781           ConvFeatureExtractionModel.forward
 *   fairseq/fairseq/models/wav2vec/wav2vec_S.py:355-440,444-489  BlockwiseTransformerEncoder / gen_block_attn_mask
 *   rain/layers/unidirect_w2v2_encoder.py:485-531,254-330        BlockWiseWav2Vec2Model.forward
 *   rain/simul/transducer_agent.py:138-167                       OnlineModels.fwd_encoder (streaming driver)
 *
 * Conventions follow the reference's own native precedent (warp_transducer/include/rnnt.h):
 * extern "C", an int status enum, the CUDA stream passed by the caller, caller-owned workspace
 * sized by *_size() queries, and no device allocation and no host synchronisation inside the
 * library.  All pointers named `d_*` are device pointers; everything else is host memory.  The
 * library is re-entrant per (stream, workspace, state) and works on whichever CUDA device is
 * current on the calling thread: the only process-wide state are idempotent per-device caches
 * (SM count, "dynamic shared memory opted in" flags) and thread-local diagnostics (last error,
 * launch counter).
 */
#ifndef W2VS_H_
#define W2VS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2VS_MAX_CONV 8
#define W2VS_ABI_VERSION 2

typedef enum {
  W2VS_OK = 0,
  W2VS_INVALID_VALUE = 1,      /* bad shape / config / NULL pointer; nothing was launched */
  W2VS_UNSUPPORTED = 2,        /* valid reference config this build does not implement */
  W2VS_WORKSPACE_TOO_SMALL = 3,
  W2VS_CUDA_ERROR = 4          /* a launch or driver call failed */
} w2vs_status_t;

typedef enum { W2VS_F32 = 0, W2VS_BF16 = 1,
               W2VS_I16 = 2,  /* waveform samples only: 16-bit PCM, read as x / 32768 (what the SimulEval agent does on
                                 the host, rain/simul/transducer_searcher.py:74-80) */
               W2VS_F16 = 3   /* input / output element type only (w2vs_config.io_dtype): IEEE half */
} w2vs_dtype_t;
typedef enum { W2VS_EXTRACTOR_DEFAULT = 0, W2VS_EXTRACTOR_LAYER_NORM = 1 } w2vs_extractor_mode_t;
typedef enum { W2VS_POS_SIN = 0, W2VS_POS_CONV = 1 } w2vs_pos_type_t;
typedef enum { W2VS_LAYOUT_BTD = 0, W2VS_LAYOUT_TBD = 1 } w2vs_layout_t;

/* Model hyper-parameters; field meaning = the reference config fields of the same name
 * (Wav2VecSConfig, wav2vec_S.py:43-311). */
typedef struct {
  int32_t abi_version;                 /* W2VS_ABI_VERSION */
  int32_t dtype;                       /* w2vs_dtype_t: arithmetic/activation type of the model */
  int32_t n_conv;                      /* conv_feature_layers */
  int32_t conv_dim[W2VS_MAX_CONV];
  int32_t conv_kernel[W2VS_MAX_CONV];
  int32_t conv_stride[W2VS_MAX_CONV];
  int32_t conv_bias;                   /* conv_bias */
  int32_t extractor_mode;              /* w2vs_extractor_mode_t */
  int32_t layer_norm_num;              /* wav2vec2.py:317: 1 if encoder_layers==12 else 7 */
  int32_t embed_dim;                   /* encoder_embed_dim */
  int32_t ffn_dim;                     /* encoder_ffn_embed_dim */
  int32_t heads;                       /* encoder_attention_heads */
  int32_t layers;                      /* encoder_layers */
  int32_t layer_norm_first;            /* 1 = pre-LN (large), 0 = post-LN (base) */
  int32_t pos_type;                    /* w2vs_pos_type_t */
  int32_t conv_pos;                    /* kernel of the positional conv (128) */
  int32_t conv_pos_groups;             /* 16 */
  int32_t seq_multiple;                /* required_seq_len_multiple (2) */
  int32_t sin_rows;                    /* rows of the sinusoidal table handed to pack (>= T+2) */
  int32_t stream_step_impl;            /* incremental mode: 0 = automatic: one kernel of thread-block clusters per decision
                                          step (k_stream_cluster.cu) for one stream of a pre-LN bf16 model of an
                                          instantiated shape with <= 32 tokens per step, else the chain; 1 = the
                                          kernel-per-operator chain with programmatic dependent launches; 2 = the first
                                          persistent cooperative kernel (bf16, <= 32 tokens per step; slower than the
                                          chain, kept for comparison); 3 = same choice as 0 */
  int32_t io_dtype;                    /* 0 = outputs in `dtype`; W2VS_F16 = the model was given in fp16 (`.half()`, as
                                          the reference trainer does under --fp16, fairseq/fairseq/trainer.py:86-90):
                                          encoder outputs are written as fp16.  Arithmetic is that of dtype = BF16:
                                          bf16 tensor-core operands (the fp16 weights are rounded to bf16 when packed),
                                          fp32 accumulation, residual stream and statistics -- the range-safe choice
                                          for a 24-layer pre-LN residual stream */
  int32_t reserved[5];
} w2vs_config;

/* ---- reference tensors handed to w2vs_weights_pack -------------------------------------
 * fp32 device tensors in the reference's own state_dict layout, in this order
 * (key names: SURVEY.md section 8(a)#1; absent optional tensors are skipped, not NULL-padded):
 *   for i in 0..n_conv-1:  feature_extractor.conv_layers.{i}.0.weight [C_out,C_in,k]
 *                          [.0.bias [C_out]]                       if conv_bias
 *                          [norm .weight,.bias [C_out]]            LayerNorm (.2.1) if layer_norm mode and
 *                                                                  i < layer_norm_num; GroupNorm (.2) if default mode and i==0
 *   layer_norm.weight, layer_norm.bias [C_last]
 *   [post_extract_proj.weight [D,C_last], .bias [D]]              if C_last != D
 *   pos_type==SIN : sinusoidal table [sin_rows, D] (sinusoidal_positional_embedding.py:36-59)
 *   pos_type==CONV: encoder.pos_conv.0.bias [D], .weight_g [1,1,k], .weight_v [D,D/groups,k]
 *   for n in 0..layers-1:  self_attn.{q,k,v,out}_proj.{weight [D,D], bias [D]}  (q,k,v,out order)
 *                          self_attn_layer_norm.{weight,bias}, fc1.{weight [F,D],bias}, fc2.{weight [D,F],bias},
 *                          final_layer_norm.{weight,bias}
 *   encoder.layer_norm.weight, .bias [D]
 */
int32_t w2vs_num_ref_tensors(const w2vs_config* cfg);
w2vs_status_t w2vs_packed_weights_size(const w2vs_config* cfg, size_t* bytes);
/* Repack (transpose conv taps to K-major, concatenate q/k/v, fold weight-norm, cast) on `stream`. */
w2vs_status_t w2vs_weights_pack(const w2vs_config* cfg, const void* const* d_ref_tensors,
                                int32_t n_tensors, void* d_packed, size_t packed_bytes,
                                void* stream);

/* ---- derived integer geometry (bit-exact; host only) -----------------------------------
 * frames       T  = conv stack output length for L samples (wav2vec2.py:725, no padding)
 * frames_pad   T' = T rounded up to seq_multiple            (wav2vec_S.py:375-385)
 * tokens       M  = T' + (T'/main)*rc                        (wav2vec_S.py:444-489)          */
typedef struct {
  int32_t frames, frames_pad, n_blocks, tokens;
  int32_t conv_len[W2VS_MAX_CONV];     /* valid rows per conv layer */
  int32_t conv_rows[W2VS_MAX_CONV];    /* allocated rows per utterance per layer (>= conv_len) */
} w2vs_geometry;
w2vs_status_t w2vs_geometry_of(const w2vs_config* cfg, int32_t L, int32_t main_ctx,
                               int32_t right_ctx, w2vs_geometry* out);

/* ---- full-utterance forward -------------------------------------------------------------- */
typedef struct {
  const void* d_wav;            /* [B, L] samples, row stride L; dtype wav_dtype */
  int32_t wav_dtype;            /* w2vs_dtype_t */
  int32_t B, L;
  /* padding: at most one of d_lengths / d_sample_pad_mask non-NULL; both NULL = no padding mask
   * (padding_mask=None in the reference). mask_len = padding_mask.size(1) (normally L). */
  const int32_t* d_lengths;         /* [B] valid samples: mask[b,i] = i >= len_b (data_utils.py:528-532) */
  const uint8_t* d_sample_pad_mask; /* [B, mask_len] 1 = padding */
  int32_t mask_len;
  int32_t main_ctx, right_ctx;  /* block size and look-ahead, in frames */
  int32_t out_layout;           /* w2vs_layout_t: BTD = fairseq extract_features, TBD = rain forward */
  int32_t drop_tail_frames;     /* rain is_infer && !finished: right_ctx, else 0 (rain :326-328) */
  void* d_out;                  /* [B,T_out,D] or [T_out,B,D], dtype = cfg.dtype; T_out = T - drop_tail_frames */
  uint8_t* d_out_pad_mask;      /* [B, T_out] 1 = padded frame (may be NULL) */
  /* optional per-stage taps for parity tests (fp32, any may be NULL) */
  float* d_tap_conv_out;        /* [B, T, C_last] output of the conv stack, channels-last */
  float* d_tap_post_proj;       /* [B, T, D] after LayerNorm + post_extract_proj */
  float* d_tap_enc_in;          /* [B, M, D] tokens entering layer 0 (incl. rc copies) */
  float* d_tap_layers;          /* [layers, B, M, D] residual stream after each layer */
  /* Waveform front end folded into the first conv layer's load: with
   * wav_normalize != 0 every utterance is standardised over its own valid samples, (x - mean) / sqrt(var + 1e-5)
   * with the biased variance, which is `F.layer_norm(feats, feats.shape)` of the reference data pipeline
   * (fairseq/data/audio/raw_audio_dataset.py:60-72, `normalize: true`); samples past d_lengths[b] are left as
   * they are (the collater pads with zeros after normalising).  Not combinable with d_sample_pad_mask. */
  int32_t wav_normalize;
  int32_t reserved[3];
} w2vs_encode_args;

w2vs_status_t w2vs_get_workspace_size(const w2vs_config* cfg, int32_t B, int32_t L,
                                      int32_t main_ctx, int32_t right_ctx, size_t* bytes);
w2vs_status_t w2vs_encode(const w2vs_config* cfg, const void* d_packed_weights,
                          const w2vs_encode_args* args, void* d_workspace, size_t workspace_bytes,
                          void* stream);

/* ---- incremental (chunk-by-chunk) forward with cached left context ------------------------
 * Exact for pos_type=SIN and extractor_mode=LAYER_NORM; other modes
 * return W2VS_UNSUPPORTED.  B streams advance in lock-step.  Host-side bookkeeping lives in the
 * caller-owned `host_state` blob, device-side caches in `d_state`. */
w2vs_status_t w2vs_stream_state_size(const w2vs_config* cfg, int32_t B, int32_t max_frames,
                                     int32_t max_new_samples, int32_t main_ctx, int32_t right_ctx,
                                     size_t* host_bytes, size_t* device_bytes,
                                     size_t* workspace_bytes);
w2vs_status_t w2vs_stream_init(const w2vs_config* cfg, int32_t B, int32_t max_frames,
                               int32_t max_new_samples, int32_t main_ctx, int32_t right_ctx,
                               void* host_state, size_t host_bytes, void* d_state,
                               size_t device_bytes, void* stream);
/* Feed n_new samples per stream (d_new_samples [B, n_new], dtype wav_dtype).  Emits every frame
 * that became final (whole blocks of main_ctx frames whose right_ctx look-ahead frames exist):
 * *n_out frames written to d_out_frames [n_out, B, D] (TBD, dtype cfg.dtype).  `flush`:
 *   W2VS_FLUSH_NONE   only final frames;
 *   W2VS_FLUSH_FINAL  end of stream: also emits the trailing frames exactly as the offline encoder
 *                     computes them on the complete utterance (rain finished=True); closes the stream;
 *   W2VS_FLUSH_PEEK   also emits the trailing frames as the offline encoder computes them on the
 *                     CURRENT prefix (the reference driver's prefix re-encoding, is_infer=True), without
 *                     committing them: the next call recomputes them with more context. */
#define W2VS_FLUSH_NONE 0
#define W2VS_FLUSH_FINAL 1
#define W2VS_FLUSH_PEEK 2
w2vs_status_t w2vs_stream_step(const w2vs_config* cfg, const void* d_packed_weights,
                               void* host_state, void* d_state, const void* d_new_samples,
                               int32_t wav_dtype, int32_t n_new, int32_t flush,
                               void* d_out_frames, int32_t out_capacity_frames, int32_t* n_out,
                               void* d_workspace, size_t workspace_bytes, void* stream);

/* Move a live stream into larger state buffers (the K/V cache and the frame buffer are sized by max_frames): copies
 * the conv carries, the projected frames and the cached K/V on `stream`, and the bookkeeping into host_state_new.
 * The new buffers are sized by w2vs_stream_state_size(..., new_max_frames, ...); cfg.sin_rows must cover
 * new_max_frames + 2 (re-pack the weights with a longer sinusoidal table first if needed).  The old buffers may be
 * released once the copies have completed (stream order). */
w2vs_status_t w2vs_stream_grow(const w2vs_config* cfg, const void* host_state_old, const void* d_state_old,
                               int32_t new_max_frames, void* host_state_new, size_t host_bytes,
                               void* d_state_new, size_t device_bytes, void* stream);

/* Host-side counters of a stream: samples consumed, frames produced by the conv stack, and frames of
 * committed (final) blocks.  Any pointer may be NULL. */
w2vs_status_t w2vs_stream_info(const void* host_state, int64_t* samples, int32_t* frames,
                               int32_t* final_frames);

/* ---- single operators (unit-test surface for the kernels; same code the forward uses) ------ */
typedef enum {
  W2VS_GEMM_AUTO = 0,
  W2VS_GEMM_SIMT = 1,          /* fp32-accurate CUDA-core kernel */
  /* 2 was the first one-CTA tcgen05 kernel, retired in ABI version 2 */
  W2VS_GEMM_TCGEN05_2CTA = 3,  /* tcgen05 cta_group::2 CTA pairs, TMA-store epilogue (default for bf16) */
  W2VS_GEMM_SKINNY = 4         /* M <= 64: weight-streaming mma.sync kernel (default for incremental steps) */
} w2vs_gemm_impl_t;
#define W2VS_EPI_GELU 1
#define W2VS_EPI_SPLITK 2   /* in-place fp32 product of a few hundred rows: the kernel may split K over several CTA pairs and
                              add the partial tiles at L2 (sums in arrival order: reproducible to rounding, not bit for
                              bit).  Set by the incremental steps of a batch of streams only; the full-utterance
                              forward keeps fixed reduction orders. */
/* C[M,N] = A[M,K] . W[N,K]^T + bias (+GELU) (+residual fp32, may alias C when C is fp32).
 * dtype_ab = dtype of A and W; dtype_c = dtype of C. lda/ldc in elements; lda may be < K*...
 * (overlapping rows: the strided-conv-as-GEMM view). */
w2vs_status_t w2vs_op_gemm(int32_t impl, int32_t dtype_ab, int32_t dtype_c, const void* d_A,
                           int64_t lda, const void* d_W, const float* d_bias,
                           const float* d_residual, void* d_C, int64_t ldc, int32_t M, int32_t N,
                           int32_t K, int32_t epilogue_flags, void* stream);
/* Row LayerNorm over N (eps 1e-5): y = LN(x)*gamma+beta, optional GELU; writes fp32 and/or
 * `dtype_act` copies. */
w2vs_status_t w2vs_op_layernorm(int32_t dtype_in, const void* d_x, int64_t ldx, const float* d_gamma,
                                const float* d_beta, float* d_out_f32, int32_t dtype_act,
                                void* d_out_act, int64_t ldo, int32_t rows, int32_t N,
                                int32_t gelu, void* stream);
/* Block-masked attention over a token buffer qkv [B, M, 3D] (q|k|v), M = T' + (T'/main)*rc,
 * key padding [B, M] (1 = masked with -inf).  ctx [B, M, D].
 * impl: 0 auto, 1 SIMT fp32-accumulate, 2 mma.sync flash kernel, 3 tcgen05/TMEM kernel (bf16). */
w2vs_status_t w2vs_op_attention(int32_t impl, int32_t dtype, const void* d_qkv,
                                const uint8_t* d_keypad, void* d_ctx, int32_t B, int32_t T_pad,
                                int32_t main_ctx, int32_t right_ctx, int32_t heads, int32_t D,
                                void* stream);

const char* w2vs_status_string(int32_t status);
/* Last CUDA error string recorded on this host thread by a failing call (diagnostics only). */
const char* w2vs_last_error(void);
/* Number of kernel launches issued by this host thread since the last reset (bench accounting). */
int64_t w2vs_launch_count(int32_t reset);

/* Per-launch device timing for bench.py (diagnostics; off by default).  w2vs_prof_enable(1, stream)
 * starts recording one CUDA event after every launch issued by this host thread; w2vs_prof_collect
 * waits for them and writes "kernel_name milliseconds launches" lines. */
void w2vs_prof_enable(int32_t on, void* stream);
int64_t w2vs_prof_collect(char* buf, int64_t capacity);

/* Diagnostics of the persistent incremental-step kernel (synchronising; tools and tests only).
 * w2vs_debug_fused_trace: globaltimer timestamps (ns) of the last launch on the current device,
 *   [2 CTAs: first, last][64 layers][12 events]; n = number of 64-bit words to copy.
 * w2vs_debug_fault_flags: bit 0 tcgen05 GEMM, bit 1 tcgen05 attention, bit 2 fused step kernel, bit 3 cluster step kernel -- set when a
 *   pipeline / barrier wait timed out inside a kernel (a bug; the kernels end instead of hanging). */
w2vs_status_t w2vs_debug_fused_trace(uint64_t* out, int32_t n);
/* same for the cluster step kernel: CTA 0, [64 layers][32 events].  The stamps are off by default:
 * w2vs_debug_cluster_trace(NULL, 1) switches them on for the process, (NULL, 0) off. */
w2vs_status_t w2vs_debug_cluster_trace(uint64_t* out, int32_t n);
w2vs_status_t w2vs_debug_fault_flags(int32_t* flags);

#ifdef __cplusplus
}
#endif
#endif /* W2VS_H_ */
