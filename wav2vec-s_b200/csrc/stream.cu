// Incremental (chunk-by-chunk) encoder forward with cached left context: w2vs_stream_*.
//
// The reference never runs its encoder incrementally: the SimulEval driver re-encodes the whole
// prefix at every decision step (rain/simul/transducer_agent.py:138-167,
// rain/simul/transducer_searcher.py:702-760).  The block mask (wav2vec_S.py:444-489) makes that
// equivalent to the schedule below, which is what this file implements (template: the fbank CAAT
// encoder's forward_infer/rollback_steps, rain/layers/unidirect_encoder.py:673-785):
//   * conv stack: per layer, carry the k-s (+remainder) trailing input rows to the next call;
//   * positions: absolute sinusoidal index frame+2 (no padding inside a live stream);
//   * block b becomes final once frames [main*b, main*(b+1)+rc) exist: its main + rc frames are run
//     as main+rc query tokens against the cached K/V of the main frames of blocks < b plus their own
//     K/V (no mask needed); only the main frames' K/V stay in the cache (the rc rows are overwritten
//     by the next block), and the main frames' outputs are emitted;
//   * flush (FINAL): the last full block runs with however many rc frames exist, then the trailing
//     partial block with none -- exactly what the offline mask gives on the complete utterance;
//   * flush (PEEK): the same tail computation without committing any state -- reproduces the
//     reference's output on a prefix that is not block aligned.
// Exact for pos_type=sin and extractor_mode=layer_norm (per-frame norms); other modes are
// non-causal by construction and return W2VS_UNSUPPORTED.
#include <string.h>
#include "common.cuh"
#include "kernels.h"
#include "layout.h"

// tcgen05 products of a decision step launched as programmatic dependents: 0 off, 1 only behind a LayerNorm, 2 all
// (16 streams, p50 per step: 2.36 / 2.22 / 2.10 ms)
#ifndef W2VS_STREAM_GEMM_PDL
#define W2VS_STREAM_GEMM_PDL 2
#endif

using namespace w2vs;

namespace {

constexpr uint32_t kMagic = 0x57325653u;  // "W2VS"

struct StreamHost {
  uint32_t magic;
  int32_t B, max_frames, max_new, main_ctx, rc;
  int64_t samples_total;
  int32_t carry[W2VS_MAX_CONV];      // rows carried at in_i[cur][carry_off ..]
  int32_t carry_off[W2VS_MAX_CONV];
  int32_t cur[W2VS_MAX_CONV];
  int32_t frames_total, blocks_done, finished;
};

struct StreamLayout {
  int cap_in[W2VS_MAX_CONV];    // rows (samples for layer 0) per stream of the layer's input buffer
  int out_cap[W2VS_MAX_CONV];   // rows per stream of the layer's output scratch
  int new_max[W2VS_MAX_CONV];
  int fcap, kv_rows, ntok_max;
  // device state
  size_t in[W2VS_MAX_CONV][2], fbuf, kv, kv_layer_bytes, sync, sync_bytes, fused_bar, dev_total;
  // workspace
  size_t out[2], normed, feats_tmp, x, xa, qkv, ctx, h, attn_part, ws_total;
};

inline bool conv_has_ln(const w2vs_config* cfg, int i) {
  return cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM && i < cfg->layer_norm_num;
}
template <typename T> inline T* at(void* base, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(base) + off);
}
template <typename T> inline const T* at(const void* base, size_t off) {
  return off == kNone ? nullptr : reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(base) + off);
}
struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; }
};

w2vs_status_t check_supported(const w2vs_config* cfg) {
  W2VS_TRY(validate_config(cfg));
  if (cfg->pos_type != W2VS_POS_SIN || cfg->extractor_mode != W2VS_EXTRACTOR_LAYER_NORM) {
    set_error("unsupported: incremental mode needs pos_type=sin and extractor_mode=layer_norm "
              "(conv positions / GroupNorm are non-causal)");
    return W2VS_UNSUPPORTED;
  }
  return W2VS_OK;
}

w2vs_status_t make_stream_layout(const w2vs_config* cfg, int B, int max_frames, int max_new, int main_ctx, int rc,
                                 StreamLayout* L) {
  W2VS_REQUIRE(B >= 1 && max_frames >= 1 && max_new >= 1, "B / max_frames / max_new_samples");
  W2VS_REQUIRE(main_ctx >= 1 && rc >= 0 && main_ctx + rc <= 64, "main_ctx + right_ctx must be <= 64");
  memset(L, 0, sizeof(*L));
  const int n = cfg->n_conv;
  const size_t as = act_size(cfg);
  int64_t new_prev = max_new;
  for (int i = 0; i < n; ++i) {
    const int k = cfg->conv_kernel[i], s = cfg->conv_stride[i];
    const int64_t n_in_max = (k - 1) + new_prev;
    W2VS_REQUIRE(n_in_max < (1 << 28), "max_new_samples too large");
    L->cap_in[i] = (int)((n_in_max + (i == 0 ? 7 : s - 1)) / (i == 0 ? 8 : s) * (i == 0 ? 8 : s));
    L->out_cap[i] = i == 0 ? (int)((n_in_max - k) / s + 1 > 1 ? (n_in_max - k) / s + 1 : 1) : L->cap_in[i] / s;
    L->new_max[i] = n_in_max >= k ? (int)((n_in_max - k) / s + 1) : 0;
    new_prev = L->new_max[i];
  }
  L->fcap = max_frames;
  L->kv_rows = max_frames + rc;
  L->ntok_max = main_ctx + rc;
  const int D = cfg->embed_dim, F = cfg->ffn_dim;
  Bump d;
  for (int i = 0; i < n; ++i) {
    const size_t row = i == 0 ? 4 : (size_t)cfg->conv_dim[i - 1] * as;
    const size_t bytes = ((size_t)B * L->cap_in[i] + 128) * row;
    L->in[i][0] = d.take(bytes);
    L->in[i][1] = d.take(bytes);
  }
  L->fbuf = d.take((size_t)B * L->fcap * D * 4);
  L->kv_layer_bytes = align_up((size_t)B * L->kv_rows * 2 * D * as, 256);
  L->kv = d.take(L->kv_layer_bytes * cfg->layers);
  // completion counters of the split-key step attention, one per (stream, head, query tile); zeroed by w2vs_stream_init
  L->sync_bytes = (size_t)B * cfg->heads * ((L->ntok_max + 63) / 64) * 4;
  L->sync = d.take(L->sync_bytes);
  // grid-barrier words of the fused step kernel (arrivals, departures): zero between launches
  L->fused_bar = d.take(64);
  L->dev_total = d.off;

  Bump w;
  size_t o[2] = {0, 0};
  for (int i = 0; i < n; ++i) {
    const size_t bytes = ((size_t)B * L->out_cap[i] + 128) * cfg->conv_dim[i] * as;
    if (bytes > o[i & 1]) o[i & 1] = bytes;
  }
  L->out[0] = w.take(o[0]);
  L->out[1] = w.take(o[1] ? o[1] : 256);
  const size_t last_rows = (size_t)B * L->out_cap[n - 1] + 128;
  L->normed = w.take(last_rows * cfg->conv_dim[n - 1] * as);
  L->feats_tmp = w.take(last_rows * D * 4);
  const size_t tok = (size_t)B * L->ntok_max + 128;
  L->x = w.take(tok * D * 4);
  L->xa = w.take(tok * D * as);
  L->qkv = w.take(tok * 3 * D * as);
  L->ctx = w.take(tok * D * as);
  L->h = w.take(tok * F * as);
  L->attn_part = w.take((size_t)B * cfg->heads * ((L->ntok_max + 63) / 64) * kAttnStepMaxSplits * 64 * 66 * 4);
  L->ws_total = w.off;
  return W2VS_OK;
}

// One block step: tokens = frames [f0, f0 + n_main + n_rc) of every stream; emits n_main frames.
w2vs_status_t block_step(const w2vs_config* cfg, const WeightLayout& wl, const StreamLayout& L, const void* W,
                         void* d_state, void* d_ws, int B, int f0, int n_main, int n_rc, void* out_frames,
                         cudaStream_t st) {
  const int adt = cfg->dtype, D = cfg->embed_dim, F = cfg->ffn_dim;
  const int ntok = n_main + n_rc, tokens = B * ntok;
  const size_t as = act_size(cfg);
  float* X = at<float>(d_ws, L.x);
  void* Xa = at<void>(d_ws, L.xa);
  void* qkv = at<void>(d_ws, L.qkv);
  void* ctx = at<void>(d_ws, L.ctx);
  void* h = at<void>(d_ws, L.h);
  const bool pre_ln = cfg->layer_norm_first != 0;
  if ((cfg->stream_step_impl == 0 || cfg->stream_step_impl == 3) && cfg->io_dtype == 0 && stream_cluster_applicable(cfg, B, ntok)) {
    // default where it applies (one stream, pre-LN bf16 model of an instantiated shape, at most 32 tokens per step, a
    // device that holds all clusters at once): the whole step as one kernel of thread-block clusters, two per
    // attention head; everything else takes the operator chain below
    StreamFusedArgs fa{};
    fa.cfg = cfg; fa.wl = &wl; fa.W = W;
    fa.B = B; fa.ntok = ntok; fa.n_main = n_main; fa.f0 = f0;
    fa.feats = at<float>(d_state, L.fbuf); fa.feat_rows = L.fcap;
    fa.R = X;
    fa.kv = at<void>(d_state, L.kv); fa.kv_layer_elems = (int64_t)(L.kv_layer_bytes / as); fa.kv_rows = L.kv_rows;
    fa.out = out_frames;
    fa.bar = at<unsigned long long>(d_state, L.fused_bar);
    return launch_stream_cluster(fa, st);
  }
  if (cfg->stream_step_impl == 2 && cfg->io_dtype == 0 && stream_fused_applicable(cfg, B, ntok)) {
    // opt-in (stream_step_impl = 2; bf16 models, at most 32 tokens per step): the whole step as one persistent
    // cooperative kernel
    StreamFusedArgs fa{};
    fa.cfg = cfg; fa.wl = &wl; fa.W = W;
    fa.B = B; fa.ntok = ntok; fa.n_main = n_main; fa.f0 = f0;
    fa.feats = at<float>(d_state, L.fbuf); fa.feat_rows = L.fcap;
    fa.R = X; fa.q = qkv; fa.ctx = ctx; fa.h = h;
    fa.kv = at<void>(d_state, L.kv); fa.kv_layer_elems = (int64_t)(L.kv_layer_bytes / as); fa.kv_rows = L.kv_rows;
    fa.partials = at<float>(d_ws, L.attn_part);
    fa.max_splits = ((L.ntok_max + 63) / 64) * kAttnStepMaxSplits * 64 / 32;
    fa.counters = at<unsigned>(d_state, L.sync);
    fa.out = out_frames;
    fa.bar = at<unsigned long long>(d_state, L.fused_bar);
    return launch_stream_fused(fa, st);
  }
  {
    EmbedArgs e{};
    e.feats = at<float>(d_state, L.fbuf) + (size_t)f0 * D; e.feat_rows = L.fcap;
    e.frame_pad = nullptr; e.pos = nullptr; e.pos_offset = f0 + 2;   // position = 1 + (1-based frame index)
    e.sin_table = at<float>(W, wl.sin_table); e.posconv = nullptr;
    e.gamma = pre_ln ? nullptr : at<float>(W, wl.enc_ln_w);
    e.beta = pre_ln ? nullptr : at<float>(W, wl.enc_ln_b);
    e.X = X; e.Xa = Xa; e.act_dtype = adt;
    e.B = B; e.T = ntok; e.T2 = ntok; e.M = ntok; e.main_ctx = ntok; e.rc = 1; e.D = D;
    W2VS_TRY(launch_embed(e, st));
  }
  auto layer_norm = [&](size_t gw, size_t gb, bool write_f32) {
    LayerNormArgs la{};
    la.x = X; la.in_dtype = W2VS_F32; la.ldx = D; la.gamma = at<float>(W, gw); la.beta = at<float>(W, gb);
    la.out_f32 = write_f32 ? X : nullptr; la.out_act = Xa; la.act_dtype = adt; la.ldo = D;
    la.rows = tokens; la.N = D; la.gelu = 0;
    return launch_layernorm(la, st);
  };
  auto gemm = [&](const void* A, int K, size_t w, size_t b, const float* res, void* C, int N, int cdt, int flags,
                  bool after_ln = false) {
    GemmArgs ga{};
    ga.pdl = (W2VS_STREAM_GEMM_PDL == 2 || (W2VS_STREAM_GEMM_PDL == 1 && after_ln)) ? 1 : 0;
    ga.A = A; ga.lda = K; ga.a_rows = tokens; ga.W = at<void>(W, w); ga.bias = at<float>(W, b);
    ga.residual = res; ga.C = C; ga.ldc = N; ga.M = tokens; ga.N = N; ga.K = K;
    ga.dtype_ab = adt; ga.dtype_c = cdt; ga.flags = flags;
    return launch_gemm(W2VS_GEMM_AUTO, ga, st);
  };
  for (int l = 0; l < cfg->layers; ++l) {
    const LayerW lw = layer_at(wl, l);
    void* cache = at<void>(d_state, L.kv + (size_t)l * L.kv_layer_bytes);
    if (pre_ln) W2VS_TRY(layer_norm(lw.ln1_w, lw.ln1_b, false));
    W2VS_TRY(gemm(Xa, D, lw.wqkv, lw.bqkv, nullptr, qkv, 3 * D, adt, 0, pre_ln));
    // (bf16: the step attention kernel appends this step's K / V to the cache itself)
    if (adt != W2VS_BF16) W2VS_TRY(launch_kv_append(qkv, cache, L.kv_rows, f0, ntok, D, (int)as, B, st));
    {
      AttnArgs aa{};
      aa.qkv = qkv; aa.ctx = ctx; aa.dtype = adt; aa.B = B; aa.heads = cfg->heads; aa.D = D;
      aa.n_step_q = ntok; aa.n_step_keys = f0 + ntok; aa.kv_cache = cache; aa.kv_rows = L.kv_rows;
      aa.step_partials = at<float>(d_ws, L.attn_part); aa.step_counters = at<unsigned>(d_state, L.sync);
      W2VS_TRY(launch_attention(0, aa, st));
    }
    W2VS_TRY(gemm(ctx, D, lw.wo, lw.bo, X, X, D, W2VS_F32, W2VS_EPI_SPLITK));
    if (pre_ln) W2VS_TRY(layer_norm(lw.ln2_w, lw.ln2_b, false));
    else W2VS_TRY(layer_norm(lw.ln1_w, lw.ln1_b, true));
    W2VS_TRY(gemm(Xa, D, lw.w1, lw.b1, nullptr, h, F, adt, W2VS_EPI_GELU, true));
    W2VS_TRY(gemm(h, F, lw.w2, lw.b2, X, X, D, W2VS_F32, W2VS_EPI_SPLITK));
    if (!pre_ln) W2VS_TRY(layer_norm(lw.ln2_w, lw.ln2_b, true));
  }
  FinalizeArgs f{};
  f.X = X; f.gamma = pre_ln ? at<float>(W, wl.enc_ln_w) : nullptr; f.beta = pre_ln ? at<float>(W, wl.enc_ln_b) : nullptr;
  f.out = out_frames; f.out_dtype = cfg->io_dtype == W2VS_F16 ? W2VS_F16 : adt; f.B = B; f.T_out = n_main; f.in_rows_per_utt = ntok; f.D = D; f.tbd = 1;
  return launch_finalize(f, st);
}

}  // namespace

extern "C" {

w2vs_status_t w2vs_stream_state_size(const w2vs_config* cfg, int32_t B, int32_t max_frames, int32_t max_new,
                                     int32_t main_ctx, int32_t rc, size_t* host_bytes, size_t* device_bytes,
                                     size_t* workspace_bytes) {
  W2VS_TRY(check_supported(cfg));
  W2VS_REQUIRE(host_bytes && device_bytes && workspace_bytes, "NULL pointer");
  StreamLayout L;
  W2VS_TRY(make_stream_layout(cfg, B, max_frames, max_new, main_ctx, rc, &L));
  *host_bytes = sizeof(StreamHost);
  *device_bytes = L.dev_total;
  *workspace_bytes = L.ws_total;
  return W2VS_OK;
}

w2vs_status_t w2vs_stream_init(const w2vs_config* cfg, int32_t B, int32_t max_frames, int32_t max_new,
                               int32_t main_ctx, int32_t rc, void* host_state, size_t host_bytes, void* d_state,
                               size_t device_bytes, void* stream) {
  W2VS_TRY(check_supported(cfg));
  W2VS_REQUIRE(host_state && d_state, "NULL pointer");
  StreamLayout L;
  W2VS_TRY(make_stream_layout(cfg, B, max_frames, max_new, main_ctx, rc, &L));
  if (host_bytes < sizeof(StreamHost) || device_bytes < L.dev_total) {
    set_error("stream state buffers too small");
    return W2VS_WORKSPACE_TOO_SMALL;
  }
  if (cfg->sin_rows < max_frames + 2) {
    set_error("invalid value: sinusoidal table (%d rows) too short for max_frames %d", cfg->sin_rows, max_frames);
    return W2VS_INVALID_VALUE;
  }
  StreamHost* hs = reinterpret_cast<StreamHost*>(host_state);
  memset(hs, 0, sizeof(*hs));
  hs->magic = kMagic;
  hs->B = B; hs->max_frames = max_frames; hs->max_new = max_new; hs->main_ctx = main_ctx; hs->rc = rc;
  // every data row is written before it is read; only the attention kernel's completion counters need zeroing
  // (L.sync and L.fused_bar are adjacent allocations: one memset covers both)
  cudaError_t e = cudaMemsetAsync(at<uint8_t>(d_state, L.sync), 0, L.fused_bar + 64 - L.sync, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("stream init memset: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  return W2VS_OK;
}

w2vs_status_t w2vs_stream_grow(const w2vs_config* cfg, const void* host_state_old, const void* d_state_old,
                               int32_t new_max_frames, void* host_state_new, size_t host_bytes, void* d_state_new,
                               size_t device_bytes, void* stream) {
  W2VS_TRY(check_supported(cfg));
  W2VS_REQUIRE(host_state_old && d_state_old && host_state_new && d_state_new, "NULL pointer");
  const StreamHost* ho = reinterpret_cast<const StreamHost*>(host_state_old);
  W2VS_REQUIRE(ho->magic == kMagic, "host_state was not initialised by w2vs_stream_init");
  W2VS_REQUIRE(new_max_frames >= ho->max_frames, "new_max_frames must not shrink the stream");
  W2VS_REQUIRE(cfg->sin_rows >= new_max_frames + 2, "sinusoidal table too short for new_max_frames (re-pack with more rows)");
  StreamLayout Lo, Ln;
  W2VS_TRY(make_stream_layout(cfg, ho->B, ho->max_frames, ho->max_new, ho->main_ctx, ho->rc, &Lo));
  W2VS_TRY(make_stream_layout(cfg, ho->B, new_max_frames, ho->max_new, ho->main_ctx, ho->rc, &Ln));
  if (host_bytes < sizeof(StreamHost) || device_bytes < Ln.dev_total) {
    set_error("stream state buffers too small");
    return W2VS_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int B = ho->B, D = cfg->embed_dim;
  const size_t as = act_size(cfg);
  auto fail = [](cudaError_t e) { set_error("stream grow copy: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; };
  cudaError_t e;
  // conv carries (their size depends on max_new_samples only): both halves of every double buffer
  for (int i = 0; i < cfg->n_conv; ++i)
    for (int h = 0; h < 2; ++h) {
      const size_t row = i == 0 ? 4 : (size_t)cfg->conv_dim[i - 1] * as;
      e = cudaMemcpyAsync(at<uint8_t>(d_state_new, Ln.in[i][h]), at<uint8_t>(d_state_old, Lo.in[i][h]),
                          ((size_t)B * Lo.cap_in[i] + 128) * row, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return fail(e);
    }
  // projected frames [B][fcap][D] fp32 and the K/V cache [layers][B][kv_rows][2D]: row pitch per stream changes
  if (ho->frames_total > 0) {
    e = cudaMemcpy2DAsync(at<uint8_t>(d_state_new, Ln.fbuf), (size_t)Ln.fcap * D * 4, at<uint8_t>(d_state_old, Lo.fbuf),
                          (size_t)Lo.fcap * D * 4, (size_t)ho->frames_total * D * 4, B, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(e);
  }
  const size_t kv_row = (size_t)2 * D * as;
  for (int l = 0; l < cfg->layers; ++l) {
    e = cudaMemcpy2DAsync(at<uint8_t>(d_state_new, Ln.kv + (size_t)l * Ln.kv_layer_bytes), (size_t)Ln.kv_rows * kv_row,
                          at<uint8_t>(d_state_old, Lo.kv + (size_t)l * Lo.kv_layer_bytes), (size_t)Lo.kv_rows * kv_row,
                          (size_t)Lo.kv_rows * kv_row, B, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(e);
  }
  e = cudaMemsetAsync(at<uint8_t>(d_state_new, Ln.sync), 0, Ln.fused_bar + 64 - Ln.sync, st);
  if (e != cudaSuccess) return fail(e);
  StreamHost hn = *ho;
  hn.max_frames = new_max_frames;
  *reinterpret_cast<StreamHost*>(host_state_new) = hn;
  return W2VS_OK;
}

w2vs_status_t w2vs_stream_info(const void* host_state, int64_t* samples, int32_t* frames, int32_t* final_frames) {
  W2VS_REQUIRE(host_state != nullptr, "host_state is NULL");
  const StreamHost* hs = reinterpret_cast<const StreamHost*>(host_state);
  W2VS_REQUIRE(hs->magic == kMagic, "host_state was not initialised by w2vs_stream_init");
  if (samples) *samples = hs->samples_total;
  if (frames) *frames = hs->frames_total;
  if (final_frames) *final_frames = hs->blocks_done * hs->main_ctx;
  return W2VS_OK;
}

w2vs_status_t w2vs_stream_step(const w2vs_config* cfg, const void* d_packed, void* host_state, void* d_state,
                               const void* d_new, int32_t wav_dtype, int32_t n_new, int32_t flush,
                               void* d_out_frames, int32_t out_capacity, int32_t* n_out, void* d_ws,
                               size_t ws_bytes, void* stream) {
  W2VS_TRY(check_supported(cfg));
  W2VS_REQUIRE(d_packed && host_state && d_state && d_ws && n_out, "NULL pointer");
  // a decision step is a chain of ~200 short kernels: launch the PDL-aware ones (common.cuh) as programmatic
  // dependents so that their launch latency hides under their predecessors
  struct PdlScope { PdlScope() { g_pdl_on = true; } ~PdlScope() { g_pdl_on = false; } } pdl_scope;
  // The step works on a copy of the host-side bookkeeping and commits it only after its last launch succeeded: a
  // launch error in the middle of the chain leaves `host_state` describing the state before the call.  The device
  // side tolerates that: conv carries are double buffered (the step writes the "next" buffers), frames and K/V rows
  // are appended behind the committed counts, so repeating the call overwrites what the failed one left.
  StreamHost* const hs_committed = reinterpret_cast<StreamHost*>(host_state);
  W2VS_REQUIRE(hs_committed->magic == kMagic, "host_state was not initialised by w2vs_stream_init");
  StreamHost hs_work = *hs_committed;
  StreamHost* const hs = &hs_work;
  W2VS_REQUIRE(!hs->finished, "stream already finished");
  W2VS_REQUIRE(n_new >= 0 && n_new <= hs->max_new, "n_new exceeds max_new_samples");
  W2VS_REQUIRE(n_new == 0 || d_new != nullptr, "d_new_samples is NULL");
  W2VS_REQUIRE(wav_dtype == W2VS_F32 || wav_dtype == W2VS_BF16 || wav_dtype == W2VS_I16, "wav_dtype");
  W2VS_REQUIRE(flush == W2VS_FLUSH_NONE || flush == W2VS_FLUSH_FINAL || flush == W2VS_FLUSH_PEEK, "flush mode");
  const int B = hs->B, main_ctx = hs->main_ctx, rc = hs->rc, n = cfg->n_conv;
  StreamLayout L;
  W2VS_TRY(make_stream_layout(cfg, B, hs->max_frames, hs->max_new, main_ctx, rc, &L));
  if (ws_bytes < L.ws_total) { set_error("workspace too small: %zu < %zu", ws_bytes, L.ws_total); return W2VS_WORKSPACE_TOO_SMALL; }
  WeightLayout wl;
  make_weight_layout(cfg, &wl);

  // ---- plan (pure host arithmetic; nothing is launched until every check has passed) ------------
  int n_in[W2VS_MAX_CONV], n_o[W2VS_MAX_CONV];
  int new_rows = n_new;
  for (int i = 0; i < n; ++i) {
    const int k = cfg->conv_kernel[i], s = cfg->conv_stride[i];
    n_in[i] = hs->carry[i] + new_rows;
    n_o[i] = n_in[i] >= k ? (n_in[i] - k) / s + 1 : 0;
    new_rows = n_o[i];
  }
  const int f_new = new_rows;
  const int frames = hs->frames_total + f_new;
  if (frames > hs->max_frames) { set_error("invalid value: stream exceeds max_frames (%d > %d)", frames, hs->max_frames); return W2VS_INVALID_VALUE; }
  int blocks = hs->blocks_done, emit = 0;
  while (frames >= (blocks + 1) * main_ctx + rc) { ++blocks; emit += main_ctx; }
  const int tail = frames - blocks * main_ctx;   // < main + rc
  if (flush != W2VS_FLUSH_NONE) emit += tail;
  if (emit > out_capacity) { set_error("invalid value: output capacity %d < %d frames", out_capacity, emit); return W2VS_INVALID_VALUE; }
  if (emit > 0) W2VS_REQUIRE(d_out_frames != nullptr, "d_out_frames is NULL");

  cudaStream_t st = (cudaStream_t)stream;
  const int adt = cfg->dtype, D = cfg->embed_dim;
  const size_t as = act_size(cfg);
  const void* W = d_packed;

  // ---- conv stack on the new samples ---------------------------------------------------------------
  if (n_new > 0) {
    const void* prev_out = nullptr;   // output rows of layer i-1
    int prev_cap = 0;
    for (int i = 0; i < n; ++i) {
      const int k = cfg->conv_kernel[i], s = cfg->conv_stride[i];
      const int cin = i == 0 ? 1 : cfg->conv_dim[i - 1], cout = cfg->conv_dim[i];
      const int added = i == 0 ? n_new : n_o[i - 1];
      if (added == 0) break;   // nothing new reaches this layer; its carry is unchanged
      const int cur = hs->cur[i], nxt = cur ^ 1;
      void* dst = at<void>(d_state, L.in[i][nxt]);
      const void* old = at<void>(d_state, L.in[i][cur]);
      if (i > 0 && B == 1 && adt == W2VS_BF16 && cfg->stream_step_impl != 1 && conv_has_ln(cfg, i) && cin == cout && n_o[i] > 0) {
        // one stream: the whole block (carry + new rows -> conv -> LayerNorm -> GELU, next carry) as one launch
        ConvStepArgs ca{};
        ca.carry = reinterpret_cast<const uint8_t*>(old) + (size_t)hs->carry_off[i] * cin * as; ca.n_carry = hs->carry[i];
        ca.fresh = prev_out; ca.n_fresh = added;
        ca.W = at<void>(W, wl.conv[i].w); ca.bias = at<float>(W, wl.conv[i].bias);
        ca.gamma = at<float>(W, wl.conv[i].norm_w); ca.beta = at<float>(W, wl.conv[i].norm_b);
        ca.out = at<void>(d_ws, L.out[i & 1]); ca.n_out = n_o[i]; ca.carry_out = dst; ca.k = k; ca.s = s; ca.C = cout;
        if (conv_step_applicable(ca)) {
          W2VS_TRY(launch_conv_step(ca, st));
          hs->cur[i] = nxt;
          hs->carry_off[i] = 0;
          hs->carry[i] = n_in[i] - n_o[i] * s;
          prev_out = ca.out;
          prev_cap = L.out_cap[i];
          continue;
        }
      }
      if (i == 0) {
        W2VS_TRY(launch_concat_wav((float*)dst, L.cap_in[0], (const float*)old, L.cap_in[0], hs->carry_off[0],
                                   hs->carry[0], d_new, wav_dtype, n_new, n_new, B, st));
      } else {
        const size_t row = (size_t)cin * as;
        W2VS_TRY(launch_concat_rows(dst, (int64_t)L.cap_in[i] * row, old, (int64_t)L.cap_in[i] * row, hs->carry_off[i],
                                    hs->carry[i], prev_out, (int64_t)prev_cap * row, added, (int)row, B, st));
      }
      hs->cur[i] = nxt;
      hs->carry_off[i] = n_o[i] * s;
      hs->carry[i] = n_in[i] - n_o[i] * s;
      if (n_o[i] == 0) { hs->carry_off[i] = 0; hs->carry[i] = n_in[i]; break; }
      void* out = at<void>(d_ws, L.out[i & 1]);
      if (i == 0) {
        Conv0Args c{};
        c.wav = dst; c.wav_dtype = W2VS_F32; c.wav_ld = L.cap_in[0];
        c.w = at<float>(W, wl.conv[0].w); c.bias = at<float>(W, wl.conv[0].bias);
        c.gamma = at<float>(W, wl.conv[0].norm_w); c.beta = at<float>(W, wl.conv[0].norm_b);
        c.out = out; c.out_dtype = adt; c.B = B; c.T0 = n_o[0]; c.rows_per_utt = L.out_cap[0];
        c.C = cout; c.k = k; c.stride = s;
        c.norm = conv_has_ln(cfg, 0) ? CONV0_NORM_LAYER : CONV0_NORM_NONE;
        W2VS_TRY(launch_conv0(c, st));
      } else {
        const bool ln = conv_has_ln(cfg, i);
        GemmArgs ga{};
        ga.A = dst; ga.lda = (int64_t)s * cin; ga.a_rows = (int64_t)B * L.out_cap[i] + 1;
        ga.W = at<void>(W, wl.conv[i].w); ga.bias = at<float>(W, wl.conv[i].bias); ga.residual = nullptr;
        ga.M = (B - 1) * L.out_cap[i] + n_o[i]; ga.N = cout; ga.K = k * cin; ga.dtype_ab = adt; ga.ldc = cout;
        ga.pdl = W2VS_STREAM_GEMM_PDL == 2;
        if (ln) {
          ga.C = out; ga.dtype_c = adt; ga.flags = 0;
          W2VS_TRY(launch_gemm(W2VS_GEMM_AUTO, ga, st));
          LayerNormArgs la{};
          la.x = out; la.in_dtype = adt; la.ldx = cout;
          la.gamma = at<float>(W, wl.conv[i].norm_w); la.beta = at<float>(W, wl.conv[i].norm_b);
          la.out_f32 = nullptr; la.out_act = out; la.act_dtype = adt; la.ldo = cout;
          la.rows = ga.M; la.N = cout; la.gelu = 1;
          W2VS_TRY(launch_layernorm(la, st));
        } else {
          ga.C = out; ga.dtype_c = adt; ga.flags = W2VS_EPI_GELU;
          W2VS_TRY(launch_gemm(W2VS_GEMM_AUTO, ga, st));
        }
      }
      prev_out = out;
      prev_cap = L.out_cap[i];
    }
  }
  // ---- feature LayerNorm + post_extract_proj on the new frames, appended to the frame buffer ----------
  if (f_new > 0) {
    const int CL = cfg->conv_dim[n - 1], ocap = L.out_cap[n - 1];
    const int rows = (B - 1) * ocap + f_new;
    const void* conv_out = at<void>(d_ws, L.out[(n - 1) & 1]);
    float* feats_tmp = at<float>(d_ws, L.feats_tmp);
    LayerNormArgs la{};
    la.x = conv_out; la.in_dtype = adt; la.ldx = CL;
    la.gamma = at<float>(W, wl.feat_ln_w); la.beta = at<float>(W, wl.feat_ln_b);
    la.rows = rows; la.N = CL; la.gelu = 0; la.ldo = CL; la.act_dtype = adt;
    bool fused_proj = false;
    if (CL != D && B == 1 && adt == W2VS_BF16 && cfg->stream_step_impl != 1) {
      // one stream: LayerNorm + projection + append to the frame buffer as one launch
      FeatProjArgs fp{};
      fp.x = conv_out; fp.rows = f_new; fp.K = CL; fp.gamma = la.gamma; fp.beta = la.beta;
      fp.W = at<void>(W, wl.proj_w); fp.bias = at<float>(W, wl.proj_b);
      fp.out = at<float>(d_state, L.fbuf) + (size_t)hs->frames_total * D; fp.N = D;
      if (feat_proj_applicable(fp)) {
        W2VS_TRY(launch_feat_proj(fp, st));
        fused_proj = true;
      }
    }
    if (fused_proj) {
    } else if (CL != D) {
      void* normed = at<void>(d_ws, L.normed);
      la.out_f32 = nullptr; la.out_act = normed;
      W2VS_TRY(launch_layernorm(la, st));
      GemmArgs ga{};
      ga.A = normed; ga.lda = CL; ga.a_rows = rows; ga.W = at<void>(W, wl.proj_w); ga.bias = at<float>(W, wl.proj_b);
      ga.residual = nullptr; ga.C = feats_tmp; ga.ldc = D; ga.M = rows; ga.N = D; ga.K = CL;
      ga.dtype_ab = adt; ga.dtype_c = W2VS_F32; ga.flags = 0; ga.pdl = W2VS_STREAM_GEMM_PDL == 2;
      W2VS_TRY(launch_gemm(W2VS_GEMM_AUTO, ga, st));
    } else {
      la.out_f32 = feats_tmp; la.out_act = nullptr;
      W2VS_TRY(launch_layernorm(la, st));
    }
    float* fbuf = at<float>(d_state, L.fbuf);
    if (!fused_proj)
    W2VS_TRY(launch_concat_rows(fbuf + (size_t)hs->frames_total * D, (int64_t)L.fcap * D * 4, feats_tmp, 0, 0, 0,
                                feats_tmp, (int64_t)ocap * D * 4, f_new, D * 4, B, st));
  }
  hs->samples_total += n_new;
  hs->frames_total = frames;

  // ---- blocks that became final -------------------------------------------------------------------------
  uint8_t* outp = reinterpret_cast<uint8_t*>(d_out_frames);
  const size_t frame_bytes = (size_t)B * D * as;
  int written = 0;
  while (hs->blocks_done < blocks) {
    W2VS_TRY(block_step(cfg, wl, L, W, d_state, d_ws, B, hs->blocks_done * main_ctx, main_ctx, rc,
                        outp + (size_t)written * frame_bytes, st));
    written += main_ctx;
    ++hs->blocks_done;
  }
  // ---- tail: final flush (commits) or peek (leaves the state untouched) -------------------------------------
  if (flush != W2VS_FLUSH_NONE && tail > 0) {
    int f0 = blocks * main_ctx, rem = tail;
    if (rem >= main_ctx) {   // a last full block with the look-ahead frames that exist
      W2VS_TRY(block_step(cfg, wl, L, W, d_state, d_ws, B, f0, main_ctx, rem - main_ctx,
                          outp + (size_t)written * frame_bytes, st));
      written += main_ctx;
      f0 += main_ctx;
      rem -= main_ctx;
    }
    if (rem > 0) {           // trailing partial block: no look-ahead
      W2VS_TRY(block_step(cfg, wl, L, W, d_state, d_ws, B, f0, rem, 0, outp + (size_t)written * frame_bytes, st));
      written += rem;
    }
  }
  if (flush == W2VS_FLUSH_FINAL) hs->finished = 1;
  *hs_committed = hs_work;
  *n_out = written;
  return W2VS_OK;
}

}  // extern "C"
