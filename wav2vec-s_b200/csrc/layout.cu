// Host-side geometry / layout (see layout.h).  Compiled by nvcc but contains no device code.
#include <stdarg.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>
#include "common.cuh"
#include "layout.h"

namespace w2vs {

thread_local char g_last_error[512] = {0};
thread_local int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

// ---- per-launch device timing ---------------------------------------------------------------------
thread_local bool g_prof_on = false;
thread_local bool g_pdl_on = false;
namespace {
struct ProfState {
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<std::string, int>> marks;  // (kernel name, event index); index 0 = base
  int used = 0;
  cudaEvent_t get() {
    if (used == (int)pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
};
thread_local ProfState g_prof;
}  // namespace

void prof_mark(const char* what, cudaStream_t st) {
  cudaEvent_t e = g_prof.get();
  cudaEventRecord(e, st);
  g_prof.marks.emplace_back(what, g_prof.used - 1);
}

extern "C" void w2vs_prof_enable(int32_t on, void* stream) {
  g_prof.used = 0;
  g_prof.marks.clear();
  g_prof_on = on != 0;
  if (g_prof_on) prof_mark("(base)", (cudaStream_t)stream);
}

// Waits for the recorded events and writes "name ms count\n" lines (device time between consecutive
// launch completions on the stream, summed per kernel name).  Returns bytes needed (incl. NUL).
extern "C" int64_t w2vs_prof_collect(char* buf, int64_t cap) {
  std::map<std::string, std::pair<double, int>> acc;
  for (size_t i = 1; i < g_prof.marks.size(); ++i) {
    cudaEvent_t a = g_prof.pool[g_prof.marks[i - 1].second], b = g_prof.pool[g_prof.marks[i].second];
    if (cudaEventSynchronize(b) != cudaSuccess) break;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) break;
    auto& r = acc[g_prof.marks[i].first];
    r.first += ms;
    r.second += 1;
  }
  std::string out;
  char line[256];
  for (auto& kv : acc) {
    snprintf(line, sizeof(line), "%s %.6f %d\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

w2vs_status_t validate_config(const w2vs_config* cfg) {
  W2VS_REQUIRE(cfg != nullptr, "cfg is NULL");
  W2VS_REQUIRE(cfg->abi_version == W2VS_ABI_VERSION, "abi_version mismatch");
  W2VS_REQUIRE(cfg->dtype == W2VS_F32 || cfg->dtype == W2VS_BF16, "dtype");
  W2VS_REQUIRE(cfg->io_dtype == 0 || (cfg->io_dtype == W2VS_F16 && cfg->dtype == W2VS_BF16), "io_dtype (fp16 I/O runs on the bf16 path)");
  W2VS_REQUIRE(cfg->n_conv >= 1 && cfg->n_conv <= W2VS_MAX_CONV, "n_conv");
  for (int i = 0; i < cfg->n_conv; ++i) {
    W2VS_REQUIRE(cfg->conv_dim[i] >= 32 && cfg->conv_dim[i] % 32 == 0 && cfg->conv_dim[i] <= 1024,
                 "conv_dim must be a multiple of 32 in [32,1024]");
    W2VS_REQUIRE(cfg->conv_kernel[i] >= 1 && cfg->conv_stride[i] >= 1, "conv kernel/stride");
    W2VS_REQUIRE(cfg->conv_kernel[i] >= cfg->conv_stride[i], "conv kernel < stride is not supported");
  }
  W2VS_REQUIRE(cfg->conv_kernel[0] <= 16, "first conv kernel must be <= 16");
  W2VS_REQUIRE(cfg->conv_dim[0] % 64 == 0, "first conv dim must be a multiple of 64");
  W2VS_REQUIRE(cfg->extractor_mode == W2VS_EXTRACTOR_DEFAULT ||
               cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM, "extractor_mode");
  W2VS_REQUIRE(cfg->embed_dim >= 64 && cfg->embed_dim % 64 == 0 && cfg->embed_dim <= 2048, "embed_dim");
  W2VS_REQUIRE(cfg->ffn_dim >= 64 && cfg->ffn_dim % 64 == 0, "ffn_dim");
  W2VS_REQUIRE(cfg->heads >= 1 && cfg->embed_dim % cfg->heads == 0, "heads");
  W2VS_REQUIRE(cfg->layers >= 1, "layers");
  W2VS_REQUIRE(cfg->seq_multiple >= 1, "seq_multiple");
  W2VS_REQUIRE(cfg->pos_type == W2VS_POS_SIN || cfg->pos_type == W2VS_POS_CONV, "pos_type");
  if (cfg->embed_dim / cfg->heads != 64) {
    set_error("unsupported: head_dim %d (only 64, as in every released wav2vec-S model)",
              cfg->embed_dim / cfg->heads);
    return W2VS_UNSUPPORTED;
  }
  if (cfg->pos_type == W2VS_POS_CONV) {
    W2VS_REQUIRE(cfg->conv_pos >= 1 && cfg->conv_pos_groups >= 1 &&
                 cfg->embed_dim % cfg->conv_pos_groups == 0, "conv_pos / conv_pos_groups");
  } else {
    W2VS_REQUIRE(cfg->sin_rows >= 3, "sin_rows");
  }
  return W2VS_OK;
}

w2vs_status_t make_geometry(const w2vs_config* cfg, int L, int main_ctx, int rc, Geometry* g) {
  W2VS_REQUIRE(L >= 1, "L");
  W2VS_REQUIRE(main_ctx >= 1 && rc >= 0, "main_ctx/right_ctx");
  memset(g, 0, sizeof(*g));
  g->n_conv = cfg->n_conv;
  int64_t t = L;
  for (int i = 0; i < cfg->n_conv; ++i) {
    W2VS_REQUIRE(t >= cfg->conv_kernel[i],
                 "waveform shorter than the receptive field of the conv stack");
    t = (t - cfg->conv_kernel[i]) / cfg->conv_stride[i] + 1;
    g->conv_len[i] = (int)t;
  }
  const int n = cfg->n_conv;
  g->T = g->conv_len[n - 1];
  // rows[i-1] = stride_i * rows[i]; smallest rows[n-1] >= T such that rows[i] >= conv_len[i] for all i
  int64_t rl = g->T;
  for (;; ++rl) {
    int64_t r = rl;
    bool ok = true;
    for (int i = n - 1; i >= 0; --i) {
      if (r < g->conv_len[i]) { ok = false; break; }
      if (i > 0) r *= cfg->conv_stride[i];
    }
    if (ok) break;
  }
  int64_t r = rl;
  for (int i = n - 1; i >= 0; --i) {
    W2VS_REQUIRE(r < (1ll << 30), "utterance too long");
    g->conv_rows[i] = (int)r;
    if (i > 0) r *= cfg->conv_stride[i];
  }
  const int mult = cfg->seq_multiple;
  g->T2 = (g->T + mult - 1) / mult * mult;
  g->main_ctx = main_ctx;
  g->rc = rc;
  g->nb = g->T2 / main_ctx;
  g->R = rc > 0 ? g->nb * rc : 0;
  g->M = g->T2 + g->R;
  return W2VS_OK;
}

namespace {
struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};
}  // namespace

void make_weight_layout(const w2vs_config* cfg, WeightLayout* wl) {
  Bump b;
  const size_t as = act_size(cfg);
  int c_in = 1;
  for (int i = 0; i < W2VS_MAX_CONV; ++i) wl->conv[i] = ConvW{kNone, kNone, kNone, kNone};
  for (int i = 0; i < cfg->n_conv; ++i) {
    const int c = cfg->conv_dim[i], k = cfg->conv_kernel[i];
    ConvW& cw = wl->conv[i];
    cw.w = b.take((size_t)c * k * c_in * (i == 0 ? 4 : as));
    cw.bias = cfg->conv_bias ? b.take((size_t)c * 4) : kNone;
    const bool has_norm = (cfg->extractor_mode == W2VS_EXTRACTOR_LAYER_NORM && i < cfg->layer_norm_num) ||
                          (cfg->extractor_mode == W2VS_EXTRACTOR_DEFAULT && i == 0);
    cw.norm_w = has_norm ? b.take((size_t)c * 4) : kNone;
    cw.norm_b = has_norm ? b.take((size_t)c * 4) : kNone;
    c_in = c;
  }
  const int D = cfg->embed_dim, F = cfg->ffn_dim;
  wl->feat_ln_w = b.take((size_t)c_in * 4);
  wl->feat_ln_b = b.take((size_t)c_in * 4);
  if (c_in != D) {
    wl->proj_w = b.take((size_t)D * c_in * as);
    wl->proj_b = b.take((size_t)D * 4);
  } else {
    wl->proj_w = wl->proj_b = kNone;
  }
  wl->sin_table = wl->posconv_w = wl->posconv_b = wl->posconv_wg = kNone;
  if (cfg->pos_type == W2VS_POS_SIN) {
    wl->sin_table = b.take((size_t)cfg->sin_rows * D * 4);
  } else {
    wl->posconv_w = b.take((size_t)D * (D / cfg->conv_pos_groups) * cfg->conv_pos * 4);
    wl->posconv_b = b.take((size_t)D * 4);
    if (posconv_tc(cfg)) wl->posconv_wg = b.take((size_t)D * cfg->conv_pos * posconv_dgp(cfg) * 2);
  }
  wl->enc_ln_w = b.take((size_t)D * 4);
  wl->enc_ln_b = b.take((size_t)D * 4);
  wl->layers_begin = b.off;
  LayerW& l = wl->layer0;
  l.wqkv = b.take((size_t)3 * D * D * as);
  l.bqkv = b.take((size_t)3 * D * 4);
  l.wo = b.take((size_t)D * D * as);
  l.bo = b.take((size_t)D * 4);
  l.ln1_w = b.take((size_t)D * 4);
  l.ln1_b = b.take((size_t)D * 4);
  l.w1 = b.take((size_t)F * D * as);
  l.b1 = b.take((size_t)F * 4);
  l.w2 = b.take((size_t)D * F * as);
  l.b2 = b.take((size_t)D * 4);
  l.ln2_w = b.take((size_t)D * 4);
  l.ln2_b = b.take((size_t)D * 4);
  l.w2s = stream_fused_model(cfg) ? b.take((size_t)D * F * 2) : kNone;
  const size_t wc_bytes = stream_cluster_layer_bytes(cfg);
  l.wc = wc_bytes ? b.take(wc_bytes) : kNone;
  wl->layer_stride = b.off - wl->layers_begin;
  wl->total = wl->layers_begin + wl->layer_stride * (size_t)cfg->layers;
}

void make_workspace(const w2vs_config* cfg, const Geometry& g, int B, Workspace* ws) {
  Bump b;
  const size_t as = act_size(cfg);
  const int n = cfg->n_conv;
  // ping holds even layers' outputs, pong odd layers'; slack rows cover the k-s overlap that the
  // strided-GEMM view of the last output row reads past the end.
  size_t ping = 0, pong = 0;
  for (int i = 0; i < n; ++i) {
    size_t rows = (size_t)B * g.conv_rows[i] + 64;
    size_t bytes = rows * cfg->conv_dim[i] * as;
    if (i % 2 == 0) ping = bytes > ping ? bytes : ping; else pong = bytes > pong ? bytes : pong;
  }
  ws->conv_a = b.take(ping);
  ws->conv_b = b.take(pong ? pong : 256);
  ws->gn_stats = cfg->extractor_mode == W2VS_EXTRACTOR_DEFAULT
                     ? b.take((size_t)B * cfg->conv_dim[0] * 2 * 4 * (1 + (size_t)(g.conv_len[0] + 255) / 256))
                     : kNone;
  const int D = cfg->embed_dim, F = cfg->ffn_dim;
  ws->wav_stats = b.take((size_t)B * 2 * 4);
  ws->feats = b.take((size_t)B * g.conv_rows[n - 1] * D * 4);
  ws->frame_pad = b.take((size_t)B * g.T);
  ws->pos = b.take((size_t)B * g.T * 4);
  ws->keypad = b.take((size_t)B * g.M);
  ws->pad_blk = b.take((size_t)B * ((g.M + 127) / 128));
  const size_t tok = (size_t)B * g.M + 128;  // slack rows: tile tails of TMA loads stay in-bounds anyway
  ws->x = b.take(tok * D * 4);
  ws->xa = b.take(tok * D * as);
  ws->qkv = b.take(tok * 3 * D * as);
  ws->ctx = b.take(tok * D * as);
  ws->h = b.take(tok * F * as);
  ws->posconv_tmp = ws->posconv_xg = kNone;
  if (posconv_tc(cfg)) {
    const size_t tp = (size_t)g.T + cfg->conv_pos;
    ws->posconv_tmp = b.take(B * tp * D * 4);
    ws->posconv_xg = b.take((size_t)cfg->conv_pos_groups * (B * tp + cfg->conv_pos) * posconv_dgp(cfg) * 2);
  } else if (cfg->pos_type == W2VS_POS_CONV) {
    ws->posconv_tmp = b.take((size_t)B * g.T * D * 4);
  }
  ws->total = b.off;
}

}  // namespace w2vs
