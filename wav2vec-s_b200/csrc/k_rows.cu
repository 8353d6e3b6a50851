// Row-wise kernels (one warp per row of N <= 2048 channels, the row lives in registers):
//   * layernorm_rows  : LayerNorm (+GELU) -> fp32 and/or activation-dtype copies
//                       (Fp32LayerNorm of the conv blocks, wav2vec2.py:739-745; feature LayerNorm
//                        :556; self_attn_layer_norm / final_layer_norm :936-976)
//   * prep_masks      : frame padding mask, sinusoidal positions, extended key-padding mask
//                       (wav2vec2.py:560-565, utils.py:250-260, wav2vec_S.py:470-476) -- integer work,
//                       bit-exact
//   * embed_tokens    : zero padded frames, add positional embedding, optional encoder LayerNorm,
//                       pad to T', append look-ahead copies (wav2vec_S.py:355-389,483-484)
//   * finalize_rows   : drop look-ahead copies / padding, optional final LayerNorm, BTD or TBD
//                       (wav2vec_S.py:425-440, wav2vec2.py:828-834, rain :314-330)
#include "common.cuh"
#include "kernels.h"

namespace w2vs {

// Normalise a row held as v[NCH][8] (lane owns chunks lane + 32*j of 8 consecutive channels).
template <int NCH>
__device__ __forceinline__ void warp_layernorm(float (&v)[NCH][8], int N, int lane,
                                               const float* __restrict__ gamma,
                                               const float* __restrict__ beta) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j)
    if ((lane + 32 * j) * 8 < N) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[j][e];
    }
  const float mean = warp_sum(s) / (float)N;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j)
    if ((lane + 32 * j) * 8 < N) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[j][e] - mean; q = fmaf(d, d, q); }
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)N + 1e-5f);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < N) {
      float g[8], bt[8];
      load8(gamma + c0, g);
      load8(beta + c0, bt);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j][e] = (v[j][e] - mean) * rstd * g[e] + bt[e];
    }
  }
}

template <typename TIn, typename TAct, int NCH>
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const TIn* x, int64_t ldx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* out_f32, TAct* out_act, int64_t ldo,
                      int rows, int N, int gelu) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[NCH][8];
  const TIn* xr = x + (size_t)row * ldx;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < N) load8(xr + c0, v[j]);
  }
  warp_layernorm<NCH>(v, N, lane, gamma, beta);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < N) {
      if (gelu) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          if (out_f32) gelu_erf2(v[j][e], v[j][e + 1]); else gelu2<TAct>(v[j][e], v[j][e + 1]);
        }
      }
      if (out_f32) store8(out_f32 + (size_t)row * ldo + c0, v[j]);
      if (out_act) store8(out_act + (size_t)row * ldo + c0, v[j]);
    }
  }
}

// bf16 -> LayerNorm -> GELU -> bf16 (the conv blocks' Fp32LayerNorm + GELU, in place over the GEMM output): same
// arithmetic as the generic kernel (fp32, mean first, then centred squares) on packed f32x2 pairs -- 8.5 issued
// instructions per element instead of 14; the generic kernel was issue bound at 79 % on these passes (ncu).
// Each warp normalises LG_ROWS rows at once, all of their loads issued before the first reduction.  Measured: alone
// (ncu) the pass stays at 5.0 of 6.5 TB/s either way; inside the power-capped step the row kernels take 7.25 instead
// of 7.5 ms (same-box A/B), so two rows it is.
#ifndef W2VS_LN_GELU_ROWS
#define W2VS_LN_GELU_ROWS 2
#endif
constexpr int LG_ROWS = W2VS_LN_GELU_ROWS;
template <int NCH>
__global__ void __launch_bounds__(256, NCH <= 2 ? 4 : 2)
ln_gelu_bf16_kernel(const bf16* x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                    bf16* out, int64_t ldo, int rows) {
  pdl_prologue();
  constexpr int N = NCH * 256;
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * LG_ROWS;
  if (row0 >= rows) return;
  uint64_t p[LG_ROWS][NCH][4];
#pragma unroll
  for (int r = 0; r < LG_ROWS; ++r) {
    const int row = min(row0 + r, rows - 1);          // a tail warp recomputes the last row and drops the result
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (size_t)row * ldx + (lane + 32 * j) * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) p[r][j][e] = pack2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
    }
  }
  uint64_t rs2[LG_ROWS];
#pragma unroll
  for (int r = 0; r < LG_ROWS; ++r) {
    uint64_t s2 = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s2 = fadd2(s2, p[r][j][e]);
    float sa, sb;
    unpack2(s2, sa, sb);
    const float mean = warp_sum(sa + sb) * (1.0f / N);
    const uint64_t nm2 = pack2(-mean, -mean);
    uint64_t q2 = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { p[r][j][e] = fadd2(p[r][j][e], nm2); q2 = ffma2(p[r][j][e], p[r][j][e], q2); }
    unpack2(q2, sa, sb);
    const float rstd = 1.0f / sqrtf(warp_sum(sa + sb) * (1.0f / N) + 1e-5f);
    rs2[r] = pack2(rstd, rstd);
  }
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    const ulonglong2 g01 = *reinterpret_cast<const ulonglong2*>(gamma + c0), g23 = *reinterpret_cast<const ulonglong2*>(gamma + c0 + 4);
    const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(beta + c0), b23 = *reinterpret_cast<const ulonglong2*>(beta + c0 + 4);
    const uint64_t g[4] = {g01.x, g01.y, g23.x, g23.y}, bt[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int r = 0; r < LG_ROWS; ++r) {
      if (row0 + r >= rows) break;
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float y0, y1;
        unpack2(gelu_tanh2p(ffma2(p[r][j][e], fmul2(rs2[r], g[e]), bt[e])), y0, y1);
        o[e] = pack_bf16x2(y0, y1);
      }
      *reinterpret_cast<uint4*>(out + (size_t)(row0 + r) * ldo + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <typename TIn, typename TAct>
static w2vs_status_t ln_dispatch(const LayerNormArgs& a, cudaStream_t st) {
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div64(a.rows, wpb));
#define W2VS_LN_CASE(NCH)                                                                      \
  launch_pdl(layernorm_rows_kernel<TIn, TAct, NCH>, grid, dim3(wpb * 32), 0, st,                \
             (const TIn*)a.x, a.ldx, a.gamma, a.beta, a.out_f32, (TAct*)a.out_act, a.ldo, a.rows, a.N, a.gelu)
  if (a.N <= 256) W2VS_LN_CASE(1);
  else if (a.N <= 512) W2VS_LN_CASE(2);
  else if (a.N <= 1024) W2VS_LN_CASE(4);
  else W2VS_LN_CASE(8);
#undef W2VS_LN_CASE
  W2VS_CHECK_LAUNCH("layernorm_rows_kernel");
  return W2VS_OK;
}

// fp32-input fast path (the 2 LayerNorms of every transformer layer): lane owns float4 slices at
// 4*lane + 128*j, so every warp load is one fully coalesced 512-byte request (the generic kernel's
// 8-element chunks cost two half-used requests per fp32 chunk).
template <typename TAct, int NV>
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* out_f32, TAct* out_act, int64_t ldo, int rows, int gelu) {
  pdl_prologue();
  constexpr int N = NV * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ldx + 4 * lane;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = *reinterpret_cast<const float4*>(xr + 128 * j);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / N);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / N) + 1e-5f);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * lane + 128 * j);
    const float4 bt = *reinterpret_cast<const float4*>(beta + 4 * lane + 128 * j);
    float4 y;
    y.x = (v[j].x - mean) * rstd * g.x + bt.x;
    y.y = (v[j].y - mean) * rstd * g.y + bt.y;
    y.z = (v[j].z - mean) * rstd * g.z + bt.z;
    y.w = (v[j].w - mean) * rstd * g.w + bt.w;
    if (gelu) {
      if (out_f32) { gelu_erf2(y.x, y.y); gelu_erf2(y.z, y.w); } else { gelu2<TAct>(y.x, y.y); gelu2<TAct>(y.z, y.w); }
    }
    const size_t o = (size_t)row * ldo + 4 * lane + 128 * j;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = y;
    if (out_act) {
      if (sizeof(TAct) == 4) {
        *reinterpret_cast<float4*>(out_act + o) = y;
      } else {
        uint2 u;
        u.x = pack_bf16x2(y.x, y.y);
        u.y = pack_bf16x2(y.z, y.w);
        *reinterpret_cast<uint2*>(out_act + o) = u;
      }
    }
  }
}

template <typename TAct>
static w2vs_status_t ln_f32_dispatch(const LayerNormArgs& a, cudaStream_t st) {
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div64(a.rows, wpb));
#define W2VS_LNF_CASE(NV)                                                                          \
  launch_pdl(layernorm_f32_kernel<TAct, NV>, grid, dim3(wpb * 32), 0, st, (const float*)a.x, a.ldx, a.gamma, a.beta, \
             a.out_f32, (TAct*)a.out_act, a.ldo, a.rows, a.gelu)
  switch (a.N / 128) {
    case 1: W2VS_LNF_CASE(1); break;
    case 2: W2VS_LNF_CASE(2); break;
    case 4: W2VS_LNF_CASE(4); break;
    case 6: W2VS_LNF_CASE(6); break;
    case 8: W2VS_LNF_CASE(8); break;
    default: return W2VS_UNSUPPORTED;
  }
#undef W2VS_LNF_CASE
  W2VS_CHECK_LAUNCH("layernorm_rows_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_layernorm(const LayerNormArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(a.N % 8 == 0 && a.N >= 8 && a.N <= 2048, "LayerNorm width must be a multiple of 8, <= 2048");
  W2VS_REQUIRE(a.ldx % 8 == 0 && a.ldo % 8 == 0, "LayerNorm leading dims must be multiples of 8");
  if (a.rows <= 0) return W2VS_OK;
  if (a.in_dtype == W2VS_F32) {
    const int nv = a.N / 128;
    if (a.N % 128 == 0 && (nv == 1 || nv == 2 || nv == 4 || nv == 6 || nv == 8))
      return a.act_dtype == W2VS_F32 ? ln_f32_dispatch<float>(a, st) : ln_f32_dispatch<bf16>(a, st);
    return a.act_dtype == W2VS_F32 ? ln_dispatch<float, float>(a, st) : ln_dispatch<float, bf16>(a, st);
  }
  if (a.act_dtype == W2VS_BF16 && a.gelu && a.out_f32 == nullptr && a.out_act != nullptr && a.N % 256 == 0 && a.N <= 1024 &&
      a.N != 768) {
    const dim3 grid((unsigned)ceil_div64(a.rows, 8 * LG_ROWS));
    switch (a.N / 256) {
      case 1: launch_pdl(ln_gelu_bf16_kernel<1>, grid, dim3(256), 0, st, (const bf16*)a.x, a.ldx, a.gamma, a.beta, (bf16*)a.out_act, a.ldo, a.rows); break;
      case 2: launch_pdl(ln_gelu_bf16_kernel<2>, grid, dim3(256), 0, st, (const bf16*)a.x, a.ldx, a.gamma, a.beta, (bf16*)a.out_act, a.ldo, a.rows); break;
      default: launch_pdl(ln_gelu_bf16_kernel<4>, grid, dim3(256), 0, st, (const bf16*)a.x, a.ldx, a.gamma, a.beta, (bf16*)a.out_act, a.ldo, a.rows); break;
    }
    W2VS_CHECK_LAUNCH("layernorm_rows_kernel");
    return W2VS_OK;
  }
  return a.act_dtype == W2VS_F32 ? ln_dispatch<bf16, float>(a, st) : ln_dispatch<bf16, bf16>(a, st);
}

// ------------------------------------------------------------------------------------------------
// prep_masks: one CTA per utterance.
//   frame_pad[b,t] = all(mask[b, t*w : (t+1)*w]),  w = mask_len / T       (wav2vec2.py:560-565)
//   pos[b,t]       = frame_pad ? 1 : 1 + #{t' <= t : !frame_pad[b,t']}    (utils.py:250-260, padding_idx=1)
//   keypad[b,m]    = extended key padding over M = T2 + R tokens           (wav2vec_S.py:470-476)
__global__ void __launch_bounds__(256)
prep_masks_kernel(const int32_t* __restrict__ lengths, const uint8_t* __restrict__ sample_mask,
                  int mask_len, uint8_t* __restrict__ frame_pad, int32_t* __restrict__ pos,
                  uint8_t* __restrict__ keypad, uint8_t* __restrict__ pad_blk, int T, int T2, int M, int main_ctx,
                  int rc) {
  __shared__ int s_part[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int w = mask_len > 0 ? mask_len / T : 0;
  const int seg = (T + 255) / 256;
  const int t_lo = min(tid * seg, T), t_hi = min(t_lo + seg, T);
  int cnt = 0;
  for (int t = t_lo; t < t_hi; ++t) {
    bool pad = false;
    if (lengths) {
      pad = (int64_t)t * w >= (int64_t)lengths[b];
    } else if (sample_mask) {
      pad = true;
      const uint8_t* mrow = sample_mask + (size_t)b * mask_len + (size_t)t * w;
      for (int i = 0; i < w; ++i)
        if (!mrow[i]) { pad = false; break; }
    }
    frame_pad[(size_t)b * T + t] = pad ? 1 : 0;
    cnt += pad ? 0 : 1;
  }
  s_part[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { const int c = s_part[i]; s_part[i] = run; run += c; }
  }
  __syncthreads();
  int run = s_part[tid];
  for (int t = t_lo; t < t_hi; ++t) {
    const bool pad = frame_pad[(size_t)b * T + t] != 0;
    if (!pad) ++run;
    pos[(size_t)b * T + t] = pad ? 1 : 1 + run;
  }
  __syncthreads();
  for (int m = tid; m < M; m += 256) {
    bool kp;
    if (m < T2) {
      kp = m >= T ? true : frame_pad[(size_t)b * T + m] != 0;
    } else {
      const int r = m - T2;
      int src = (r / rc + 1) * main_ctx + (r % rc);
      const bool oor = src > T2 - 1;
      src = min(src, T2 - 1);
      kp = oor || (src >= T ? true : frame_pad[(size_t)b * T + src] != 0);
    }
    keypad[(size_t)b * M + m] = kp ? 1 : 0;
  }
  // pad_blk[b][j] = any padded key among tokens [128 j, 128 j + 128): lets the attention kernel skip the
  // per-key padding bytes of key tiles that have none
  if (pad_blk != nullptr) {
    __syncthreads();
    const int nblk = (M + 127) / 128;
    for (int j = tid; j < nblk; j += 256) {
      uint8_t any = 0;
      for (int m = 128 * j; m < min(M, 128 * j + 128); ++m) any |= keypad[(size_t)b * M + m];
      pad_blk[(size_t)b * nblk + j] = any;
    }
  }
}

w2vs_status_t launch_prep_masks(const PrepArgs& a, cudaStream_t st) {
  prep_masks_kernel<<<a.B, 256, 0, st>>>(a.lengths, a.sample_mask, a.mask_len, a.frame_pad, a.pos,
                                         a.keypad, a.pad_blk, a.T, a.T2, a.M, a.main_ctx, a.rc);
  W2VS_CHECK_LAUNCH("prep_masks_kernel");
  return W2VS_OK;
}

// ------------------------------------------------------------------------------------------------
// embed_tokens: one warp per token (b, m).
template <typename TAct, int NCH>
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const float* __restrict__ feats, int feat_rows, const uint8_t* __restrict__ frame_pad,
                    const int32_t* __restrict__ pos, int pos_offset, const float* __restrict__ sin_table,
                    const float* __restrict__ posconv, int posconv_rows, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ X, TAct* __restrict__ Xa,
                    int B, int T, int T2, int M, int main_ctx, int rc, int D) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= (int64_t)B * M) return;
  const int b = (int)(tok / M), m = (int)(tok % M);
  int t = m;
  if (m >= T2) {
    const int r = m - T2;
    t = min((r / rc + 1) * main_ctx + (r % rc), T2 - 1);
  }
  float v[NCH][8];
  const bool zero_row = t >= T;  // sequence padding added after the LayerNorm: exact zeros
  const bool pad = !zero_row && frame_pad != nullptr && frame_pad[(size_t)b * T + t] != 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
      if (!zero_row) {
        if (!pad) load8(feats + ((size_t)b * feat_rows + t) * D + c0, v[j]);
        float p[8];
        if (posconv) {
          load8(posconv + ((size_t)b * posconv_rows + t) * D + c0, p);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[j][e] += p[e];
        } else if (!pad) {
          const int pi = pos ? pos[(size_t)b * T + t] : t + pos_offset;
          load8(sin_table + (size_t)pi * D + c0, p);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[j][e] += p[e];
        }
      }
    }
  }
  if (gamma != nullptr && !zero_row) warp_layernorm<NCH>(v, D, lane, gamma, beta);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < D) {
      store8(X + (size_t)tok * D + c0, v[j]);
      if (Xa) store8(Xa + (size_t)tok * D + c0, v[j]);
    }
  }
}

template <typename TAct>
static w2vs_status_t embed_dispatch(const EmbedArgs& a, cudaStream_t st) {
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div64((int64_t)a.B * a.M, wpb));
#define W2VS_EMB_CASE(NCH)                                                                        \
  launch_pdl(embed_tokens_kernel<TAct, NCH>, grid, dim3(wpb * 32), 0, st,                         \
             a.feats, a.feat_rows, a.frame_pad, a.pos, a.pos_offset, a.sin_table, a.posconv, a.posconv_rows > 0 ? a.posconv_rows : a.T, a.gamma, \
             a.beta, a.X, (TAct*)a.Xa, a.B, a.T, a.T2, a.M, a.main_ctx, a.rc, a.D)
  if (a.D <= 256) W2VS_EMB_CASE(1);
  else if (a.D <= 512) W2VS_EMB_CASE(2);
  else if (a.D <= 1024) W2VS_EMB_CASE(4);
  else W2VS_EMB_CASE(8);
#undef W2VS_EMB_CASE
  W2VS_CHECK_LAUNCH("embed_tokens_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_embed(const EmbedArgs& a, cudaStream_t st) {
  return a.act_dtype == W2VS_F32 ? embed_dispatch<float>(a, st) : embed_dispatch<bf16>(a, st);
}

// ------------------------------------------------------------------------------------------------
// finalize_rows: one warp per output frame (b, t), t < T_out.
template <typename TOut, int NCH>
__global__ void __launch_bounds__(256)
finalize_rows_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                     const float* __restrict__ beta, TOut* __restrict__ out, int B, int T_out,
                     int64_t in_rows_per_utt, int D, int tbd) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= (int64_t)B * T_out) return;
  const int b = (int)(r / T_out), t = (int)(r % T_out);
  float v[NCH][8];
  const float* xr = X + ((size_t)b * in_rows_per_utt + t) * D;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < D) load8(xr + c0, v[j]);
  }
  if (gamma != nullptr) warp_layernorm<NCH>(v, D, lane, gamma, beta);
  TOut* o = out + (tbd ? ((size_t)t * B + b) : ((size_t)b * T_out + t)) * D;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    if (c0 < D) store8(o + c0, v[j]);
  }
}

template <typename TOut>
static w2vs_status_t finalize_dispatch(const FinalizeArgs& a, cudaStream_t st) {
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div64((int64_t)a.B * a.T_out, wpb));
#define W2VS_FIN_CASE(NCH)                                                            \
  launch_pdl(finalize_rows_kernel<TOut, NCH>, grid, dim3(wpb * 32), 0, st,             \
             a.X, a.gamma, a.beta, (TOut*)a.out, a.B, a.T_out, a.in_rows_per_utt, a.D, a.tbd)
  if (a.D <= 256) W2VS_FIN_CASE(1);
  else if (a.D <= 512) W2VS_FIN_CASE(2);
  else if (a.D <= 1024) W2VS_FIN_CASE(4);
  else W2VS_FIN_CASE(8);
#undef W2VS_FIN_CASE
  W2VS_CHECK_LAUNCH("finalize_rows_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_finalize(const FinalizeArgs& a, cudaStream_t st) {
  if (a.B * a.T_out <= 0) return W2VS_OK;
  if (a.out_dtype == W2VS_F16) return finalize_dispatch<__half>(a, st);     // fp16 models: w2vs_config.io_dtype
  return a.out_dtype == W2VS_F32 ? finalize_dispatch<float>(a, st) : finalize_dispatch<bf16>(a, st);
}

// ------------------------------------------------------------------------------------------------
// small utility kernels
template <typename TIn>
__global__ void copy_rows_to_f32_kernel(const TIn* __restrict__ src, int64_t src_rows_per_utt,
                                        float* __restrict__ dst, int B, int T, int C) {
  const int64_t n = (int64_t)B * T * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bt = i / C;
    const int t = (int)(bt % T), b = (int)(bt / T);
    dst[i] = to_f32(src[((size_t)b * src_rows_per_utt + t) * C + c]);
  }
}

w2vs_status_t launch_tap_rows(const void* src, int src_dtype, int64_t src_rows_per_utt, float* dst,
                              int B, int T, int C, cudaStream_t st) {
  const int64_t n = (int64_t)B * T * C;
  if (n <= 0) return W2VS_OK;
  const int grid = (int)(ceil_div64(n, 256) < 4096 ? ceil_div64(n, 256) : 4096);
  if (src_dtype == W2VS_F32)
    copy_rows_to_f32_kernel<float><<<grid, 256, 0, st>>>((const float*)src, src_rows_per_utt, dst, B, T, C);
  else
    copy_rows_to_f32_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)src, src_rows_per_utt, dst, B, T, C);
  W2VS_CHECK_LAUNCH("copy_rows_to_f32_kernel");
  return W2VS_OK;
}

__global__ void copy_u8_rows_kernel(const uint8_t* __restrict__ src, int src_ld, uint8_t* __restrict__ dst,
                                    int dst_ld, int B, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * n; i += gridDim.x * blockDim.x) {
    const int b = i / n, t = i % n;
    dst[(size_t)b * dst_ld + t] = src[(size_t)b * src_ld + t];
  }
}

w2vs_status_t launch_copy_mask(const uint8_t* src, int src_ld, uint8_t* dst, int dst_ld, int B, int n,
                               cudaStream_t st) {
  if (B * n <= 0) return W2VS_OK;
  copy_u8_rows_kernel<<<(B * n + 255) / 256, 256, 0, st>>>(src, src_ld, dst, dst_ld, B, n);
  W2VS_CHECK_LAUNCH("copy_u8_rows_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
