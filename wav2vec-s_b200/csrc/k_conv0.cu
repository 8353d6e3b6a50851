// Layer-0 feature-extractor convolution: Conv1d(1 -> C, k, stride s), no padding, with its
// normalisation + GELU fused (ConvFeatureExtractionModel block 0, wav2vec2.py:715-752,773-781).
//
// HBM-streaming kernel: reads the waveform once, writes the channels-last activation [B, rows, C]
// once.  One warp produces whole frames: lane l owns channel pairs {2l + 64 i, 2l + 64 i + 1}, so a
// warp-wide store of one `i` is one contiguous 128 B (bf16) / 256 B (fp32) segment, and the
// per-frame LayerNorm over C is a warp-shuffle reduction.  The C*k filter taps of a lane's channels
// live in registers for the whole CTA lifetime (160 registers for C=512,k=10), so the inner loop is
// pure FFMA on broadcast waveform samples.
//
// Modes:  PLAIN     y = gelu(conv + bias)
//         LN        y = gelu(LayerNorm_C(conv + bias))            (extractor_mode = layer_norm)
//         GN_STATS  per-CTA partial sum / sum of squares per (utterance, channel); a small finalize
//                   kernel reduces them in a fixed order (fp64) -> bit-reproducible scale / shift
//         GN_APPLY  y = gelu(GroupNorm(conv + bias)), groups == channels   (extractor_mode = default)
#include "common.cuh"
#include "kernels.h"

namespace w2vs {

enum { C0_PLAIN = 0, C0_LN = 1, C0_GN_STATS = 2, C0_GN_APPLY = 3 };

template <typename TIn, typename TOut, int NI, int KW, int MODE>
__global__ void __launch_bounds__(256, 1)
conv0_kernel(const TIn* __restrict__ wav, int64_t wav_ld, const float* __restrict__ w,
             const float* __restrict__ bias, const float* __restrict__ gamma,
             const float* __restrict__ beta, TOut* __restrict__ out, int rows_per_utt, int T0,
             int stride, float* __restrict__ gn_stats, int frames_per_cta) {
  constexpr int C = NI * 64;
  __shared__ float s_bias[C];
  __shared__ float s_scale[C];   // LN: gamma; GN_APPLY: gamma * rstd
  __shared__ float s_shift[C];   // LN: beta;  GN_APPLY: beta - mean * gamma * rstd
  __shared__ float s_red[MODE == C0_GN_STATS ? 8 * 2 * C : 1];  // GN_STATS: per-warp partials

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_bias[c] = bias ? bias[c] : 0.f;
    if (MODE == C0_LN) {
      s_scale[c] = gamma[c];
      s_shift[c] = beta[c];
    } else if (MODE == C0_GN_APPLY) {
      s_scale[c] = gn_stats[((size_t)b * C + c) * 2 + 0];
      s_shift[c] = gn_stats[((size_t)b * C + c) * 2 + 1];
    }
  }
  // filter taps of this lane's channels -> registers
  float wr[NI][2][KW];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < KW; ++j) wr[i][h][j] = w[(size_t)(2 * lane + 64 * i + h) * KW + j];
  __syncthreads();

  float st_sum[NI][2], st_sq[NI][2];
  if (MODE == C0_GN_STATS) {
#pragma unroll
    for (int i = 0; i < NI; ++i) { st_sum[i][0] = st_sum[i][1] = 0.f; st_sq[i][0] = st_sq[i][1] = 0.f; }
  }

  const int t_begin = blockIdx.x * frames_per_cta;
  const int t_end = min(t_begin + frames_per_cta, T0);
  const TIn* x = wav + (size_t)b * wav_ld;
  const int nwarps = blockDim.x >> 5;

  for (int t0 = t_begin + 2 * warp; t0 < t_end; t0 += 2 * nwarps) {
    const bool has2 = (t0 + 1) < t_end;
    float acc[2][NI][2];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        acc[f][i][0] = s_bias[2 * lane + 64 * i];
        acc[f][i][1] = s_bias[2 * lane + 64 * i + 1];
      }
    const int64_t base = (int64_t)t0 * stride;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
      const float x0 = to_f32(x[base + j]);
      const float x1 = has2 ? to_f32(x[base + stride + j]) : 0.f;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        acc[0][i][0] = fmaf(wr[i][0][j], x0, acc[0][i][0]);
        acc[0][i][1] = fmaf(wr[i][1][j], x0, acc[0][i][1]);
        acc[1][i][0] = fmaf(wr[i][0][j], x1, acc[1][i][0]);
        acc[1][i][1] = fmaf(wr[i][1][j], x1, acc[1][i][1]);
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      if (f == 1 && !has2) break;
      if (MODE == C0_GN_STATS) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            st_sum[i][h] += acc[f][i][h];
            st_sq[i][h] = fmaf(acc[f][i][h], acc[f][i][h], st_sq[i][h]);
          }
        continue;
      }
      float mean = 0.f, rstd = 1.f;
      if (MODE == C0_LN) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) s += acc[f][i][0] + acc[f][i][1];
        mean = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const float d0 = acc[f][i][0] - mean, d1 = acc[f][i][1] - mean;
          q = fmaf(d0, d0, q);
          q = fmaf(d1, d1, q);
        }
        rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
      }
      TOut* o = out + ((size_t)b * rows_per_utt + (t0 + f)) * C;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = 2 * lane + 64 * i;
        float v0 = acc[f][i][0], v1 = acc[f][i][1];
        if (MODE == C0_LN) {
          v0 = (v0 - mean) * rstd * s_scale[c] + s_shift[c];
          v1 = (v1 - mean) * rstd * s_scale[c + 1] + s_shift[c + 1];
        } else if (MODE == C0_GN_APPLY) {
          v0 = fmaf(v0, s_scale[c], s_shift[c]);
          v1 = fmaf(v1, s_scale[c + 1], s_shift[c + 1]);
        }
        v0 = gelu_erf(v0);
        v1 = gelu_erf(v1);
        if (sizeof(TOut) == 2) {
          *reinterpret_cast<uint32_t*>(o + c) = pack_bf16x2(v0, v1);
        } else {
          *reinterpret_cast<float2*>(o + c) = make_float2(v0, v1);
        }
      }
    }
  }

  if (MODE == C0_GN_STATS) {
    // per-warp partials -> fixed-order sum over the 8 warps -> one partial row per CTA
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        s_red[(warp * 2 + 0) * C + 2 * lane + 64 * i + h] = st_sum[i][h];
        s_red[(warp * 2 + 1) * C + 2 * lane + 64 * i + h] = st_sq[i][h];
      }
    __syncthreads();
    float* part = gn_stats + (size_t)gridDim.y * C * 2 + ((size_t)b * gridDim.x + blockIdx.x) * 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
      float acc = 0.f;
      for (int w = 0; w < 8; ++w) acc += s_red[(w * 2 + c / C) * C + (c % C)];
      part[c] = acc;
    }
  }
}

// gn_stats layout (floats): [B][C][2] scale/shift, then [B][n_cta][2][C] partials.
__global__ void __launch_bounds__(256)
conv0_gn_finalize_kernel(float* __restrict__ gn_stats, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int B, int C, int n_cta, int T0) {
  const int b = blockIdx.x;
  const float* part = gn_stats + (size_t)B * C * 2 + (size_t)b * n_cta * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double sum = 0.0, sq = 0.0;
    for (int i = 0; i < n_cta; ++i) {
      sum += (double)part[(size_t)i * 2 * C + c];
      sq += (double)part[(size_t)i * 2 * C + C + c];
    }
    const double mean = sum / T0;
    double var = sq / T0 - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float sc = gamma[c] * rstd;
    gn_stats[((size_t)b * C + c) * 2 + 0] = sc;
    gn_stats[((size_t)b * C + c) * 2 + 1] = beta[c] - (float)mean * sc;
  }
}

constexpr int kFramesPerCta = 256;

template <typename TIn, typename TOut, int NI, int MODE>
static w2vs_status_t launch_one(const Conv0Args& a, cudaStream_t st) {
  const int frames_per_cta = kFramesPerCta;
  dim3 grid((unsigned)ceil_div64(a.T0, frames_per_cta), (unsigned)a.B);
  conv0_kernel<TIn, TOut, NI, 10, MODE><<<grid, 256, 0, st>>>(
      (const TIn*)a.wav, a.wav_ld, a.w, a.bias, a.gamma, a.beta, (TOut*)a.out, a.rows_per_utt, a.T0,
      a.stride, a.gn_stats, frames_per_cta);
  W2VS_CHECK_LAUNCH("conv0_kernel");
  return W2VS_OK;
}

template <typename TIn, typename TOut, int NI>
static w2vs_status_t launch_mode(const Conv0Args& a, cudaStream_t st) {
  switch (a.norm) {
    case CONV0_NORM_NONE: return launch_one<TIn, TOut, NI, C0_PLAIN>(a, st);
    case CONV0_NORM_LAYER: return launch_one<TIn, TOut, NI, C0_LN>(a, st);
    case CONV0_NORM_GROUP: {
      W2VS_TRY((launch_one<TIn, TOut, NI, C0_GN_STATS>(a, st)));
      conv0_gn_finalize_kernel<<<a.B, 256, 0, st>>>(a.gn_stats, a.gamma, a.beta, a.B, a.C,
                                                    (int)ceil_div64(a.T0, kFramesPerCta), a.T0);
      W2VS_CHECK_LAUNCH("conv0_gn_finalize_kernel");
      return launch_one<TIn, TOut, NI, C0_GN_APPLY>(a, st);
    }
  }
  return W2VS_INVALID_VALUE;
}

template <typename TIn, typename TOut>
static w2vs_status_t launch_ni(const Conv0Args& a, cudaStream_t st) {
  switch (a.C / 64) {
    case 1: return launch_mode<TIn, TOut, 1>(a, st);
    case 2: return launch_mode<TIn, TOut, 2>(a, st);
    case 4: return launch_mode<TIn, TOut, 4>(a, st);
    case 8: return launch_mode<TIn, TOut, 8>(a, st);
  }
  set_error("unsupported: conv0 channels %d (supported 64/128/256/512)", a.C);
  return W2VS_UNSUPPORTED;
}

w2vs_status_t launch_conv0(const Conv0Args& a, cudaStream_t st) {
  if (a.k != 10 || a.C % 64 != 0) {
    set_error("unsupported: first conv layer (%d,%d,%d); this build implements k=10", a.C, a.k, a.stride);
    return W2VS_UNSUPPORTED;
  }
  if (a.wav_dtype == W2VS_F32) {
    return a.out_dtype == W2VS_F32 ? launch_ni<float, float>(a, st) : launch_ni<float, bf16>(a, st);
  }
  return a.out_dtype == W2VS_F32 ? launch_ni<bf16, float>(a, st) : launch_ni<bf16, bf16>(a, st);
}

}  // namespace w2vs
