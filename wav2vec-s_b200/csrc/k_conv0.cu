// Layer-0 feature-extractor convolution: Conv1d(1 -> C, k, stride s), no padding, with its
// normalisation + GELU fused (ConvFeatureExtractionModel block 0, wav2vec2.py:715-752,773-781).
//
// Streaming kernel: reads the waveform once, writes the channels-last activation [B, rows, C] once.
// One warp produces whole frames: lane l owns channel pairs {2l + 64 i, 2l + 64 i + 1}, so a warp-wide
// store of one `i` is one contiguous 128 B (bf16) / 256 B (fp32) segment and the per-frame LayerNorm
// over C is a warp-shuffle reduction.  The filter taps (tap-major) and the CTA's waveform slice sit
// in shared memory; a warp iteration computes 4 frames so that each 8-byte weight read feeds 8 FFMAs.
// Registers stay under 128 (two CTAs = 16 warps per SM): the kernel is bound by the ~30 ALU
// instructions per output element of LayerNorm + erf-GELU, not by HBM (65.5 MB per 20 s utterance).
//
// Modes:  PLAIN     y = gelu(conv + bias)
//         LN        y = gelu(LayerNorm_C(conv + bias))            (extractor_mode = layer_norm)
//         GN_STATS  per-CTA partial sum / sum of squares per (utterance, channel); a small finalize
//                   kernel reduces them in a fixed order (fp64) -> bit-reproducible scale / shift
//         GN_APPLY  y = gelu(GroupNorm(conv + bias)), groups == channels   (extractor_mode = default)
#include <limits.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {

enum { C0_PLAIN = 0, C0_LN = 1, C0_GN_STATS = 2, C0_GN_APPLY = 3 };

constexpr int kFramesPerCta = 256;
#ifndef W2VS_CONV0_FPI
#define W2VS_CONV0_FPI 4
#endif
#ifndef W2VS_CONV0_CTAS
#define W2VS_CONV0_CTAS 2
#endif
constexpr int FPI = W2VS_CONV0_FPI;   // frames per warp iteration

template <typename TIn, typename TOut, int NI, int KW, int MODE>
__global__ void __launch_bounds__(256, W2VS_CONV0_CTAS)
conv0_kernel(const TIn* __restrict__ wav, int64_t wav_ld, const float* __restrict__ w,
             const float* __restrict__ bias, const float* __restrict__ gamma,
             const float* __restrict__ beta, TOut* __restrict__ out, int rows_per_utt, int T0,
             int stride, float* __restrict__ gn_stats, int frames_per_cta, const float* __restrict__ wav_stats,
             const int32_t* __restrict__ wav_lengths) {
  constexpr int C = NI * 64;
  pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                       // [KW][C]  tap-major: a lane's channel pairs are 8-byte contiguous
  float* s_bias = s_w + KW * C;            // [C]
  float* s_scale = s_bias + C;             // LN: gamma; GN_APPLY: gamma * rstd
  float* s_shift = s_scale + C;            // LN: beta;  GN_APPLY: beta - mean * gamma * rstd
  float* s_x = s_shift + C;                // waveform samples of this CTA: (frames_per_cta - 1) * stride + KW
  float* s_red = s_x + ((kFramesPerCta - 1) * 8 + KW + 3) / 4 * 4;   // GN_STATS: [8 warps][2][C]

  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t_begin = blockIdx.x * frames_per_cta;
  const int t_end = min(t_begin + frames_per_cta, T0);
  const int n_frames = t_end - t_begin;
  const TIn* x = wav + (size_t)b * wav_ld + (size_t)t_begin * stride;
  const int n_samples = (n_frames - 1) * stride + KW;
  if (wav_stats != nullptr) {
    // waveform front end: standardise the utterance's valid samples on load (kernels.h: launch_wav_stats)
    const float mean = wav_stats[2 * b], rstd = wav_stats[2 * b + 1];
    const int n_valid = (wav_lengths != nullptr ? wav_lengths[b] : INT_MAX) - t_begin * stride;
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x) {
      const float v = to_f32(x[i]);
      s_x[i] = i < n_valid ? (v - mean) * rstd : v;
    }
  } else {
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x) s_x[i] = to_f32(x[i]);
  }
  for (int i = threadIdx.x; i < KW * C; i += blockDim.x) {
    const int j = i / C, c = i % C;
    s_w[i] = w[(size_t)c * KW + j];
  }
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
  __syncthreads();

  float st_sum[NI][2], st_sq[NI][2];
  if (MODE == C0_GN_STATS) {
#pragma unroll
    for (int i = 0; i < NI; ++i) { st_sum[i][0] = st_sum[i][1] = 0.f; st_sq[i][0] = st_sq[i][1] = 0.f; }
  }
  const int nwarps = blockDim.x >> 5;

  for (int f0 = FPI * warp; f0 < n_frames; f0 += FPI * nwarps) {
    // accumulators as packed f32x2 pairs (channels 2*lane, 2*lane+1 of each 64-channel group): one FFMA2 per tap
    // and pair instead of two FFMAs, and the LayerNorm / GELU arithmetic below stays packed as well
    uint64_t acc[FPI][NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const uint64_t bb = *reinterpret_cast<const uint64_t*>(s_bias + 64 * i + 2 * lane);
#pragma unroll
      for (int f = 0; f < FPI; ++f) acc[f][i] = bb;
    }
#pragma unroll
    for (int j = 0; j < KW; ++j) {
      uint64_t xv[FPI];
#pragma unroll
      for (int f = 0; f < FPI; ++f) {
        const float xs = (f0 + f) < n_frames ? s_x[(f0 + f) * stride + j] : 0.f;   // broadcast
        xv[f] = pack2(xs, xs);
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const uint64_t ww = *reinterpret_cast<const uint64_t*>(s_w + j * C + 64 * i + 2 * lane);
#pragma unroll
        for (int f = 0; f < FPI; ++f) acc[f][i] = ffma2(ww, xv[f], acc[f][i]);
      }
    }
#pragma unroll
    for (int f = 0; f < FPI; ++f) {
      if (f0 + f >= n_frames) break;
      if (MODE == C0_GN_STATS) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          float a0, a1;
          unpack2(acc[f][i], a0, a1);
          st_sum[i][0] += a0; st_sum[i][1] += a1;
          st_sq[i][0] = fmaf(a0, a0, st_sq[i][0]);
          st_sq[i][1] = fmaf(a1, a1, st_sq[i][1]);
        }
        continue;
      }
      uint64_t rs2 = pack2(1.f, 1.f);
      if (MODE == C0_LN) {
        uint64_t s2 = pack2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NI; ++i) s2 = fadd2(s2, acc[f][i]);
        float sa, sb;
        unpack2(s2, sa, sb);
        const float mean = warp_sum(sa + sb) * (1.0f / C);
        const uint64_t nm2 = pack2(-mean, -mean);
        uint64_t q2 = pack2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NI; ++i) { acc[f][i] = fadd2(acc[f][i], nm2); q2 = ffma2(acc[f][i], acc[f][i], q2); }
        unpack2(q2, sa, sb);
        const float rstd = 1.0f / sqrtf(warp_sum(sa + sb) * (1.0f / C) + 1e-5f);
        rs2 = pack2(rstd, rstd);
      }
      TOut* o = out + ((size_t)b * rows_per_utt + (t_begin + f0 + f)) * C;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = 2 * lane + 64 * i;
        uint64_t v = acc[f][i];
        if (MODE == C0_LN || MODE == C0_GN_APPLY) {
          const uint64_t sc = *reinterpret_cast<const uint64_t*>(s_scale + c);
          const uint64_t sh = *reinterpret_cast<const uint64_t*>(s_shift + c);
          v = ffma2(v, MODE == C0_LN ? fmul2(rs2, sc) : sc, sh);     // LN: v is already centred
        }
        float v0, v1;
        if (sizeof(TOut) == 2) {
          unpack2(gelu_tanh2p(v), v0, v1);
          *reinterpret_cast<uint32_t*>(o + c) = pack_bf16x2(v0, v1);
        } else {
          unpack2(v, v0, v1);
          gelu_erf2(v0, v1);
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
      for (int w8 = 0; w8 < 8; ++w8) acc += s_red[(w8 * 2 + c / C) * C + (c % C)];
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

template <typename TIn, typename TOut, int NI, int MODE>
static w2vs_status_t launch_one(const Conv0Args& a, cudaStream_t st) {
  constexpr int C = NI * 64, KW = 10;
  // 256 frames per CTA amortise the tap / parameter loads over a long utterance; a decision step of the incremental
  // mode has ~1000 frames in all, which would be four CTAs: spread those over the SMs (multiples of 16 frames = one
  // pass of the 8 warps).  The GroupNorm passes keep 256 (their partial-statistics layout is sized by it).
  int frames_per_cta = kFramesPerCta;
  if (MODE == C0_PLAIN || MODE == C0_LN) {
    const int64_t per_cta = ceil_div64((int64_t)a.T0 * a.B, 2 * num_sms());
    if (per_cta < kFramesPerCta) frames_per_cta = (int)((per_cta + 15) / 16 * 16);
  }
  W2VS_REQUIRE(a.stride >= 1 && a.stride <= 8, "first conv stride must be <= 8");
  size_t smem = (size_t)(KW * C + 3 * C + ((kFramesPerCta - 1) * 8 + KW + 3) / 4 * 4) * sizeof(float);
  if (MODE == C0_GN_STATS) smem += (size_t)8 * 2 * C * sizeof(float);
  static PerDeviceOnce attr_once;   // the attribute belongs to the current device's copy of the kernel
  bool& attr_done = attr_once.here();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv0_kernel<TIn, TOut, NI, KW, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv0 smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    attr_done = true;
  }
  dim3 grid((unsigned)ceil_div64(a.T0, frames_per_cta), (unsigned)a.B);
  launch_pdl(conv0_kernel<TIn, TOut, NI, KW, MODE>, grid, dim3(256), smem, st,
             (const TIn*)a.wav, a.wav_ld, a.w, a.bias, a.gamma, a.beta, (TOut*)a.out, a.rows_per_utt, a.T0,
             a.stride, a.gn_stats, frames_per_cta, a.wav_stats, a.wav_lengths);
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
  if (a.wav_dtype == W2VS_I16) {
    return a.out_dtype == W2VS_F32 ? launch_ni<int16_t, float>(a, st) : launch_ni<int16_t, bf16>(a, st);
  }
  return a.out_dtype == W2VS_F32 ? launch_ni<bf16, float>(a, st) : launch_ni<bf16, bf16>(a, st);
}

// ---- waveform front end: per-utterance statistics ------------------------------------------------------------------
// One CTA per utterance, two passes (mean, then centred squares) in fp32 with a fixed reduction order.
template <typename TIn>
__global__ void __launch_bounds__(1024)
wav_stats_kernel(const TIn* __restrict__ wav, int64_t wav_ld, const int32_t* __restrict__ lengths, int L,
                 float* __restrict__ stats) {
  __shared__ float s_red[32];
  __shared__ float s_bcast;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = lengths != nullptr ? min(max(lengths[b], 1), L) : L;
  const TIn* x = wav + (size_t)b * wav_ld;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      float t = lane < (int)(blockDim.x >> 5) ? s_red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) s_bcast = t;
    }
    __syncthreads();
    return s_bcast;
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += to_f32(x[i]);
  const float mean = block_sum(s) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = to_f32(x[i]) - mean; q = fmaf(d, d, q); }
  const float var = block_sum(q) / (float)n;
  if (threadIdx.x == 0) { stats[2 * b] = mean; stats[2 * b + 1] = 1.0f / sqrtf(var + 1e-5f); }
}

w2vs_status_t launch_wav_stats(const void* wav, int wav_dtype, int64_t wav_ld, const int32_t* lengths, int L, int B,
                               float* stats, cudaStream_t st) {
  if (wav_dtype == W2VS_F32) wav_stats_kernel<float><<<B, 1024, 0, st>>>((const float*)wav, wav_ld, lengths, L, stats);
  else if (wav_dtype == W2VS_I16) wav_stats_kernel<int16_t><<<B, 1024, 0, st>>>((const int16_t*)wav, wav_ld, lengths, L, stats);
  else wav_stats_kernel<bf16><<<B, 1024, 0, st>>>((const bf16*)wav, wav_ld, lengths, L, stats);
  W2VS_CHECK_LAUNCH("wav_stats_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
