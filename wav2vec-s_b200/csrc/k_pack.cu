// Weight repacking (run once per load_state_dict) and the optional convolutional positional
// embedding (pos_type="conv": wav2vec2.py:791-804; not used by the released wav2vec-S models): a direct fp32
// kernel for the fp32 parity mode, and the operand packing of the tensor-core path (bf16 models), where every
// group is an implicit GEMM on the tcgen05 kernel (api.cu).
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
namespace {

template <typename TOut>
__global__ void pack_copy_kernel(const float* __restrict__ src, TOut* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f32<TOut>(src[i]);
}

// Linear weight [N][K] fp32 -> bf16 slabs [N/8][K/KC][8][KC] (every 8-row x KC-wide slab contiguous: one bulk copy
// per slab in the fused incremental step)
__global__ void pack_slabs_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int N, int K, int KC) {
  const int64_t n = (int64_t)N * K;
  const int nch = K / KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), row = (int)(i / K);
    const int u = row >> 3, r = row & 7, j = k / KC, kk = k - j * KC;
    dst[(((size_t)u * nch + j) * 8 + r) * KC + kk] = __float2bfloat16_rn(src[i]);
  }
}

// Conv1d weight [C_out, C_in, k] -> K-major GEMM operand [C_out][j*C_in + ci]
template <typename TOut>
__global__ void pack_conv_kernel(const float* __restrict__ src, TOut* __restrict__ dst, int C_out, int C_in, int k) {
  const int64_t n = (int64_t)C_out * C_in * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C_in);
    const int j = (int)((i / C_in) % k);
    const int co = (int)(i / ((int64_t)C_in * k));
    dst[i] = from_f32<TOut>(src[((size_t)co * C_in + ci) * k + j]);
  }
}

// weight_norm(dim=2): w[co,ci,j] = g[j] * v[co,ci,j] / ||v[:,:,j]||_2   (torch.nn.utils.weight_norm)
// One CTA per tap j; output layout [group][j][ci][co_in_group] so that consecutive output channels are
// consecutive in memory.
__global__ void __launch_bounds__(256)
pack_posconv_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ dst, int D,
                    int groups, int k) {
  __shared__ float s_red[256];
  const int j = blockIdx.x, Dg = D / groups;
  const int64_t n = (int64_t)D * Dg;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) { const float x = v[i * k + j]; s = fmaf(x, x, s); }
  s_red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  const float scale = g[j] / sqrtf(s_red[0]);
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const int ci = (int)(i % Dg), co = (int)(i / Dg);
    const int grp = co / Dg, cog = co % Dg;
    dst[(((size_t)grp * k + j) * Dg + ci) * Dg + cog] = v[i * k + j] * scale;
  }
}

// y[b,t,co] = gelu(bias[co] + sum_{j<k} sum_{ci<Dg} w[g][j][ci][cog] * x[b, t + j - k/2, g*Dg + ci]),
// x = features with padded frames zeroed; output frame T of the (even k) padded conv is dropped (SamePad).
__global__ void __launch_bounds__(256)
posconv_kernel(const float* __restrict__ feats, int feat_rows, const uint8_t* __restrict__ frame_pad,
               const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int T,
               int D, int k, int groups, int t_per_cta) {
  const int Dg = D / groups;
  const int grp = blockIdx.y, b = blockIdx.z;
  const int t_begin = blockIdx.x * t_per_cta;
  const int n_out = t_per_cta * Dg;
  for (int idx = threadIdx.x; idx < n_out; idx += blockDim.x) {
    const int cog = idx % Dg, t = t_begin + idx / Dg;
    if (t >= T) continue;
    float acc = bias[grp * Dg + cog];
    for (int j = 0; j < k; ++j) {
      const int ts = t + j - k / 2;
      if (ts < 0 || ts >= T) continue;
      if (frame_pad && frame_pad[(size_t)b * T + ts]) continue;
      const float* xr = feats + ((size_t)b * feat_rows + ts) * D + grp * Dg;
      const float* wr = w + (((size_t)grp * k + j) * Dg) * Dg + cog;
      for (int ci = 0; ci < Dg; ++ci) acc = fmaf(wr[(size_t)ci * Dg], xr[ci], acc);
    }
    out[((size_t)b * T + t) * D + grp * Dg + cog] = gelu_erf(acc);
  }
}

// ---- tensor-core path of the positional conv (bf16 models) --------------------------------------------------------
// Group g is the implicit GEMM  out[b*Tp + t, g*Dg + o] = sum_{j<k, i<Dgp} xg[g][b*Tp + t + j][i] * wg[g][o][j*Dgp + i]:
// with the group's channels stored contiguously per frame, row m of the A operand is the k*Dgp contiguous elements
// starting at frame m (the wrapped-row view the conv stack uses, lda = Dgp < K = k*Dgp), and the k/2 zero frames in
// front of every utterance are SamePad's padding.
//
// folded fp32 weights [grp][j][ci][cog] -> bf16 [grp][cog][j*Dgp + ci], zero for ci >= Dg
__global__ void pack_posconv_tc_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Dg, int Dgp,
                                       int groups, int k) {
  const int64_t n = (int64_t)groups * Dg * k * Dgp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Dgp);
    const int j = (int)((i / Dgp) % k);
    const int cog = (int)((i / ((int64_t)Dgp * k)) % Dg);
    const int grp = (int)(i / ((int64_t)Dgp * k * Dg));
    dst[i] = ci < Dg ? from_f32<bf16>(w[(((size_t)grp * k + j) * Dg + ci) * Dg + cog]) : from_f32<bf16>(0.f);
  }
}

// xg[grp][b*Tp + u][c] = x[b, u - k/2, grp*Dg + c] for a real, unpadded frame and c < Dg, else 0; 8 channels per thread
__global__ void posconv_pack_x_kernel(const float* __restrict__ feats, int feat_rows,
                                      const uint8_t* __restrict__ frame_pad, uint4* __restrict__ xg, int B, int T,
                                      int Tp, int D, int Dg, int Dgp, int k, int64_t rows_tot, int64_t total) {
  const int v_per_row = Dgp / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % v_per_row) * 8;
    const int64_t row = (i / v_per_row) % rows_tot;
    const int grp = (int)(i / ((int64_t)v_per_row * rows_tot));
    const int b = (int)(row / Tp), t = (int)(row % Tp) - k / 2;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (b < B && t >= 0 && t < T && c < Dg && !(frame_pad != nullptr && frame_pad[(size_t)b * T + t])) {
      const float4* src = reinterpret_cast<const float4*>(feats + ((size_t)b * feat_rows + t) * D + grp * Dg + c);
      const float4 lo = src[0], hi = src[1];
      o = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
    }
    xg[i] = o;
  }
}
}  // namespace

static int grid_for(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  return (int)(g < 1 ? 1 : (g > 8192 ? 8192 : g));
}

w2vs_status_t launch_pack_slabs(const float* src, void* dst, int N, int K, int KC, cudaStream_t st) {
  pack_slabs_kernel<<<grid_for((int64_t)N * K), 256, 0, st>>>(src, (bf16*)dst, N, K, KC);
  W2VS_CHECK_LAUNCH("pack_slabs_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_pack_copy(const float* src, void* dst, int dst_dtype, int64_t n, cudaStream_t st) {
  if (n <= 0) return W2VS_OK;
  if (dst_dtype == W2VS_F32) pack_copy_kernel<float><<<grid_for(n), 256, 0, st>>>(src, (float*)dst, n);
  else pack_copy_kernel<bf16><<<grid_for(n), 256, 0, st>>>(src, (bf16*)dst, n);
  W2VS_CHECK_LAUNCH("pack_copy_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_pack_conv(const float* src, void* dst, int dst_dtype, int C_out, int C_in, int k,
                               cudaStream_t st) {
  const int64_t n = (int64_t)C_out * C_in * k;
  if (dst_dtype == W2VS_F32) pack_conv_kernel<float><<<grid_for(n), 256, 0, st>>>(src, (float*)dst, C_out, C_in, k);
  else pack_conv_kernel<bf16><<<grid_for(n), 256, 0, st>>>(src, (bf16*)dst, C_out, C_in, k);
  W2VS_CHECK_LAUNCH("pack_conv_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_pack_posconv(const float* g, const float* v, float* dst, int D, int groups, int k,
                                  cudaStream_t st) {
  pack_posconv_kernel<<<k, 256, 0, st>>>(g, v, dst, D, groups, k);
  W2VS_CHECK_LAUNCH("pack_posconv_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_pack_posconv_tc(const float* w_folded, void* dst, int D, int groups, int k, int Dgp,
                                     cudaStream_t st) {
  const int Dg = D / groups;
  pack_posconv_tc_kernel<<<grid_for((int64_t)groups * Dg * k * Dgp), 256, 0, st>>>(w_folded, (bf16*)dst, Dg, Dgp, groups, k);
  W2VS_CHECK_LAUNCH("pack_posconv_tc_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_posconv_pack_x(const float* feats, int feat_rows, const uint8_t* frame_pad, void* xg, int B,
                                    int T, int D, int k, int groups, int Dgp, cudaStream_t st) {
  const int Dg = D / groups, Tp = T + k;
  W2VS_REQUIRE(Dg % 8 == 0 && Dgp % 64 == 0 && Dgp >= Dg, "positional conv group width");
  const int64_t rows_tot = (int64_t)B * Tp + k;
  const int64_t total = (int64_t)groups * rows_tot * (Dgp / 8);
  posconv_pack_x_kernel<<<grid_for(total), 256, 0, st>>>(feats, feat_rows, frame_pad, (uint4*)xg, B, T, Tp, D, Dg, Dgp, k,
                                                        rows_tot, total);
  W2VS_CHECK_LAUNCH("posconv_pack_x_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_posconv(const PosConvArgs& a, cudaStream_t st) {
  const int t_per_cta = 16;
  dim3 grid((unsigned)ceil_div64(a.T, t_per_cta), (unsigned)a.groups, (unsigned)a.B);
  posconv_kernel<<<grid, 256, 0, st>>>(a.feats, a.feat_rows, a.frame_pad, a.w, a.bias, a.out, a.T, a.D, a.k,
                                       a.groups, t_per_cta);
  W2VS_CHECK_LAUNCH("posconv_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
