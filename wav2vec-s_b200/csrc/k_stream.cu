// Small data-movement kernels of incremental mode (stream.cu): carry/append row buffers of the conv
// stack, and the per-layer K/V cache append.
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
namespace {

// dst[b][0:cA) = srcA[b][offA:offA+cA);  dst[b][cA:cA+nB) = srcB[b][0:nB)   rows of `row_vecs` uint4 each
__global__ void concat_rows_kernel(uint4* __restrict__ dst, int64_t dst_bs, const uint4* __restrict__ srcA,
                                   int64_t a_bs, int offA, int cA, const uint4* __restrict__ srcB, int64_t b_bs,
                                   int nB, int row_vecs, int B) {
  pdl_prologue();
  const int64_t per_b = (int64_t)(cA + nB) * row_vecs;
  const int64_t total = per_b * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const int64_t r = i % per_b;
    const int row = (int)(r / row_vecs), v = (int)(r % row_vecs);
    const uint4 val = row < cA ? srcA[(size_t)b * a_bs + (size_t)(offA + row) * row_vecs + v]
                               : srcB[(size_t)b * b_bs + (size_t)(row - cA) * row_vecs + v];
    dst[(size_t)b * dst_bs + (size_t)row * row_vecs + v] = val;
  }
}

// waveform carry buffer (fp32): dst[b][0:cA) = srcA[b][offA:..), dst[b][cA:cA+n) = float(new[b][0:n))
template <typename TIn>
__global__ void concat_wav_kernel(float* __restrict__ dst, int64_t dst_bs, const float* __restrict__ srcA,
                                  int64_t a_bs, int offA, int cA, const TIn* __restrict__ src_new, int64_t new_bs,
                                  int n, int B) {
  pdl_prologue();
  const int64_t per_b = (int64_t)cA + n;
  const int64_t total = per_b * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const int r = (int)(i % per_b);
    dst[(size_t)b * dst_bs + r] = r < cA ? srcA[(size_t)b * a_bs + offA + r] : to_f32(src_new[(size_t)b * new_bs + (r - cA)]);
  }
}

// cache[b][row0 + j][0:2D) = qkv[b*n_tok + j][D:3D)   (K | V of this step's tokens)
__global__ void kv_append_kernel(const uint4* __restrict__ qkv, uint4* __restrict__ cache, int64_t cache_bs_vecs,
                                 int row0, int n_tok, int d_vecs /* D*sizeof/16 */, int B) {
  pdl_prologue();
  const int64_t per_b = (int64_t)n_tok * 2 * d_vecs;
  const int64_t total = per_b * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const int64_t r = i % per_b;
    const int j = (int)(r / (2 * d_vecs)), v = (int)(r % (2 * d_vecs));
    cache[(size_t)b * cache_bs_vecs + (size_t)(row0 + j) * 2 * d_vecs + v] =
        qkv[((size_t)b * n_tok + j) * 3 * d_vecs + d_vecs + v];
  }
}

inline int grid_for(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  return (int)(g < 1 ? 1 : (g > 2368 ? 2368 : g));
}
}  // namespace

w2vs_status_t launch_concat_rows(void* dst, int64_t dst_bs_bytes, const void* srcA, int64_t a_bs_bytes, int offA,
                                 int cA, const void* srcB, int64_t b_bs_bytes, int nB, int row_bytes, int B,
                                 cudaStream_t st) {
  W2VS_REQUIRE(row_bytes % 16 == 0 && dst_bs_bytes % 16 == 0 && a_bs_bytes % 16 == 0 && b_bs_bytes % 16 == 0,
               "concat_rows alignment");
  const int64_t total = (int64_t)(cA + nB) * (row_bytes / 16) * B;
  if (total <= 0) return W2VS_OK;
  launch_pdl(concat_rows_kernel, dim3(grid_for(total)), dim3(256), 0, st, (uint4*)dst, dst_bs_bytes / 16,
             (const uint4*)srcA, a_bs_bytes / 16, offA, cA, (const uint4*)srcB, b_bs_bytes / 16, nB, row_bytes / 16, B);
  W2VS_CHECK_LAUNCH("concat_rows_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_concat_wav(float* dst, int64_t dst_bs, const float* srcA, int64_t a_bs, int offA, int cA,
                                const void* src_new, int new_dtype, int64_t new_bs, int n, int B, cudaStream_t st) {
  const int64_t total = ((int64_t)cA + n) * B;
  if (total <= 0) return W2VS_OK;
  if (new_dtype == W2VS_F32)
    launch_pdl(concat_wav_kernel<float>, dim3(grid_for(total)), dim3(256), 0, st, dst, dst_bs, srcA, a_bs, offA, cA, (const float*)src_new, new_bs, n, B);
  else if (new_dtype == W2VS_I16)   // 16-bit PCM chunks straight from the audio source (x / 32768)
    launch_pdl(concat_wav_kernel<int16_t>, dim3(grid_for(total)), dim3(256), 0, st, dst, dst_bs, srcA, a_bs, offA, cA, (const int16_t*)src_new, new_bs, n, B);
  else
    launch_pdl(concat_wav_kernel<bf16>, dim3(grid_for(total)), dim3(256), 0, st, dst, dst_bs, srcA, a_bs, offA, cA, (const bf16*)src_new, new_bs, n, B);
  W2VS_CHECK_LAUNCH("concat_wav_kernel");
  return W2VS_OK;
}

w2vs_status_t launch_kv_append(const void* qkv, void* cache, int64_t cache_rows, int row0, int n_tok, int D,
                               int elem_bytes, int B, cudaStream_t st) {
  W2VS_REQUIRE((D * elem_bytes) % 16 == 0, "kv_append alignment");
  const int d_vecs = D * elem_bytes / 16;
  const int64_t total = (int64_t)n_tok * 2 * d_vecs * B;
  if (total <= 0) return W2VS_OK;
  launch_pdl(kv_append_kernel, dim3(grid_for(total)), dim3(256), 0, st, (const uint4*)qkv, (uint4*)cache,
             cache_rows * 2 * d_vecs, row0, n_tok, d_vecs, B);
  W2VS_CHECK_LAUNCH("kv_append_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
