// Weight-streaming GEMM for a handful of rows (incremental inference: 24 tokens per stream and decision step):
//     C[M,N] = A[M,K] . W[N,K]^T + bias  (+GELU)  (+fp32 residual, in place)        M <= 64, bf16 operands
//
// Same contract as k_gemm_tc2.cu (nn.Linear of wav2vec2.py:568,950-973 / multihead_attention.py:162-194) for the
// shapes where a 256-row tensor-core tile is almost empty: the product is bound by reading W once from HBM
// (613 MB of weights per decision step for the large model), so the kernel is laid out to get every byte of W in
// flight at once and to spread it over all SMs:
//   * one CTA per 8 output columns, 8 warps per CTA; warp w owns the K range [w K/8, (w+1) K/8) of those 8 rows of
//     W and issues all of its 16-byte loads before the first MMA (K = 4096: 16 loads = 64 registers per lane);
//   * mma.sync.m16n8k16 (bf16, fp32 accumulate).  The k index inside a 32-wide block is permuted identically for
//     A and W (a dot product does not care): lane (g = lane/4, q = lane%4) takes the 8 consecutive elements
//     k = 32 blk + 8 q .. + 7 of row g, which feed two k16 steps -- so both operands are read with plain 16-byte
//     vector loads and no shared-memory staging; A (at most 64 x K bf16) is re-read by every CTA through L1/L2;
//   * the 8 partial accumulators are reduced through shared memory in a fixed order (deterministic), then bias,
//     GELU, residual and the store.
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
namespace {

constexpr int SK_WARPS = 8, SK_THREADS = 32 * SK_WARPS, SK_MAX_MT = 4, SK_MAX_BLK = 16;   // K/8/32 <= 16: K <= 4096

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MT, typename TC>
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ W, const float* __restrict__ bias,
                   int has_residual, TC* C, int64_t ldc, int M, int N, int K, int gelu) {
  __shared__ float part[SK_WARPS][MT * 16][8 + 1];
  // No griddepcontrol.launch_dependents here: the successors of a product (LayerNorm, step attention) have nothing
  // to fetch ahead of their dependency, and released early they only get in the way -- one stream through the chain
  // 1.354 -> 1.263 ms per step without it (same finding as for the tcgen05 products).  The kernels in
  // FRONT of a product (LayerNorm, attention) do release early: this kernel's weight loads below run ahead of its wait.
#ifndef W2VS_SKINNY_TRIGGER
#define W2VS_SKINNY_TRIGGER 0
#endif
  if (W2VS_SKINNY_TRIGGER) pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int n0 = blockIdx.x * 8;
  const int kw = K / SK_WARPS, nblk = kw / 32;          // this warp's K range, in blocks of 32
  const int kbeg = warp * kw + 8 * q;

  // ---- W: every load of this warp's slice is issued up front (rows n0 + g, 16 bytes per lane and block)
  const bf16* wrow = W + (size_t)(n0 + g) * K + kbeg;
  uint4 wv[SK_MAX_BLK];
#pragma unroll
  for (int b = 0; b < SK_MAX_BLK; ++b)
    if (b < nblk) wv[b] = __ldg(reinterpret_cast<const uint4*>(wrow + 32 * b));
  // The weights do not depend on the previous kernel: under a programmatic dependent launch they stream in while
  // it is still running; A (and C, for the residual) are touched only after the dependency has resolved.
  pdl_wait();

  float acc[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[mt][e] = 0.f;
  // rows past M are clamped (their results are dropped in the epilogue)
  const bf16* arow[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    arow[mt][0] = A + (size_t)min(mt * 16 + g, M - 1) * lda + kbeg;
    arow[mt][1] = A + (size_t)min(mt * 16 + g + 8, M - 1) * lda + kbeg;
  }
#pragma unroll
  for (int b = 0; b < SK_MAX_BLK; ++b) {
    if (b < nblk) {
      uint4 a_lo[MT], a_hi[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        a_lo[mt] = *reinterpret_cast<const uint4*>(arow[mt][0] + 32 * b);
        a_hi[mt] = *reinterpret_cast<const uint4*>(arow[mt][1] + 32 * b);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // step 0: elements 0..3 of the lane's 8 (a0 = row g {0,1}, a1 = row g+8 {0,1}, a2 = row g {2,3}, a3 = row g+8 {2,3})
        mma_bf16_16816(acc[mt], a_lo[mt].x, a_hi[mt].x, a_lo[mt].y, a_hi[mt].y, wv[b].x, wv[b].y);
        // step 1: elements 4..7
        mma_bf16_16816(acc[mt], a_lo[mt].z, a_hi[mt].z, a_lo[mt].w, a_hi[mt].w, wv[b].z, wv[b].w);
      }
    }
  }
  // ---- accumulator fragment: c0,c1 = row g, cols 2q, 2q+1; c2,c3 = row g+8
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    part[warp][mt * 16 + g][2 * q] = acc[mt][0];
    part[warp][mt * 16 + g][2 * q + 1] = acc[mt][1];
    part[warp][mt * 16 + g + 8][2 * q] = acc[mt][2];
    part[warp][mt * 16 + g + 8][2 * q + 1] = acc[mt][3];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < MT * 16 * 8; o += SK_THREADS) {
    const int r = o >> 3, c = o & 7;
    if (r >= M) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += part[w][r][c];       // fixed order
    if (bias != nullptr) v += bias[n0 + c];
    if (gelu) v = gelu1<TC>(v);
    TC* dst = C + (size_t)r * ldc + n0 + c;
    if (has_residual) v += to_f32(*dst);
    *dst = from_f32<TC>(v);
  }
}

template <typename TC>
w2vs_status_t launch_mt(const GemmArgs& g, cudaStream_t st) {
  const int mt = (g.M + 15) / 16;
  const dim3 grid((unsigned)(g.N / 8));
  const int has_res = g.residual != nullptr ? 1 : 0, gelu = (g.flags & W2VS_EPI_GELU) ? 1 : 0;
#define W2VS_SK_CASE(MT_)                                                                                   \
  launch_pdl(gemm_skinny_kernel<MT_, TC>, grid, dim3(SK_THREADS), 0, st, (const bf16*)g.A, g.lda, (const bf16*)g.W, \
             g.bias, has_res, (TC*)g.C, g.ldc, g.M, g.N, g.K, gelu)
  switch (mt) {
    case 1: W2VS_SK_CASE(1); break;
    case 2: W2VS_SK_CASE(2); break;
    case 3: W2VS_SK_CASE(3); break;
    default: W2VS_SK_CASE(4); break;
  }
#undef W2VS_SK_CASE
  if (g_prof_on) {
    char name[96];
    snprintf(name, sizeof(name), "gemm_skinny_kernel[M=%d,N=%d,K=%d,%s%s%s]", g.M, g.N, g.K,
             sizeof(TC) == 4 ? "f32" : "bf16", g.residual ? ",res" : "", gelu ? ",gelu" : "");
    W2VS_CHECK_LAUNCH(name);
  } else {
    W2VS_CHECK_LAUNCH("gemm_skinny_kernel");
  }
  return W2VS_OK;
}

}  // namespace

bool gemm_skinny_applicable(const GemmArgs& g) {
  return g.dtype_ab == W2VS_BF16 && g.M >= 1 && g.M <= 16 * SK_MAX_MT && g.K <= g.lda && g.K % (32 * SK_WARPS) == 0 &&
         g.K <= 32 * SK_WARPS * SK_MAX_BLK && g.N % 8 == 0 && g.lda % 8 == 0 &&
         (g.residual == nullptr || (g.dtype_c == W2VS_F32 && g.residual == (const float*)g.C)) &&
         ((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0;
}

w2vs_status_t launch_gemm_skinny(const GemmArgs& g, cudaStream_t st) {
  W2VS_REQUIRE(gemm_skinny_applicable(g), "skinny GEMM: M <= 64, bf16 operands, K % 256 == 0, K <= 4096, plain rows");
  return g.dtype_c == W2VS_F32 ? launch_mt<float>(g, st) : launch_mt<bf16>(g, st);
}

}  // namespace w2vs
