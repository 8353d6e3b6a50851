// fp32-accurate GEMM on the CUDA cores:  C[M,N] = A[M,K] . W[N,K]^T + bias (+GELU) (+residual).
//
// This is the arithmetic path of the fp32 parity mode (tolerance 1e-4 rules out bf16/tf32 tensor
// cores), and the small-M path of incremental inference.  Both operands are K-major, exactly the
// nn.Linear layout; `lda` may be smaller than K (overlapping rows) which is how the strided convs
// k=3/s=2 and k=2/s=2 become plain GEMMs over the channels-last activation.
// Classic 128x128x16 register-tiled kernel, 8x8 outputs per thread, register-prefetch double
// buffering; bf16 operands are widened on load and accumulated in fp32.
#include "common.cuh"
#include "kernels.h"

namespace w2vs {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4;

template <typename TA, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int64_t lda, const TA* __restrict__ W, int64_t ldw,
                 const float* __restrict__ bias, const float* residual, TC* C, int64_t ldc, int M,
                 int N, int K, int gelu) {
  __shared__ __align__(16) float As[2][BK][LDS_];
  __shared__ __align__(16) float Bs[2][BK][LDS_];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // loader mapping: row = tid/2 (0..127), 8 consecutive k at (tid%2)*8
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  const bool a_ok = (m0 + lrow) < M, b_ok = (n0 + lrow) < N;
  const TA* a_ptr = A + (size_t)(m0 + lrow) * lda + lk;
  const TA* b_ptr = W + (size_t)(n0 + lrow) * ldw + lk;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto gload = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { ra[e] = 0.f; rb[e] = 0.f; }
    if (k0 + lk < K) {  // K % 8 == 0 is required, so a chunk is either fully in or fully out
      if (a_ok) load8(a_ptr + k0, ra);
      if (b_ok) load8(b_ptr + k0, rb);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      As[buf][lk + e][lrow] = ra[e];
      Bs[buf][lk + e][lrow] = rb[e];
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {ty*4..+3, 64+ty*4..+3}, cols {tx*4..+3, 64+tx*4..+3}
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= N) continue;  // N % 4 == 0
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][jh * 4 + j] + (bias ? bias[col + j] : 0.f);
        if (gelu) v[j] = gelu_erf(v[j]);
      }
      if (residual) {
        const float4 r = *reinterpret_cast<const float4*>(residual + (size_t)row * ldc + col);
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
      TC* cp = C + (size_t)row * ldc + col;
      if (sizeof(TC) == 4) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        uint2 u;
        u.x = pack_bf16x2(v[0], v[1]);
        u.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(cp) = u;
      }
    }
  }
}
}  // namespace

w2vs_status_t launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  W2VS_REQUIRE(g.K % 8 == 0 && g.N % 4 == 0, "GEMM needs K % 8 == 0 and N % 4 == 0");
  W2VS_REQUIRE(g.lda % 8 == 0 && g.ldc % 4 == 0, "GEMM leading dims alignment");
  if (g.M <= 0) return W2VS_OK;
  dim3 grid((unsigned)ceil_div64(g.N, BN), (unsigned)ceil_div64(g.M, BM));
  const int gelu = (g.flags & W2VS_EPI_GELU) ? 1 : 0;
#define W2VS_SIMT(TA, TC)                                                                    \
  gemm_simt_kernel<TA, TC><<<grid, 256, 0, st>>>((const TA*)g.A, g.lda, (const TA*)g.W, g.K, \
                                                 g.bias, g.residual, (TC*)g.C, g.ldc, g.M, g.N, g.K, gelu)
  if (g.dtype_ab == W2VS_F32 && g.dtype_c == W2VS_F32) W2VS_SIMT(float, float);
  else if (g.dtype_ab == W2VS_F32) W2VS_SIMT(float, bf16);
  else if (g.dtype_c == W2VS_F32) W2VS_SIMT(bf16, float);
  else W2VS_SIMT(bf16, bf16);
#undef W2VS_SIMT
  W2VS_CHECK_LAUNCH("gemm_simt_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
