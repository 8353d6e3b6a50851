// Block-mask-aware fused attention, fp32 CUDA-core version (fp32 parity mode; reference for the
// tensor-core version).
//
// Replaces MultiheadAttention fast path -> F.multi_head_attention_forward with the dense additive
// mask of gen_block_attn_mask (modules/multihead_attention.py:162-194, wav2vec_S.py:444-489).  The
// mask is never materialised: visibility of key token k for query token q is computed from
// (T', main, rc):   main key   k <  T' : block(k) <= qblock(q)
//                   look-ahead k >= T' : owner(k) == qblock(q)
// with qblock(q) = q/main for main tokens and owner(q) for look-ahead copies.  Key tiles that are
// fully masked for a whole query tile are skipped (the additive -1e4 of the reference underflows to
// exactly 0 after softmax, so skipping is numerically identical); key padding (-inf) is applied per
// element.  Online softmax in fp32, scale head_dim^-0.5 applied to q.
#include <math.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
namespace {

constexpr int QT = 64, KT = 64, HD = 64, LDT = 68;  // tiles and padded smem row (floats)

template <typename T>
__device__ __forceinline__ void load_tile(float (*dst)[LDT], const T* __restrict__ src, int64_t row_stride,
                                          int first_row, int n_rows, float scale, int tid) {
  // 64 rows x 64 dims; chunk c -> row c/8, dims (c%8)*8..+7 (one 128 B row per 8 lanes for bf16)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int c = tid + 256 * r;
    const int row = c >> 3, d0 = (c & 7) * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (row < n_rows) load8(src + (size_t)(first_row + row) * row_stride + d0, v);
    *reinterpret_cast<float4*>(&dst[row][d0]) = make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
    *reinterpret_cast<float4*>(&dst[row][d0 + 4]) = make_float4(v[4] * scale, v[5] * scale, v[6] * scale, v[7] * scale);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
attn_simt_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ keypad, T* __restrict__ ctx,
                 int T2, int M, int main_ctx, int rc, int D, int n_main_tiles, float scale,
                 const T* __restrict__ kv_cache, int64_t kv_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*Qs)[LDT] = reinterpret_cast<float (*)[LDT]>(smem_raw);
  float (*Ks)[LDT] = Qs + QT;   // reused as Ps after the score tile is in registers
  float (*Vs)[LDT] = Ks + KT;
  __shared__ uint8_t s_kpad[KT];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nb = T2 / main_ctx;
  // step mode (kv_cache != nullptr): M = query tokens per stream, T2 = keys visible to all of them
  const bool step = kv_cache != nullptr;
  const int64_t rs = 3 * (int64_t)D;  // token row stride in qkv
  const int64_t krs = step ? 2 * (int64_t)D : rs;
  const T* qbase = qkv + (size_t)b * M * rs + (size_t)h * HD;
  const T* kbase = step ? kv_cache + (size_t)b * kv_rows * krs + (size_t)h * HD : qbase + D;
  const T* vbase = kbase + D;

  // ---- query tile
  int q_first, q_count;
  if (step) {
    q_first = blockIdx.x * QT;
    q_count = min(QT, M - q_first);
  } else if ((int)blockIdx.x < n_main_tiles) {
    q_first = blockIdx.x * QT;
    q_count = min(QT, T2 - q_first);
  } else {
    q_first = T2 + (blockIdx.x - n_main_tiles) * QT;
    q_count = min(QT, M - q_first);
  }
  const int rcd = rc > 0 ? rc : 1;
  auto qblock = [&](int m) { return step ? 0 : (m < T2 ? m / main_ctx : (m - T2) / rcd); };
  const int qb_lo = qblock(q_first), qb_hi = qblock(q_first + q_count - 1);
  // visible key segments for the whole tile
  const int seg0_end = step ? T2 : min(main_ctx * (qb_hi + 1), T2);  // main keys [0, seg0_end)
  int seg1_begin = 0, seg1_end = 0;                                   // look-ahead keys
  if (!step && rc > 0 && qb_lo <= nb - 1) {
    seg1_begin = T2 + rc * qb_lo;
    seg1_end = T2 + rc * (min(qb_hi, nb - 1) + 1);
  }

  load_tile(Qs, qbase, rs, q_first, q_count, scale, tid);

  int my_qb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = ty + 16 * i;
    my_qb[i] = q < q_count ? qblock(q_first + q) : -1;
  }
  float o[4][4], m_run[4], l_run[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int seg = 0; seg < 2; ++seg) {
    const int s_begin = seg == 0 ? 0 : seg1_begin;
    const int s_end = seg == 0 ? seg0_end : seg1_end;
    for (int k0 = s_begin; k0 < s_end; k0 += KT) {
      const int k_count = min(KT, s_end - k0);
      __syncthreads();  // previous tile's Ps / Vs fully consumed (also orders the Q load)
      load_tile(Ks, kbase, krs, k0, k_count, 1.0f, tid);
      load_tile(Vs, vbase, krs, k0, k_count, 1.0f, tid);
      if (tid < KT) s_kpad[tid] = tid < k_count ? (step ? 0 : keypad[(size_t)b * M + k0 + tid]) : 1;
      __syncthreads();

      // ---- S = Q K^T for queries ty+16i, keys tx+16j
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
      for (int d = 0; d < HD; d += 4) {
        float4 qv[4], kv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[ty + 16 * i][d]);
#pragma unroll
        for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(&Ks[tx + 16 * j][d]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
            s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
            s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
            s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
          }
      }
      // ---- mask
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = tx + 16 * j;
        const int k = k0 + kk;
        const bool kpad = s_kpad[kk] != 0;
        const int kb = step ? 0 : (seg == 0 ? k / main_ctx : (k - T2) / rcd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool vis = !kpad && my_qb[i] >= 0 && (seg == 0 ? kb <= my_qb[i] : kb == my_qb[i]);
          if (!vis) s[i][j] = -INFINITY;
        }
      }
      // ---- online softmax (row statistics shared by the 16 lanes with equal ty)
      float alpha[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float m_new = fmaxf(m_run[i], mx);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        alpha[i] = __expf(m_run[i] - m_safe);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = __expf(s[i][j] - m_safe);
          sum += s[i][j];
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        l_run[i] = l_run[i] * alpha[i] + sum;
        m_run[i] = m_new;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] *= alpha[i];
      }
      __syncthreads();  // all reads of Ks done -> reuse as Ps[q][k]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Ks[ty + 16 * i][tx + 16 * j] = s[i][j];
      __syncthreads();
      // ---- O += P V : queries ty+16i, dims tx*4..+3
#pragma unroll 4
      for (int k = 0; k < KT; k += 4) {
        float4 pv[4], vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(&Ks[ty + 16 * i][k]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) vv[kk] = *reinterpret_cast<const float4*>(&Vs[k + kk][tx * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p0 = pv[i].x, p1 = pv[i].y, p2 = pv[i].z, p3 = pv[i].w;
          o[i][0] = fmaf(p0, vv[0].x, fmaf(p1, vv[1].x, fmaf(p2, vv[2].x, fmaf(p3, vv[3].x, o[i][0]))));
          o[i][1] = fmaf(p0, vv[0].y, fmaf(p1, vv[1].y, fmaf(p2, vv[2].y, fmaf(p3, vv[3].y, o[i][1]))));
          o[i][2] = fmaf(p0, vv[0].z, fmaf(p1, vv[1].z, fmaf(p2, vv[2].z, fmaf(p3, vv[3].z, o[i][2]))));
          o[i][3] = fmaf(p0, vv[0].w, fmaf(p1, vv[1].w, fmaf(p2, vv[2].w, fmaf(p3, vv[3].w, o[i][3]))));
        }
      }
    }
  }

  // ---- write context rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = ty + 16 * i;
    if (q >= q_count) continue;
    const float inv = l_run[i] > 0.f ? 1.0f / l_run[i] : 0.f;
    T* dst = ctx + ((size_t)b * M + q_first + q) * D + (size_t)h * HD + tx * 4;
    const float v0 = o[i][0] * inv, v1 = o[i][1] * inv, v2 = o[i][2] * inv, v3 = o[i][3] * inv;
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
    } else {
      uint2 u;
      u.x = pack_bf16x2(v0, v1);
      u.y = pack_bf16x2(v2, v3);
      *reinterpret_cast<uint2*>(dst) = u;
    }
  }
}
}  // namespace

w2vs_status_t launch_attention_simt(const AttnArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(a.D == a.heads * HD, "attention head_dim must be 64");
  const bool step = a.n_step_q > 0;
  const int M = step ? a.n_step_q : a.T2 + (a.rc > 0 ? (a.T2 / a.main_ctx) * a.rc : 0);
  const int T2 = step ? a.n_step_keys : a.T2;
  const int n_main = step ? (M + QT - 1) / QT : (a.T2 + QT - 1) / QT;
  const int n_rc = step ? 0 : (M - a.T2 + QT - 1) / QT;
  dim3 grid((unsigned)(n_main + n_rc), (unsigned)a.heads, (unsigned)a.B);
  const size_t smem = (size_t)(QT + KT + KT) * LDT * sizeof(float);
  const float scale = 1.0f / sqrtf((float)HD);
  cudaError_t e;
  if (a.dtype == W2VS_F32) {
    e = cudaFuncSetAttribute(attn_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      attn_simt_kernel<float><<<grid, 256, smem, st>>>((const float*)a.qkv, a.keypad, (float*)a.ctx, T2, M,
                                                       step ? 1 : a.main_ctx, step ? 0 : a.rc, a.D, n_main, scale,
                                                       (const float*)a.kv_cache, a.kv_rows);
  } else {
    e = cudaFuncSetAttribute(attn_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      attn_simt_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)a.qkv, a.keypad, (bf16*)a.ctx, T2, M,
                                                      step ? 1 : a.main_ctx, step ? 0 : a.rc, a.D, n_main, scale,
                                                      (const bf16*)a.kv_cache, a.kv_rows);
  }
  if (e != cudaSuccess) { set_error("attn_simt attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  W2VS_CHECK_LAUNCH("attn_simt_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
