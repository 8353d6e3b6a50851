// Block-mask-aware fused attention on the tensor cores (bf16 operands, fp32 softmax/accumulate).
//
// Same contract as k_attn_simt.cu (MultiheadAttention fast path + gen_block_attn_mask,
// modules/multihead_attention.py:162-194, wav2vec_S.py:444-489): token buffer qkv [B, M, 3D],
// M = T' + nb*rc, mask derived from (T', main, rc) and the key-padding bytes; key tiles that are
// invisible to a whole query tile are never loaded.
//
// Flash-style schedule: one CTA = 64 query tokens of one (utterance, head), 4 warps x 16 rows.  K/V
// tiles of 64 keys stream through a 2-stage cp.async pipeline into XOR-swizzled shared memory,
// S = Q K^T and O += P V run as mma.sync.m16n8k16 (ldmatrix-fed), the running max / sum live in
// registers.  Query tiles are issued heaviest-first (late blocks see the most keys).
#include <math.h>
#include <limits.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
namespace {

constexpr int QT = 64, KT = 64, HD = 64;
constexpr int TILE_ELEMS = 64 * 64;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element offset of (row, 16-byte chunk) inside a 64 x 64 bf16 tile with the chunk index XOR-swizzled
__device__ __forceinline__ int sw(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 rows x 128 B from a strided global matrix into a swizzled tile; rows >= n_rows are zero-filled.  Rows at or past
// `split` come from a second matrix (step mode: keys of this step's own tokens are read from the qkv buffer, earlier
// keys from the K/V cache).
__device__ __forceinline__ void load_tile_async(bf16* tile, const bf16* src, int64_t row_stride, int first_row,
                                                int n_rows, int tid, const bf16* src2 = nullptr,
                                                int64_t row_stride2 = 0, int split = INT_MAX) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = tid + 128 * r;
    const int row = c >> 3, chunk = c & 7;
    const bool ok = row < n_rows;
    const int ar = first_row + (ok ? row : 0);
    const bf16* g = ar < split ? src + (size_t)ar * row_stride + chunk * 8
                               : src2 + (size_t)(ar - split) * row_stride2 + chunk * 8;
    cp_async16(smem_addr(tile + sw(row, chunk)), g, ok);
  }
}

// step mode (kv_cache != nullptr): M = query tokens per stream, T2 = keys visible to all of them.
__global__ void __launch_bounds__(128)
attn_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ keypad, bf16* __restrict__ ctx,
                int T2, int M, int main_ctx, int rc, int D, int n_main_tiles, int n_tiles, float scale_log2,
                const bf16* __restrict__ kv_cache, int64_t kv_rows, int n_splits, float* __restrict__ partials,
                unsigned* __restrict__ counters) {
  __shared__ __align__(128) bf16 Qs[TILE_ELEMS];
  __shared__ int s_last;
  __shared__ __align__(128) bf16 Ks[2][TILE_ELEMS];
  __shared__ __align__(128) bf16 Vs[2][TILE_ELEMS];
  __shared__ int s_kinfo[2][KT];
  pdl_prologue();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  // step mode with n_splits > 1: blockIdx.x = query tile * n_splits + split; each split takes a contiguous range of
  // key tiles and leaves (m, l, O) per row, the last CTA of a (stream, head, query tile) to finish combines them
  const int tile_lin = (int)blockIdx.x / n_splits, split = (int)blockIdx.x % n_splits;
  const int tile_id = n_tiles - 1 - tile_lin;  // heaviest first
  const int nb = T2 / main_ctx;
  const int rcd = rc > 0 ? rc : 1;
  const bool step = kv_cache != nullptr;
  const int64_t rs = 3 * (int64_t)D;
  const int64_t krs = step ? 2 * (int64_t)D : rs;
  const bf16* qbase = qkv + (size_t)b * M * rs + (size_t)h * HD;
  const bf16* kbase = step ? kv_cache + (size_t)b * kv_rows * krs + (size_t)h * HD : qbase + D;
  const bf16* vbase = kbase + D;
  const uint8_t* kp = step ? nullptr : keypad + (size_t)b * M;
  // step mode: the M query tokens of this step are the last M of the T2 keys.  Their K / V are read from the qkv
  // buffer, and the first query tile's CTA of each (head, stream) also appends them to the cache for later steps
  // (what a separate kv_append launch used to do: one kernel less per layer in a latency-bound chain).
  const int f0 = step ? T2 - M : INT_MAX;
  if (step && blockIdx.x == 0) {      // (query tile 0, split 0)
    bf16* cache = const_cast<bf16*>(kv_cache) + (size_t)b * kv_rows * krs + (size_t)h * HD;
    for (int i = tid; i < M * 16; i += 128) {          // 8 chunks of K + 8 chunks of V per token
      const int row = i >> 4, part = (i >> 3) & 1, chunk = i & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(qbase + (size_t)row * rs + (1 + part) * D + chunk * 8);
      *reinterpret_cast<uint4*>(cache + (size_t)(f0 + row) * krs + part * D + chunk * 8) = v;
    }
  }

  int q_first, q_count;
  if (step) { q_first = tile_id * QT; q_count = min(QT, M - q_first); }
  else if (tile_id < n_main_tiles) { q_first = tile_id * QT; q_count = min(QT, T2 - q_first); }
  else { q_first = T2 + (tile_id - n_main_tiles) * QT; q_count = min(QT, M - q_first); }
  auto qblock = [&](int m) { return step ? 0 : (m < T2 ? m / main_ctx : (m - T2) / rcd); };
  const int qb_lo = qblock(q_first), qb_hi = qblock(q_first + q_count - 1);
  const int seg0_end = step ? T2 : min(main_ctx * (qb_hi + 1), T2);
  int seg1_begin = 0, seg1_end = 0;
  if (!step && rc > 0 && qb_lo <= nb - 1) { seg1_begin = T2 + rc * qb_lo; seg1_end = T2 + rc * (min(qb_hi, nb - 1) + 1); }
  const int n0 = (seg0_end + KT - 1) / KT;
  const int n1 = (seg1_end - seg1_begin + KT - 1) / KT;
  const int n_kt = n0 + n1;
  int it_begin = 0, it_end = n_kt;
  if (n_splits > 1) {
    const int per = (n_kt + n_splits - 1) / n_splits;
    it_begin = min(split * per, n_kt);
    it_end = min(it_begin + per, n_kt);
  }

  auto issue_tile = [&](int it, int buf) {
    const bool s1 = it >= n0;
    const int k0 = s1 ? seg1_begin + (it - n0) * KT : it * KT;
    const int cnt = min(KT, (s1 ? seg1_end : seg0_end) - k0);
    load_tile_async(Ks[buf], kbase, krs, k0, cnt, tid, qbase + D, rs, f0);
    load_tile_async(Vs[buf], vbase, krs, k0, cnt, tid, qbase + 2 * D, rs, f0);
    if (tid < KT) {
      int info;
      if (step) info = tid < cnt ? 0 : INT_MAX;
      else if (tid < cnt && !kp[k0 + tid]) info = s1 ? (k0 + tid - T2) / rcd : (k0 + tid) / main_ctx;
      else info = s1 ? -2 : INT_MAX;
      s_kinfo[buf][tid] = info;
    }
  };

  load_tile_async(Qs, qbase, rs, q_first, q_count, tid);
  if (it_begin < it_end) issue_tile(it_begin, 0);
  cp_async_commit();

  // this thread's two query rows: warp*16 + g and + 8
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const int qb0 = r0 < q_count ? qblock(q_first + r0) : -1;
  const int qb1 = r1 < q_count ? qblock(q_first + r1) : -1;

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int it = it_begin; it < it_end; ++it) {
    const int buf = (it - it_begin) & 1;
    if (it + 1 < it_end) {
      issue_tile(it + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == it_begin) {
      const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ldsm_x4(smem_addr(Qs + sw(row, kk * 2 + (lane >> 4))), qf[kk]);
    }
    const bool seg1 = it >= n0;

    // ---- S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t kf[4];
        const int mi = lane >> 3;
        const int row = jp * 16 + (lane & 7) + (mi >> 1) * 8;
        ldsm_x4(smem_addr(Ks[buf] + sw(row, kk * 2 + (mi & 1))), kf);
        mma_bf16(s[2 * jp], qf[kk], kf[0], kf[1]);
        mma_bf16(s[2 * jp + 1], qf[kk], kf[2], kf[3]);
      }
    }
    // ---- mask + online softmax
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ki = s_kinfo[buf][j * 8 + 2 * t4 + e];
        const bool v0 = seg1 ? (ki == qb0) : (ki <= qb0);
        const bool v1 = seg1 ? (ki == qb1) : (ki <= qb1);
        s[j][e] = v0 ? s[j][e] : -INFINITY;
        s[j][2 + e] = v1 ? s[j][2 + e] : -INFINITY;
        mx0 = fmaxf(mx0, s[j][e]);
        mx1 = fmaxf(mx1, s[j][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float ms0 = mn0 == -INFINITY ? 0.f : mn0 * scale_log2;
    const float ms1 = mn1 == -INFINITY ? 0.f : mn1 * scale_log2;
    const float a0 = exp2f(m0 * scale_log2 - ms0), a1 = exp2f(m1 * scale_log2 - ms1);
    m0 = mn0; m1 = mn1;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pf[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(fmaf(s[j][0], scale_log2, -ms0)), p1 = exp2f(fmaf(s[j][1], scale_log2, -ms0));
      const float p2 = exp2f(fmaf(s[j][2], scale_log2, -ms1)), p3 = exp2f(fmaf(s[j][3], scale_log2, -ms1));
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      pf[j][0] = pack_bf16x2(p0, p1);
      pf[j][1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * a0 + sum0;
    l1 = l1 * a1 + sum1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }

    // ---- O += P V : keys are the contraction dimension
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t af[4] = {pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1]};
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t vf[4];
        const int mi = lane >> 3;
        const int row = kk * 16 + (lane & 7) + (mi & 1) * 8;
        ldsm_x4_trans(smem_addr(Vs[buf] + sw(row, jp * 2 + (mi >> 1))), vf);
        mma_bf16(o[2 * jp], af, vf[0], vf[1]);
        mma_bf16(o[2 * jp + 1], af, vf[2], vf[3]);
      }
    }
    __syncthreads();  // tile `buf` fully consumed before the next iteration's prefetch overwrites it
  }

  cp_async_wait<0>();   // an empty key range (a trailing split) never waited for its Q tile
  // ---- normalise, stage through this warp's 16 rows of Qs, 16-byte coalesced stores
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  if (n_splits > 1) {
    // ---- partial state of this split: per row 64 un-normalised outputs, the running maximum (raw score units)
    //      and the row sum; the last CTA of this (stream, head, query tile) merges the splits
    const size_t grp = ((size_t)b * gridDim.y + h) * n_tiles + tile_lin;
    float* P = partials + (grp * n_splits + split) * (size_t)(QT * 66);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float2*>(P + r0 * 66 + j * 8 + 2 * t4) = make_float2(o[j][0], o[j][1]);
      *reinterpret_cast<float2*>(P + r1 * 66 + j * 8 + 2 * t4) = make_float2(o[j][2], o[j][3]);
    }
    if (t4 == 0) {
      *reinterpret_cast<float2*>(P + r0 * 66 + 64) = make_float2(m0, l0);
      *reinterpret_cast<float2*>(P + r1 * 66 + 64) = make_float2(m1, l1);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(counters + grp, 1u) == (unsigned)n_splits - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* P0 = partials + grp * n_splits * (size_t)(QT * 66);
    for (int idx = tid; idx < q_count * 8; idx += 128) {
      const int row = idx >> 3, chunk = idx & 7;
      float mx = -INFINITY;
      for (int sp = 0; sp < n_splits; ++sp) mx = fmaxf(mx, __ldcg(P0 + (size_t)sp * QT * 66 + row * 66 + 64));
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, lsum = 0.f;
      for (int sp = 0; sp < n_splits; ++sp) {
        const float* pr = P0 + (size_t)sp * QT * 66 + row * 66;
        const float2 ml = __ldcg(reinterpret_cast<const float2*>(pr + 64));
        if (ml.x == -INFINITY) continue;
        const float w = exp2f((ml.x - mx) * scale_log2);
        lsum = fmaf(ml.y, w, lsum);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float2 ov = __ldcg(reinterpret_cast<const float2*>(pr + chunk * 8 + e));
          acc[e] = fmaf(ov.x, w, acc[e]);
          acc[e + 1] = fmaf(ov.y, w, acc[e + 1]);
        }
      }
      const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
      uint4 v;
      v.x = pack_bf16x2(acc[0] * inv, acc[1] * inv); v.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
      v.z = pack_bf16x2(acc[4] * inv, acc[5] * inv); v.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
      *reinterpret_cast<uint4*>(ctx + ((size_t)b * M + q_first + row) * D + (size_t)h * HD + chunk * 8) = v;
    }
    if (tid == 0) counters[grp] = 0u;            // ready for the next launch (stream order)
    return;
  }
  const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // columns j*8 + 2*t4, +1  -> chunk j, element offset 2*t4 inside the chunk
    *reinterpret_cast<uint32_t*>(Qs + sw(r0, j) + 2 * t4) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
    *reinterpret_cast<uint32_t*>(Qs + sw(r1, j) + 2 * t4) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = lane + 32 * r;
    const int row = warp * 16 + (c >> 3), chunk = c & 7;
    if (row < q_count) {
      const uint4 v = *reinterpret_cast<const uint4*>(Qs + sw(row, chunk));
      *reinterpret_cast<uint4*>(ctx + ((size_t)b * M + q_first + row) * D + (size_t)h * HD + chunk * 8) = v;
    }
  }
}
// Step mode with at most 32 query tokens per stream (a decision step has main + rc = 24): the four warps share the
// QUERY rows and split the KEYS of every tile instead (16 each), so a tile costs a warp 32 MMAs instead of 64 (and none
// of them on the 40 padding rows of a 64-row query tile) -- the kernel is a serial walk over up to 24 key tiles per
// (stream, head) at one or two CTAs per SM, so the length of that per-tile chain is its run time.  Every warp keeps its
// own running (max, sum, O) over its quarter of the keys; the four states are merged through shared memory at the end,
// then either written out or handed to the cross-CTA merge of the split-key scheme (same partial-state layout as
// attn_mma_kernel).
constexpr int SQ = 32;                 // query rows per CTA
constexpr int OP_LD = 68;              // floats per row of a warp's partial O in shared memory
// NG groups of four warps walk the key tiles in turns (group q takes tiles q, q + NG, ...), each with its own two-stage
// K/V ring, so NG tiles are in flight per CTA; all 4 * NG warp states are merged at the end.  NG = 2 keeps the Q
// fragments in shared memory (128 registers per thread for two CTAs of 256 threads per SM).
// NSTG: depth of a group's K/V ring (cp.async groups, NSTG - 1 tiles in flight ahead of the one being consumed): with
// 256 CTAs of one group (16 streams) two stages keep ~28 KB per SM in flight, less than half of what the HBM latency
// needs (the kernel read 76 MB at 36 % of the DRAM peak); four stages are used there.
template <int NG, int NSTG>
__global__ void __launch_bounds__(128 * NG, NG == 1 ? 3 : 2)
attn_step32_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ ctx, int T2, int M, int D, float scale_log2,
                   const bf16* __restrict__ kv_cache, int64_t kv_rows, int n_splits, float* __restrict__ partials,
                   unsigned* __restrict__ counters) {
  extern __shared__ __align__(128) uint8_t sm_raw[];                      // Q tile | [group][stage][K | V]; later the O partials
  __shared__ float s_ml[4 * NG][SQ][2];
  __shared__ int s_last;
  bf16* Qs = reinterpret_cast<bf16*>(sm_raw);
  float* Op = reinterpret_cast<float*>(sm_raw);                            // [4 * NG warps][SQ][OP_LD]
  static_assert(4 * NG * SQ * OP_LD * 4 <= (1 + 2 * NSTG * NG) * TILE_ELEMS * 2, "partial O does not fit");
  pdl_launch_dependents();

  const int tid = threadIdx.x, wall = tid >> 5, lane = tid & 31;
  const int grp = wall >> 2, warp = wall & 3, gtid = tid & 127;      // tile group, key quarter inside a tile
  const int g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z, split = blockIdx.x;
  bf16* KVs = Qs + TILE_ELEMS + (size_t)grp * 2 * NSTG * TILE_ELEMS;   // this group's [stage][K | V][64 x 64]
  auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };
  const int64_t rs = 3 * (int64_t)D, krs = 2 * (int64_t)D;
  const bf16* qbase = qkv + (size_t)b * M * rs + (size_t)h * HD;
  const bf16* kbase = kv_cache + (size_t)b * kv_rows * krs + (size_t)h * HD;
  const bf16* vbase = kbase + D;
  const int f0 = T2 - M;               // keys at or past f0 are this step's own tokens (read from qkv)
  const int n_kt = (T2 + KT - 1) / KT;
  int it_begin = 0, it_end = n_kt;
  if (n_splits > 1) {
    const int per = (n_kt + n_splits - 1) / n_splits;
    it_begin = min(split * per, n_kt);
    it_end = min(it_begin + per, n_kt);
  }
  auto issue_tile = [&](int it, int buf) {
    const int k0 = it * KT, cnt = min(KT, T2 - k0);
    load_tile_async(KVs + (buf * 2 + 0) * TILE_ELEMS, kbase, krs, k0, cnt, gtid, qbase + D, rs, f0);
    load_tile_async(KVs + (buf * 2 + 1) * TILE_ELEMS, vbase, krs, k0, cnt, gtid, qbase + 2 * D, rs, f0);
  };
  const int it_first = it_begin + grp;
  // Keys before f0 were cached by earlier steps: a first tile made of those alone is requested before this kernel
  // waits for its predecessor (the QKV product), everything that reads this step's qkv rows after.
  const bool early = it_first < it_end && (it_first + 1) * KT <= f0;
  if (early) issue_tile(it_first, 0);
  pdl_wait();
  if (blockIdx.x == 0) {               // append this step's K / V to the cache for later steps
    bf16* cache = const_cast<bf16*>(kv_cache) + (size_t)b * kv_rows * krs + (size_t)h * HD;
    for (int i = tid; i < M * 16; i += 128 * NG) {
      const int row = i >> 4, part = (i >> 3) & 1, chunk = i & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(qbase + (size_t)row * rs + (1 + part) * D + chunk * 8);
      *reinterpret_cast<uint4*>(cache + (size_t)(f0 + row) * krs + part * D + chunk * 8) = v;
    }
  }
  if (grp == 0) load_tile_async(Qs, qbase, rs, 0, M, gtid);
  if (!early && it_first < it_end) issue_tile(it_first, 0);
  cp_async_commit();
#pragma unroll
  for (int st = 1; st < NSTG - 1; ++st) {          // the ring's other leading stages, one cp.async group each
    if (it_first + st * NG < it_end) issue_tile(it_first + st * NG, st);
    cp_async_commit();
  }
  if (NG > 1) {                        // Q comes from group 0's threads: everybody sees it before the first tile
    cp_async_wait<0>();
    __syncthreads();
  }

  uint32_t qf[NG == 1 ? 2 : 1][4][4];
  float o[2][8][4];
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[rb][j][0] = o[rb][j][1] = o[rb][j][2] = o[rb][j][3] = 0.f; }
  float m[4], l[4];                    // rows g, g + 8, 16 + g, 24 + g
#pragma unroll
  for (int r = 0; r < 4; ++r) { m[r] = -INFINITY; l[r] = 0.f; }

  int iter = 0;
  for (int it = it_first; it < it_end; it += NG, ++iter) {
    const int buf = iter % NSTG;
    if (it + (NSTG - 1) * NG < it_end) issue_tile(it + (NSTG - 1) * NG, (iter + NSTG - 1) % NSTG);
    cp_async_commit();                 // (possibly empty: the group count is what cp.async.wait_group counts)
    cp_async_wait<NSTG - 1>();         // tile `it` has landed
    group_sync();
    if (NG == 1 && iter == 0) {
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        const int row = rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ldsm_x4(smem_addr(Qs + sw(row, kk * 2 + (lane >> 4))), qf[rb][kk]);
      }
    }
    const int cnt = min(KT, T2 - it * KT);
    if (warp * 16 < cnt) {             // this warp's 16 keys of the tile: 16 * warp ..
      const bf16* Kt = KVs + (buf * 2 + 0) * TILE_ELEMS;
      const bf16* Vt = KVs + (buf * 2 + 1) * TILE_ELEMS;
      float sc[2][2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) { sc[rb][nb][0] = sc[rb][nb][1] = sc[rb][nb][2] = sc[rb][nb][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t kf[4];
        const int mi = lane >> 3;
        const int row = warp * 16 + (lane & 7) + (mi >> 1) * 8;
        ldsm_x4(smem_addr(Kt + sw(row, kk * 2 + (mi & 1))), kf);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          if (NG > 1) {
            const int qrow = rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            ldsm_x4(smem_addr(Qs + sw(qrow, kk * 2 + (lane >> 4))), qf[0][kk]);
          }
          mma_bf16(sc[rb][0], qf[NG == 1 ? rb : 0][kk], kf[0], kf[1]);
          mma_bf16(sc[rb][1], qf[NG == 1 ? rb : 0][kk], kf[2], kf[3]);
        }
      }
      // ---- keys past the end of the cache are masked; online softmax per row over this warp's keys
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool vis = warp * 16 + nb * 8 + 2 * t4 + (e & 1) < cnt;
            sc[rb][nb][e] = vis ? sc[rb][nb][e] : -INFINITY;
            mx[rb * 2 + (e >> 1)] = fmaxf(mx[rb * 2 + (e >> 1)], sc[rb][nb][e]);
          }
      float al[4], ms[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float mn = fmaxf(m[r], mx[r]);
        ms[r] = mn == -INFINITY ? 0.f : mn * scale_log2;
        al[r] = exp2f(m[r] * scale_log2 - ms[r]);
        m[r] = mn;
      }
      uint32_t af[2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        float p[2][4], sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          p[nb][0] = exp2f(fmaf(sc[rb][nb][0], scale_log2, -ms[rb * 2]));
          p[nb][1] = exp2f(fmaf(sc[rb][nb][1], scale_log2, -ms[rb * 2]));
          p[nb][2] = exp2f(fmaf(sc[rb][nb][2], scale_log2, -ms[rb * 2 + 1]));
          p[nb][3] = exp2f(fmaf(sc[rb][nb][3], scale_log2, -ms[rb * 2 + 1]));
          sum0 += p[nb][0] + p[nb][1];
          sum1 += p[nb][2] + p[nb][3];
        }
        l[rb * 2] = l[rb * 2] * al[rb * 2] + sum0;
        l[rb * 2 + 1] = l[rb * 2 + 1] * al[rb * 2 + 1] + sum1;
        af[rb][0] = pack_bf16x2(p[0][0], p[0][1]); af[rb][1] = pack_bf16x2(p[0][2], p[0][3]);
        af[rb][2] = pack_bf16x2(p[1][0], p[1][1]); af[rb][3] = pack_bf16x2(p[1][2], p[1][3]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[rb][j][0] *= al[rb * 2]; o[rb][j][1] *= al[rb * 2];
          o[rb][j][2] *= al[rb * 2 + 1]; o[rb][j][3] *= al[rb * 2 + 1];
        }
      }
      // ---- O += P V over this warp's 16 keys (one k16 step)
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t vf[4];
        const int mi = lane >> 3;
        const int row = warp * 16 + (lane & 7) + (mi & 1) * 8;
        ldsm_x4_trans(smem_addr(Vt + sw(row, jp * 2 + (mi >> 1))), vf);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          mma_bf16(o[rb][2 * jp], af[rb], vf[0], vf[1]);
          mma_bf16(o[rb][2 * jp + 1], af[rb], vf[2], vf[3]);
        }
      }
    }
    group_sync();     // tile `buf` fully consumed before the next iteration's prefetch overwrites it
  }
  cp_async_wait<0>();
  __syncthreads();    // (an empty key range never entered the loop) nobody reads Q / K / V any more: Op may overwrite them

  // ---- the four warps' states -> shared memory -> one state per row
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  float* myO = Op + (size_t)wall * SQ * OP_LD;
#pragma unroll
  for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float2*>(myO + (rb * 16 + g) * OP_LD + j * 8 + 2 * t4) = make_float2(o[rb][j][0], o[rb][j][1]);
      *reinterpret_cast<float2*>(myO + (rb * 16 + g + 8) * OP_LD + j * 8 + 2 * t4) = make_float2(o[rb][j][2], o[rb][j][3]);
    }
    if (t4 == 0) {
      s_ml[wall][rb * 16 + g][0] = m[rb * 2]; s_ml[wall][rb * 16 + g][1] = l[rb * 2];
      s_ml[wall][rb * 16 + g + 8][0] = m[rb * 2 + 1]; s_ml[wall][rb * 16 + g + 8][1] = l[rb * 2 + 1];
    }
  }
  __syncthreads();
  const int row = gtid >> 2, c0 = (gtid & 3) * 16;    // 32 rows x 4 column quarters (the first 128 threads merge)
  const bool merger = tid < 128;
  float mxr = -INFINITY;
#pragma unroll
  for (int w = 0; w < 4 * NG; ++w) mxr = fmaxf(mxr, s_ml[w][row][0]);
  float acc[16], lsum = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
  for (int w = 0; w < 4 * NG; ++w) {
    const float mw = s_ml[w][row][0];
    if (mw == -INFINITY || !merger) continue;
    const float wgt = exp2f((mw - mxr) * scale_log2);
    lsum = fmaf(s_ml[w][row][1], wgt, lsum);
    const float* src = Op + ((size_t)w * SQ + row) * OP_LD + c0;
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      acc[c] = fmaf(v.x, wgt, acc[c]); acc[c + 1] = fmaf(v.y, wgt, acc[c + 1]);
      acc[c + 2] = fmaf(v.z, wgt, acc[c + 2]); acc[c + 3] = fmaf(v.w, wgt, acc[c + 3]);
    }
  }
  if (n_splits == 1) {
    if (row < M && merger) {
      const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
      bf16* dst = ctx + ((size_t)b * M + row) * D + (size_t)h * HD + c0;
#pragma unroll
      for (int c = 0; c < 16; c += 8) {
        uint4 v;
        v.x = pack_bf16x2(acc[c] * inv, acc[c + 1] * inv); v.y = pack_bf16x2(acc[c + 2] * inv, acc[c + 3] * inv);
        v.z = pack_bf16x2(acc[c + 4] * inv, acc[c + 5] * inv); v.w = pack_bf16x2(acc[c + 6] * inv, acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c) = v;
      }
    }
    return;
  }
  // ---- split keys over CTAs: this CTA's state per row (layout of attn_mma_kernel: 64 rows x 66 floats per split);
  //      the CTA of this (stream, head) that finishes last merges the splits in a fixed order
  const size_t bh = (size_t)b * gridDim.y + h;
  float* P = partials + (bh * n_splits + split) * (size_t)(QT * 66);
  if (merger) {
#pragma unroll
    for (int c = 0; c < 16; c += 2) *reinterpret_cast<float2*>(P + row * 66 + c0 + c) = make_float2(acc[c], acc[c + 1]);
    if ((tid & 3) == 0) *reinterpret_cast<float2*>(P + row * 66 + 64) = make_float2(mxr, lsum);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(counters + bh, 1u) == (unsigned)n_splits - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* P0 = partials + bh * n_splits * (size_t)(QT * 66);
  for (int idx = tid; idx < M * 8; idx += 128 * NG) {
    const int r = idx >> 3, chunk = idx & 7;
    float mx = -INFINITY;
    for (int sp = 0; sp < n_splits; ++sp) mx = fmaxf(mx, __ldcg(P0 + (size_t)sp * QT * 66 + r * 66 + 64));
    float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ls = 0.f;
    for (int sp = 0; sp < n_splits; ++sp) {
      const float* pr = P0 + (size_t)sp * QT * 66 + r * 66;
      const float2 ml = __ldcg(reinterpret_cast<const float2*>(pr + 64));
      if (ml.x == -INFINITY) continue;
      const float w = exp2f((ml.x - mx) * scale_log2);
      ls = fmaf(ml.y, w, ls);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float2 ov = __ldcg(reinterpret_cast<const float2*>(pr + chunk * 8 + e));
        a8[e] = fmaf(ov.x, w, a8[e]);
        a8[e + 1] = fmaf(ov.y, w, a8[e + 1]);
      }
    }
    const float inv = ls > 0.f ? 1.0f / ls : 0.f;
    uint4 v;
    v.x = pack_bf16x2(a8[0] * inv, a8[1] * inv); v.y = pack_bf16x2(a8[2] * inv, a8[3] * inv);
    v.z = pack_bf16x2(a8[4] * inv, a8[5] * inv); v.w = pack_bf16x2(a8[6] * inv, a8[7] * inv);
    *reinterpret_cast<uint4*>(ctx + ((size_t)b * M + r) * D + (size_t)h * HD + chunk * 8) = v;
  }
  if (tid == 0) counters[bh] = 0u;            // ready for the next launch (stream order)
}
}  // namespace

w2vs_status_t launch_attention_mma(const AttnArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(a.D == a.heads * HD, "attention head_dim must be 64");
  W2VS_REQUIRE(a.D % 8 == 0, "attention D alignment");
  const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  if (a.n_step_q > 0) {
    const int nt = (a.n_step_q + QT - 1) / QT;
    // long left contexts: split the keys of a (stream, head) over up to 8 CTAs (two key tiles per split at least)
    const int n_kt = (a.n_step_keys + KT - 1) / KT;
    int splits = 1;
    if (a.step_partials != nullptr && a.step_counters != nullptr) {
      // ... as long as the streams and heads alone do not fill the GPU (measured: 16 streams are slower split)
#ifndef W2VS_ATTN_STEP_FILL
#define W2VS_ATTN_STEP_FILL 296
#endif
      const int fill = W2VS_ATTN_STEP_FILL / (a.B * a.heads * nt);
      splits = n_kt / 2 < fill ? n_kt / 2 : fill;
      splits = splits < 1 ? 1 : (splits > kAttnStepMaxSplits ? kAttnStepMaxSplits : splits);
    }
#ifndef W2VS_ATTN_STEP32
#define W2VS_ATTN_STEP32 1
#endif
    if (W2VS_ATTN_STEP32 && a.n_step_q <= SQ) {     // the usual decision step: keys split over the warps of a CTA
      // two tile groups per CTA while the CTAs do not fill the GPU (one stream through the chain: 1.38 -> 1.35 ms per
      // step); with 256 CTAs (16 streams) one group is as fast (1.78 vs 1.79 ms) and leaves three CTAs per SM
      dim3 grid32((unsigned)splits, (unsigned)a.heads, (unsigned)a.B);
      const bool two = (int64_t)splits * a.heads * a.B <= num_sms();
#ifndef W2VS_ATTN_STEP_STAGES
#define W2VS_ATTN_STEP_STAGES 4
#endif
      constexpr int NSTG1 = W2VS_ATTN_STEP_STAGES;
      constexpr size_t smem1 = (size_t)(1 + 2 * NSTG1) * TILE_ELEMS * 2, smem2 = (size_t)(1 + 8) * TILE_ELEMS * 2;
      static PerDeviceOnce attr_once;
      bool& attr_done = attr_once.here();
      if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(attn_step32_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e == cudaSuccess)
          e = cudaFuncSetAttribute(attn_step32_kernel<1, NSTG1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) { set_error("attn_step32 smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
        attr_done = true;
      }
      if (two)
        launch_pdl(attn_step32_kernel<2, 2>, grid32, dim3(256), smem2, st, (const bf16*)a.qkv, (bf16*)a.ctx, a.n_step_keys,
                   a.n_step_q, a.D, scale_log2, (const bf16*)a.kv_cache, a.kv_rows, splits, a.step_partials,
                   a.step_counters);
      else
        launch_pdl(attn_step32_kernel<1, NSTG1>, grid32, dim3(128), smem1, st, (const bf16*)a.qkv, (bf16*)a.ctx, a.n_step_keys,
                   a.n_step_q, a.D, scale_log2, (const bf16*)a.kv_cache, a.kv_rows, splits, a.step_partials,
                   a.step_counters);
      W2VS_CHECK_LAUNCH("attn_step32_kernel");
      return W2VS_OK;
    }
    dim3 grid((unsigned)(nt * splits), (unsigned)a.heads, (unsigned)a.B);
    launch_pdl(attn_mma_kernel, grid, dim3(128), 0, st, (const bf16*)a.qkv, (const uint8_t*)nullptr, (bf16*)a.ctx,
               a.n_step_keys, a.n_step_q, 1, 0, a.D, nt, nt, scale_log2, (const bf16*)a.kv_cache, a.kv_rows, splits,
               a.step_partials, a.step_counters);
    W2VS_CHECK_LAUNCH("attn_mma_kernel");
    return W2VS_OK;
  }
  const int M = a.T2 + (a.rc > 0 ? (a.T2 / a.main_ctx) * a.rc : 0);
  const int n_main = (a.T2 + QT - 1) / QT, n_rc = (M - a.T2 + QT - 1) / QT;
  dim3 grid((unsigned)(n_main + n_rc), (unsigned)a.heads, (unsigned)a.B);
  attn_mma_kernel<<<grid, 128, 0, st>>>((const bf16*)a.qkv, a.keypad, (bf16*)a.ctx, a.T2, M, a.main_ctx, a.rc,
                                        a.D, n_main, n_main + n_rc, scale_log2, nullptr, 0, 1, nullptr, nullptr);
  W2VS_CHECK_LAUNCH("attn_mma_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
