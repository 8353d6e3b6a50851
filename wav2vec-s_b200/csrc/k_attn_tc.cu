// Block-mask-aware fused attention on the 5th-generation tensor cores (tcgen05 + TMEM), bf16, head_dim 64.
//
// Same contract as k_attn_simt.cu / k_attn_mma.cu (MultiheadAttention fast path + gen_block_attn_mask,
// modules/multihead_attention.py:162-194, wav2vec_S.py:444-489): token buffer qkv [B, M, 3D], M = T' + nb*rc;
// visibility derived from (T', main, rc) and the key-padding bytes; key tiles invisible to a whole query tile
// are never loaded, tiles entirely below the block diagonal skip the per-element mask.
//
// One CTA = 128 query tokens of one (utterance, head); two CTAs are resident per SM so that one CTA's
// softmax overlaps the other's MMAs.  Roles (320 threads):
//   (warp ids: softmax 0-7, loader 8, MMA 9 -- the SMSP arbiter favours high warp ids and the two
//    single-thread roles are on the critical path of every tile)
//   warp 8   TMA loader   Q once; K and V tiles of 128 keys through two 2-stage rings (128B-swizzled smem)
//   warp 9   MMA issuer   S = Q K^T  (tcgen05.mma M=128,N=128,K=64: A,B K-major from smem -> TMEM cols 0..127)
//                         O += P V   (M=128,N=64,K=128: A = P bf16 from TMEM cols 128..191, B = V MN-major from
//                         smem -> TMEM cols 192..255); also owns the TMEM allocation (256 columns)
//   warps 0-7 softmax     two threads per query row (warp w and w+4 share TMEM lane quarter w&3, each owns 64 of
//                         the 128 key columns): tcgen05.ld S, mask, running max (halves exchanged through smem;
//                         lazy rescale of O: only when the max grows by more than 2^8), exp2 (ex2.approx), row
//                         sum, P -> TMEM (tcgen05.st); final O / l -> bf16 -> global.  16 softmax warps per SM
//                         keep the MUFU / FMA pipes busy (4 warps per SMSP issued only ~0.25 IPC each).
// S(i+1) is issued as soon as the softmax threads have read S(i), i.e. it overlaps softmax(i)'s exponentials
// and PV(i).  TMEM budget: S 128 + P 64 + O 64 = 256 columns per CTA.
#include <math.h>
#include <limits.h>
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
__device__ int g_attn_tc_fault = 0;
#ifdef W2VS_ATTN_TRACE
__device__ long long g_attn_trace[2][64][8];   // [role: 0 softmax warp 0, 1 MMA thread][tile][event] (clock64)
#define TRACE(role, it, ev) do { if (trace_on && (it) < 64) g_attn_trace[role][it][ev] = clock64(); } while (0)
#else
#define TRACE(role, it, ev) do { } while (0)
#endif
}
#define W2VS_TC_FAULT_FLAG (&::w2vs::g_attn_tc_fault)
#include "tc_common.cuh"

namespace w2vs {
namespace {
using namespace tc;

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;     // 16 KB: Q, K or V tile
constexpr int NS = 2;                        // K ring and V ring depth
constexpr int N_SOFTMAX_WARPS = 8, LOADER_WARP = 8, MMA_WARP = 9, N_THREADS = 320;
constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * NS) + 2 * KT * 4 /*info*/ + 2 * 2 * QT * 4 /*exchange*/ + 256 + 1024;
constexpr uint32_t TMEM_COLS = 256, S_COL = 0, P_COL = 128, O_COL = 192;
constexpr float RESCALE_THRESHOLD = 8.0f;    // log2 units

__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_1sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// MN-major operand tile with 128-byte swizzle: rows = K index (128 B each), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset: next 64-element MN group (single group here)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: next group of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// named barrier over the 256 softmax threads with an OR reduction of a predicate
__device__ __forceinline__ bool softmax_bar_or(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\t"
      "barrier.cta.red.or.pred.aligned p, 1, 256, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out) : "r"((uint32_t)pred) : "memory");
  return out != 0;
}

// barrier between the two warps that share a TMEM lane quarter (ids 2..5, 64 threads)
__device__ __forceinline__ void pair_bar(int quarter) {
  asm volatile("barrier.cta.sync.aligned %0, 64;" ::"r"(2 + quarter) : "memory");
}
// packed fp32 pairs (sm_100): one instruction for two lanes of the softmax arithmetic
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct TileSeq {   // key tiles visible to one query tile
  int seg0_end, seg1_begin, seg1_end, n0, n_kt;
  __device__ __forceinline__ void get(int it, int& k0, int& cnt, bool& s1) const {
    s1 = it >= n0;
    k0 = s1 ? seg1_begin + (it - n0) * KT : it * KT;
    cnt = min(KT, (s1 ? seg1_end : seg0_end) - k0);
  }
};

__global__ void __launch_bounds__(N_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const uint8_t* __restrict__ keypad,
               bf16* __restrict__ ctx, int T2, int M, int main_ctx, int rc, int D, int n_main_tiles, int n_tiles,
               float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + TILE_BYTES;
  const uint32_t sV = sK + NS * TILE_BYTES;
  const uint32_t sInfo = sV + NS * TILE_BYTES;              // int [2][KT]
  const uint32_t sXch = sInfo + 2 * KT * 4;                 // float [2 tile parities][2 halves][QT]
  const uint32_t bars = sXch + 2 * 2 * QT * 4;
  const uint32_t bar_q = bars, bar_kfull = bars + 8, bar_kempty = bar_kfull + 8 * NS, bar_vfull = bar_kempty + 8 * NS,
                 bar_vempty = bar_vfull + 8 * NS, bar_sfull = bar_vempty + 8 * NS, bar_sfree = bar_sfull + 8,
                 bar_pfull = bar_sfree + 8, bar_pvdone = bar_pfull + 8;
  const uint32_t tmem_slot = bar_pvdone + 8;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  int* s_info = reinterpret_cast<int*>(gen_base + (sInfo - smem_base));
  float* s_xch = reinterpret_cast<float*>(gen_base + (sXch - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int tile_id = n_tiles - 1 - (int)blockIdx.x;        // heaviest query tiles first
#ifdef W2VS_ATTN_TRACE
  const bool trace_on = blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == 1 && lane == 0 && (warp == 0 || warp == MMA_WARP);
#endif
  const int nb = T2 / main_ctx;
  const int rcd = rc > 0 ? rc : 1;

  int q_first, q_count;
  if (tile_id < n_main_tiles) { q_first = tile_id * QT; q_count = min(QT, T2 - q_first); }
  else { q_first = T2 + (tile_id - n_main_tiles) * QT; q_count = min(QT, M - q_first); }
  auto qblock = [&](int m) { return m < T2 ? m / main_ctx : (m - T2) / rcd; };
  const int qb_lo = qblock(q_first), qb_hi = qblock(q_first + q_count - 1);
  TileSeq ts;
  ts.seg0_end = min(main_ctx * (qb_hi + 1), T2);
  ts.seg1_begin = ts.seg1_end = 0;
  if (rc > 0 && qb_lo <= nb - 1) { ts.seg1_begin = T2 + rc * qb_lo; ts.seg1_end = T2 + rc * (min(qb_hi, nb - 1) + 1); }
  ts.n0 = (ts.seg0_end + KT - 1) / KT;
  ts.n_kt = ts.n0 + (ts.seg1_end - ts.seg1_begin + KT - 1) / KT;
  const int n_kt = ts.n_kt;
  const int row_base = b * M;                                // first token row of this utterance in qkv

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_kfull + 8 * s, 1); mbar_init(bar_kempty + 8 * s, 1);
      mbar_init(bar_vfull + 8 * s, 1); mbar_init(bar_vempty + 8 * s, 1);
    }
    mbar_init(bar_sfull, 1); mbar_init(bar_sfree, N_SOFTMAX_WARPS); mbar_init(bar_pfull, N_SOFTMAX_WARPS); mbar_init(bar_pvdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == LOADER_WARP) {
    // ===================== TMA loader =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQKV) : "memory");
      mbar_expect_tx(bar_q, TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, bar_q, h * HD, row_base + q_first);
      bool ok = true;
      for (int it = 0; it < n_kt && ok; ++it) {
        const int s = it % NS;
        const uint32_t par = ((it / NS) & 1) ^ 1;
        int k0, cnt; bool s1;
        ts.get(it, k0, cnt, s1);
        if (!(ok = mbar_wait(bar_kempty + 8 * s, par))) break;
        mbar_expect_tx(bar_kfull + 8 * s, TILE_BYTES);
        tma_load_2d(sK + s * TILE_BYTES, &tmQKV, bar_kfull + 8 * s, D + h * HD, row_base + k0);
        if (!(ok = mbar_wait(bar_vempty + 8 * s, par))) break;
        mbar_expect_tx(bar_vfull + 8 * s, TILE_BYTES);
        tma_load_2d(sV + s * TILE_BYTES, &tmQKV, bar_vfull + 8 * s, 2 * D + h * HD, row_base + k0);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // D=f32, A=B=bf16; QK: both K-major, N=128; PV: B MN-major (bit 16), N=64; M=128
      constexpr uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      constexpr uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                    ((uint32_t)(QT >> 4) << 24);
      bool ok = mbar_wait(bar_q, 0);
      auto issue_s = [&](int it) {
        const int s = it % NS;
        if (!mbar_wait(bar_kfull + 8 * s, (it / NS) & 1)) return false;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss(tmem_base + S_COL, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + s * TILE_BYTES + k * 32),
                  idesc_qk, k > 0 ? 1u : 0u);
        tc_commit_1sm(bar_sfull);
        tc_commit_1sm(bar_kempty + 8 * s);
        return true;
      };
      if (ok && n_kt > 0) ok = issue_s(0);
      for (int it = 0; it < n_kt && ok; ++it) {
        TRACE(1, it, 0);
        if (it + 1 < n_kt) {
          if (!(ok = mbar_wait(bar_sfree, it & 1))) break;      // softmax has read S(it)
          tc_fence_after();
          TRACE(1, it, 1);
          if (!(ok = issue_s(it + 1))) break;
          TRACE(1, it, 2);
        }
        const int s = it % NS;
        if (!(ok = mbar_wait(bar_vfull + 8 * s, (it / NS) & 1))) break;
        TRACE(1, it, 3);
        if (!(ok = mbar_wait(bar_pfull, it & 1))) break;         // P(it) is in TMEM, O has been rescaled
        tc_fence_after();
        TRACE(1, it, 4);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_ts(tmem_base + O_COL, tmem_base + P_COL + k * 8, umma_desc_mn_sw128(sV + s * TILE_BYTES + k * 2048),
                  idesc_pv, (it > 0 || k > 0) ? 1u : 0u);
        tc_commit_1sm(bar_pvdone);
        tc_commit_1sm(bar_vempty + 8 * s);
        TRACE(1, it, 5);
      }
    }
  } else {
    // ===================== softmax warps: two threads per query row =====================
    // Visibility of the 128 key columns of a tile for one query row is a contiguous column range [lo, hi):
    //   main keys: columns whose block <= qblock(row)  ->  [0, (qb+1)*main - k0)
    //   look-ahead keys: the rc copies owned by qblock(row)  ->  [T2 + qb*rc - k0, +rc)
    // so 32-column chunks are classified per warp as all-visible (no masking), none-visible (no loads, no
    // exponentials: P = 0) or partial (one range compare per element).  Tiles that contain padded keys take
    // the general per-key path (s_info), which also covers arbitrary (non-prefix) padding masks.
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;                 // row inside the query tile == TMEM lane
    const int st = warp * 32 + lane;                     // 0..255: threads 0..127 describe the key columns
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // rows past the end of the tile behave like the last valid row (their output is never stored); this keeps
    // the chunk classification uniform across the warp
    const int my_qb = row < q_count ? qblock(q_first + row) : qb_hi;
    const uint8_t* kp = keypad + (size_t)b * M;
    float m_ref = -INFINITY;
    uint64_t l2 = pack2(0.f, 0.f);                       // partial row sums (two accumulators)
    bool ok = true;

    int k0, cnt; bool s1;
    uint8_t kp_next = 0;
    if (n_kt > 0 && st < KT) { ts.get(0, k0, cnt, s1); kp_next = st < cnt ? kp[k0 + st] : 0; }

    for (int it = 0; it < n_kt && ok; ++it) {
      ts.get(it, k0, cnt, s1);
      // ---- per-key description for the general path; the barrier publishes it and ORs "tile has padding"
      const bool padded = st < cnt && kp_next != 0;
      int* info_t = s_info + (it & 1) * KT;
      if (st < KT) {
        int info;
        if (st < cnt && !padded) info = s1 ? (k0 + st - T2) / rcd : (k0 + st) / main_ctx;
        else info = s1 ? -2 : INT_MAX;
        info_t[st] = info;
      }
      TRACE(0, it, 0);
      const bool has_pad = softmax_bar_or(padded);
      TRACE(0, it, 1);
      if (it + 1 < n_kt && st < KT) {   // prefetch the padding byte of the next tile's column
        int k0n, cntn; bool s1n;
        ts.get(it + 1, k0n, cntn, s1n);
        kp_next = st < cntn ? kp[k0n + st] : 0;
      }
      // ---- visible column range of this row
      int lo = 0, hi = 0;
      if (!s1) { hi = min(max((my_qb + 1) * main_ctx - k0, 0), cnt); }
      else if (my_qb <= nb - 1) { lo = min(max(T2 + my_qb * rc - k0, 0), cnt); hi = min(max(T2 + (my_qb + 1) * rc - k0, 0), cnt); }
      const uint32_t span = (uint32_t)(hi - lo);
      // classes of this thread's two 32-column chunks, warp-uniform
      bool all_vis[2], none_vis[2];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = (2 * half + cc) * 32;
        const bool a = !has_pad && lo <= c0 && c0 + 32 <= hi;
        const bool n = (!has_pad && (hi <= c0 || lo >= c0 + 32 || span == 0)) || c0 >= cnt;
        all_vis[cc] = __all_sync(0xffffffffu, a);
        none_vis[cc] = __all_sync(0xffffffffu, n);
      }
      ok = mbar_wait(bar_sfull, it & 1);
      tc_fence_after();
      TRACE(0, it, 2);

      // ---- sweep 1: maximum of the visible scores of this half row
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (none_vis[cc]) continue;
        const int c0 = (2 * half + cc) * 32;
        uint32_t r[32];
        tmem_ld32(tlane + S_COL + c0, r);
        tmem_ld_wait();
        if (all_vis[cc]) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            mxa = fmaxf(mxa, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
            mxb = fmaxf(mxb, fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
          }
        } else if (!has_pad) {
          const int off = c0 - lo;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            mxa = fmaxf(mxa, (uint32_t)(off + j) < span ? __uint_as_float(r[j]) : -INFINITY);
            mxb = fmaxf(mxb, (uint32_t)(off + j + 1) < span ? __uint_as_float(r[j + 1]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ki = info_t[c0 + j];
            const bool vis = s1 ? (ki == my_qb) : (ki <= my_qb);
            mxa = fmaxf(mxa, vis ? __uint_as_float(r[j]) : -INFINITY);
          }
        }
      }
      // ---- combine the two half-row maxima through shared memory
      float* xch = s_xch + (it & 1) * 2 * QT;
      xch[half * QT + row] = fmaxf(mxa, mxb);
      pair_bar(quarter);
      const float mx = fmaxf(fmaxf(mxa, mxb), xch[(half ^ 1) * QT + row]);
      TRACE(0, it, 3);
      // ---- running maximum with lazy rescale (exact: the final normalisation uses the same reference)
      const float m_tile = mx * scale_log2;               // -inf stays -inf
      const bool grow = m_tile > m_ref + RESCALE_THRESHOLD;
      float alpha = 1.0f;
      if (grow) {
        alpha = ex2_approx(m_ref - m_tile);
        m_ref = m_tile;
        l2 = ffma2(l2, pack2(alpha, alpha), pack2(0.f, 0.f));
      }
      if (it > 0) {
        ok = mbar_wait(bar_pvdone, (it - 1) & 1) && ok;   // PV(it-1) retired: P is free, O is stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {               // this thread rescales its 32 columns of O
          uint32_t r[32];
          tmem_ld32(tlane + O_COL + half * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
          tmem_st32(tlane + O_COL + half * 32, r);
        }
      }
      TRACE(0, it, 4);
      const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
      const uint64_t sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_use, -m_use);

      // ---- sweep 2: P = exp2(s * scale - m), row sum, bf16 P -> TMEM
      const int last_read = none_vis[1] ? (none_vis[0] ? -1 : 0) : 1;
      if (last_read < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(bar_sfree);
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * half + cc, c0 = c * 32;
        uint32_t pk[16];
        if (none_vis[cc]) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
          tmem_st16(tlane + P_COL + c * 16, pk);
          continue;
        }
        uint32_t r[32];
        tmem_ld32(tlane + S_COL + c0, r);
        tmem_ld_wait();
        if (cc == last_read) {        // every score of this tile has been read by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_local(bar_sfree);
        }
        if (!all_vis[cc]) {           // partial chunk: replace invisible scores by -inf
          if (!has_pad) {
            const int off = c0 - lo;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              r[j] = (uint32_t)(off + j) < span ? r[j] : 0xff800000u;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int ki = info_t[c0 + j];
              const bool vis = s1 ? (ki == my_qb) : (ki <= my_qb);
              r[j] = vis ? r[j] : 0xff800000u;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float a0, a1;
          unpack2(ffma2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sc2, nm2), a0, a1);
          const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
          l2 = fadd2(l2, pack2(p0, p1));
          pk[j >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tlane + P_COL + c * 16, pk);
      }
      TRACE(0, it, 5);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(bar_pfull);
      TRACE(0, it, 6);
    }

    // ---- epilogue: O / l -> bf16 -> ctx (each thread: 32 of the 64 output columns of its row)
    float la, lb;
    unpack2(l2, la, lb);
    float* xch = s_xch + (n_kt & 1) * 2 * QT;
    xch[half * QT + row] = la + lb;
    pair_bar(quarter);
    const float l = (la + lb) + xch[(half ^ 1) * QT + row];
    if (n_kt > 0) ok = mbar_wait(bar_pvdone, (n_kt - 1) & 1) && ok;
    tc_fence_after();
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    bf16* dst = ctx + ((size_t)row_base + q_first + row) * D + (size_t)h * HD + half * 32;
    {
      uint32_t r[32];
      tmem_ld32(tlane + O_COL + half * 32, r);
      tmem_ld_wait();
      if (row < q_count && n_kt > 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}
}  // namespace

w2vs_status_t launch_attention_tc(const AttnArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(a.dtype == W2VS_BF16, "tcgen05 attention takes bf16");
  W2VS_REQUIRE(a.D == a.heads * HD, "attention head_dim must be 64");
  W2VS_REQUIRE(a.n_step_q == 0, "tcgen05 attention implements the full-utterance (block mask) mode");
  W2VS_REQUIRE(((uintptr_t)a.qkv & 15) == 0 && ((uintptr_t)a.ctx & 15) == 0, "attention buffers must be 16-byte aligned");
  const int M = a.T2 + (a.rc > 0 ? (a.T2 / a.main_ctx) * a.rc : 0);
  const int n_main = (a.T2 + QT - 1) / QT, n_rc = (M - a.T2 + QT - 1) / QT;
  alignas(64) CUtensorMap tm;
  W2VS_TRY(tc::make_map(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.qkv, (uint64_t)3 * a.D, (uint64_t)a.B * M,
                        (uint64_t)3 * a.D, HD, 128, CU_TENSOR_MAP_SWIZZLE_128B));
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_error("attn_tc smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    attr_done = true;
  }
  dim3 grid((unsigned)(n_main + n_rc), (unsigned)a.heads, (unsigned)a.B);
  const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  attn_tc_kernel<<<grid, N_THREADS, SMEM_BYTES, st>>>(tm, a.keypad, (bf16*)a.ctx, a.T2, M, a.main_ctx, a.rc, a.D, n_main,
                                                n_main + n_rc, scale_log2);
  W2VS_CHECK_LAUNCH("attn_tc_kernel");
  return W2VS_OK;
}

}  // namespace w2vs
