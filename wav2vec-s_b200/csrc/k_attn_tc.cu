// Block-mask-aware fused attention on the 5th-generation tensor cores (tcgen05 + TMEM), bf16, head_dim 64.
//
// Same contract as k_attn_simt.cu / k_attn_mma.cu (MultiheadAttention fast path + gen_block_attn_mask,
// modules/multihead_attention.py:162-194, wav2vec_S.py:444-489): token buffer qkv [B, M, 3D], M = T' + nb*rc;
// visibility derived from (T', main, rc) and the key-padding bytes; key tiles invisible to a whole query tile
// are never loaded, tiles entirely below the block diagonal skip the per-element mask.
//
// Three persistent CTAs per SM, each an independent pipeline over its own list of work items (one item = 128
// query tokens of one (utterance, head)); key tiles are 64 tokens.  The kernel is latency bound per warp (TMEM
// load -> row max -> exponentials -> P store is one dependent chain per tile) and throughput bound by the MUFU /
// FMA pipes (16 exp2 / clk / SM), so what matters is how many independent softmax warps each SM sub-partition has
// to interleave: three here (a 128-column score tile per thread needs 128+ live registers and 256 TMEM columns
// per pipeline, which caps an SM at two pipelines; measured 2150 cycles per 128x128 tile against a pipe bound of
// about 900).  Per CTA: TMEM 128 columns (S 64 + O 64), smem 64 KB (Q 16, K 2x8, V 2x8, P 16), 160 threads.
// Roles (warps 0-3 softmax, warp 4 MMA issuer):
//   4 warps    softmax     one query row per thread (TMEM lane = row): the 64 scores of a key tile are read from
//                          TMEM exactly once into registers (tcgen05.ld) and S is handed back to the MMA thread at
//                          once, so S(g+1) is computed while this tile's exponentials run; range mask, row max
//                          (thread-local, no cross-thread traffic), lazy rescale of O (only when the max grows by
//                          more than 2^8), exp2 (ex2.approx, part of it emulated on the FMA pipe), row sum, bf16 P
//                          -> smem in the K-major 128B-swizzled operand layout; per item: O / l -> bf16 -> staged
//                          in the P buffer -> TMA store.
//   1 warp     MMA issuer  (one thread; it also issues the TMA loads -- no loader warp, no "empty" barriers: it
//                          commits the MMAs that read a ring slot, so it knows when the slot is free.  The refills
//                          are deferred by one step -- K's slot of S(g+1) after PV(g) has been issued, V's slot of
//                          PV(g-1) at the start of tile g, the next Q after the item's last S -- so the thread never
//                          waits for an MMA it has just issued).  The role is entered through elect_one(): behind
//                          `lane == 0` ptxas wrapped every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY
//                          loop and this thread needed 500 cycles to issue four MMAs -- its loop, 2500 cycles per
//                          tile, was the kernel's period (clock64 trace); now 160 cycles per four MMAs.
//                          S = Q K^T  (tcgen05.mma M=128,N=64,K=64: A,B K-major from smem -> TMEM cols 0..63)
//                          O += P V   (M=128,N=64,K=64: A = bf16 P K-major from smem, B = V MN-major from smem
//                          -> TMEM cols 64..127); owns the TMEM allocation.
#include <math.h>
#include <limits.h>
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
__device__ int g_attn_tc_fault = 0;
#ifdef W2VS_ATTN_TRACE
__device__ long long g_attn_trace[2][64][16];   // [role: 0 softmax warp 0, 1 MMA thread][tile][event] (clock64)
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

constexpr int QT = 128, KT = 64, HD = 64;
constexpr int Q_BYTES = QT * HD * 2;         // 16 KB
constexpr int KV_BYTES = KT * HD * 2;        // 8 KB: K or V tile
constexpr int P_BYTES = QT * KT * 2;         // 16 KB: bf16 P (A operand of PV), reused as the output staging tile
// Ring depth and CTAs per SM as macros for the A/B harness (tools/build_attn_variants.sh).  Measured per layer call
// at cfg3 (round 2, same box): 2 slots x 3 CTAs 466 us; 1 slot x 3 CTAs 500 us; 1 slot x 4 CTAs (96 registers,
// 572 bytes of spills in the tile loop) 1066 us.  F2FP packing is not on the MUFU pipe (tools/mufu_bw.cu: ex2 alone
// 16 / clk / SM, with one bf16x2 pack per two ex2 still 16), and cutting P to bf16 with PRMT instead changes nothing.
// Skipping the row maximum on every tile but an item's first (the reference only has to keep 2^(s - m_ref) inside the
// exponent range; a tile whose row sum overflows would be redone): 475 -> 459 us without the redo path, 552 us with it
// (spills); the tile period of a softmax warp stays at ~2350 cycles either way -- the exponential phase stretches from
// 1050 to 1210 cycles, i.e. the three warps of a sub-partition queue on the MUFU pipe in step.  Not kept.
// Two passes over the scores in TMEM, 32 columns at a time (row maximum, then exponentials: 32 score registers, 96
// registers per thread without spills, bit-identical output): 495 us at three CTAs per SM, 488 us at four -- four
// CTAs only fit with one-slot K / V rings (50 KB each), and then PV(g-1) waits for its V tile: the softmax warps sit
// ~1100 cycles per tile in the "PV retired" wait and the tile period per CTA grows from 2350 to 3400 cycles.  Not kept.
// The chain of ONE softmax warp, uncontended (one CTA per SM, W2VS_ATTN_CTAS=1): 1860 cycles per tile, of which the
// exponential phase is 805 (64 MUFU.EX2 at 8 cycles each plus the dependent FFMA2 / FADD2 / F2FP), TMEM load 70, row
// maximum 300-350, waits and fences ~500.  Per layer call: 929 us at one CTA per SM, 552 at two, 466 at three (tile
// periods 1860 / 2140 / 2350 cycles): the kernel is latency bound per warp and every further resident pipeline still
// adds throughput; starting the co-resident pipelines 600-800 cycles apart changes nothing (466.4 vs 466.5 us).
#ifndef W2VS_ATTN_NS
#define W2VS_ATTN_NS 2
#endif
#ifndef W2VS_ATTN_CTAS
#define W2VS_ATTN_CTAS 3
#endif
constexpr int NS = W2VS_ATTN_NS;             // K ring and V ring depth
constexpr int N_SOFTMAX_WARPS = 4;
constexpr int N_THREADS = 160;               // warps 0-3 softmax, warp 4 MMA
constexpr int CTAS_PER_SM = W2VS_ATTN_CTAS;
constexpr int SMEM_BYTES = Q_BYTES + 2 * NS * KV_BYTES + P_BYTES + 256 /*barriers*/ + 256 /*item ring*/ + 1024 /*align*/;
constexpr uint32_t TMEM_COLS = 128, S_COL = 0, O_COL = 64;
// barrier waits per role: parked (suspend-time hint) or spinning, see tc_common.cuh
#ifndef W2VS_ATTN_MMA_PARK
#define W2VS_ATTN_MMA_PARK 0
#endif
#ifndef W2VS_ATTN_SOFTMAX_PARK
#define W2VS_ATTN_SOFTMAX_PARK 0
#endif
#if W2VS_ATTN_MMA_PARK
#define MMA_WAIT mbar_wait_park
#else
#define MMA_WAIT mbar_wait
#endif
#if W2VS_ATTN_SOFTMAX_PARK
#define SOFTMAX_WAIT mbar_wait_park
#else
#define SOFTMAX_WAIT mbar_wait
#endif
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
// MN-major operand tile with 128-byte swizzle: rows = K index (128 B each), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset: next 64-element MN group (single group here)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: next group of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// exp2 on the FMA pipe for a fraction of the scores (the MUFU pipe, 16 ex2/clk/SM, is the kernel's bound):
// round-to-nearest split x = n + f, |f| <= 0.5, 2^f by a degree-3 minimax polynomial (relative error 7.5e-5,
// far below the bf16 rounding of P), 2^n applied through the exponent bits.  Inputs are clamped at -125.
__device__ __forceinline__ void ex2_emul2(float x0, float x1, float& p0, float& p1) {
  const uint64_t x = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t magic = pack2(12582912.f, 12582912.f);            // 1.5 * 2^23: low mantissa bits = round(x)
  const uint64_t t = fadd2(x, magic);
  const uint64_t n = fadd2(t, pack2(-12582912.f, -12582912.f));
  const uint64_t f = fadd2(x, fmul2(n, pack2(-1.f, -1.f)));
  uint64_t p = ffma2(pack2(0.055171407759189606f, 0.055171407759189606f), f, pack2(0.24261075258255005f, 0.24261075258255005f));
  p = ffma2(p, f, pack2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, pack2(0.9999281167984009f, 0.9999281167984009f));
  float t0, t1, q0, q1;
  unpack2(t, t0, t1);
  unpack2(p, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
__device__ __forceinline__ void tma_store_2d_(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(x), "r"(y) : "memory");
}

struct TileSeq {   // key tiles visible to one query tile
  int seg0_end, seg1_begin, seg1_end, n0, n_kt;
  __device__ __forceinline__ void get(int it, int& k0, int& cnt, bool& s1) const {
    s1 = it >= n0;
    k0 = s1 ? seg1_begin + (it - n0) * KT : it * KT;
    cnt = min(KT, (s1 ? seg1_end : seg0_end) - k0);
  }
};

// Runtime divisors get host-computed magic multipliers (floor(x/d) == umulhi(x, ceil(2^32/d)) for x < 2^32/d;
// the launcher checks the ranges): a single thread walks the work list, and a hardware-less 32-bit division
// costs ~50 dependent instructions -- with half a dozen of them per work item the item switch used to take
// longer than two key tiles.
struct Shape {
  int T2, M, main_ctx, rc, rcd, nb, D, H, B, n_main_tiles, n_tiles;
  int HB, n_vcta;                       // H*B; pipelines in the grid (work-list stride)
  int flag_fold;                        // log2 of the 128-token padding-flag blocks folded into one mask bit
  int G, S, g_div, g_mod;               // (b,h) pairs per block, items per block (G * n_tiles), stride / S, stride % S
  uint32_t magic_main, magic_rcd, magic_H, magic_G;
};
__device__ __forceinline__ int fdiv(int x, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)x, magic) : x; }  // magic 0: d == 1

// One work item = one query tile of one (utterance, head), dealt round-robin to the pipelines: pipeline c takes the
// items w = c, c + n_vcta, ...  The list is ordered in blocks of G (b,h) pairs; inside a block by tile rank (heaviest
// query tile first), then by pair: w = blk * S + rank * G + j.  All query tiles of a (b,h) pair are therefore in
// flight at about the same time on neighbouring CTAs and its K/V (383 KB at cfg3) is fetched from HBM once and then
// hit in L2.  (Ordered by rank across ALL pairs, as this kernel first was, the live K/V set was the whole 392 MB
// token buffer: ncu showed 2.26 GB of DRAM reads per call for 0.59 GB of input and a 20 % L2 hit rate.)  The
// stride n_vcta is not a multiple of S, so a pipeline drifts through the ranks and the load stays balanced.
// Walker keeps (blk, within) = (w / S, w % S) incrementally.
struct Walker {
  int w, blk, within;
  __device__ __forceinline__ void init(const Shape& sh, int vcta) {
    w = vcta;
    blk = 0; within = w;
    while (within >= sh.S) { within -= sh.S; ++blk; }
  }
  __device__ __forceinline__ void next(const Shape& sh) {
    w += sh.n_vcta;
    blk += sh.g_div; within += sh.g_mod;
    if (within >= sh.S) { within -= sh.S; ++blk; }
  }
  __device__ __forceinline__ int rank(const Shape& sh) const { return fdiv(within, sh.magic_G); }
  __device__ __forceinline__ int hb(const Shape& sh) const { return blk * sh.G + within - rank(sh) * sh.G; }
};
struct Item {
  int q_first, q_count, qb_lo, qb_hi, h, row_base, b;
  TileSeq ts;
};
__device__ __forceinline__ int qblock_of(const Shape& sh, int m) {
  return m < sh.T2 ? fdiv(m, sh.magic_main) : fdiv(m - sh.T2, sh.magic_rcd);
}
__device__ __forceinline__ Item item_of(const Shape& sh, const Walker& wk) {
  Item it;
  const int tile_id = sh.n_tiles - 1 - wk.rank(sh);
  const int hb = wk.hb(sh);
  const int b = fdiv(hb, sh.magic_H);
  it.h = hb - b * sh.H;
  it.b = b;
  it.row_base = b * sh.M;
  if (tile_id < sh.n_main_tiles) { it.q_first = tile_id * QT; it.q_count = min(QT, sh.T2 - it.q_first); }
  else { it.q_first = sh.T2 + (tile_id - sh.n_main_tiles) * QT; it.q_count = min(QT, sh.M - it.q_first); }
  it.qb_lo = qblock_of(sh, it.q_first);
  it.qb_hi = qblock_of(sh, it.q_first + it.q_count - 1);
  it.ts.seg0_end = min(sh.main_ctx * (it.qb_hi + 1), sh.T2);
  it.ts.seg1_begin = it.ts.seg1_end = 0;
  if (sh.rc > 0 && it.qb_lo <= sh.nb - 1) {
    it.ts.seg1_begin = sh.T2 + sh.rc * it.qb_lo;
    it.ts.seg1_end = sh.T2 + sh.rc * (min(it.qb_hi, sh.nb - 1) + 1);
  }
  it.ts.n0 = (it.ts.seg0_end + KT - 1) / KT;
  it.ts.n_kt = it.ts.n0 + (it.ts.seg1_end - it.ts.seg1_begin + KT - 1) / KT;
  return it;
}

__global__ void __launch_bounds__(N_THREADS, CTAS_PER_SM)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
               const __grid_constant__ CUtensorMap tmCtx, const uint8_t* __restrict__ keypad,
               const uint8_t* __restrict__ pad_blk, bf16* __restrict__ ctx, Shape sh, int n_items, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_softmax = warp < N_SOFTMAX_WARPS;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + Q_BYTES;
  const uint32_t sV = sK + NS * KV_BYTES;
  const uint32_t sP = sV + NS * KV_BYTES;                   // [128 rows][128 B], 128B-swizzled
  const uint32_t bars = sP + P_BYTES;
  const uint32_t bar_qfull = bars, bar_kfull = bars + 8, bar_vfull = bar_kfull + 8 * NS, bar_sfull = bar_vfull + 8 * NS,
                 bar_sfree = bar_sfull + 8, bar_pfull = bar_sfree + 8, bar_pvdone = bar_pfull + 8;
  const uint32_t tmem_slot = bars + 96;
  const uint32_t item_ring = bars + 256;                    // 8 x 32 bytes, used by the MMA thread only
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  const int D = sh.D;
  const int vcta = blockIdx.x;
#ifdef W2VS_ATTN_TRACE
  const bool trace_on = blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == N_SOFTMAX_WARPS);
#endif

  if (threadIdx.x == N_SOFTMAX_WARPS * 32) {
    mbar_init(bar_qfull, 1);
    for (int s = 0; s < NS; ++s) { mbar_init(bar_kfull + 8 * s, 1); mbar_init(bar_vfull + 8 * s, 1); }
    mbar_init(bar_sfull, 1); mbar_init(bar_sfree, N_SOFTMAX_WARPS); mbar_init(bar_pfull, N_SOFTMAX_WARPS); mbar_init(bar_pvdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == N_SOFTMAX_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (!is_softmax) {
    // ===================== MMA issuer + TMA loader (one thread) =====================
    // The thread that issues the MMAs also knows when each ring slot becomes free (it commits the MMAs that
    // read them), so it issues the TMA loads in its idle time: after S(g+1) is issued it waits for that MMA
    // (it would be waiting for P(g) anyway) and loads K(g+1+NS); after PV(g) it loads V(g+NS).
    if (elect_one()) {
      // D=f32, A=B=bf16; QK: both K-major, N=64; PV: A K-major (P), B MN-major (bit 16), N=64; M=128
      constexpr uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      constexpr uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                    ((uint32_t)(QT >> 4) << 24);
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmKV) : "memory");
      const uint64_t dQ = umma_desc_sw128(sQ), dK = umma_desc_sw128(sK), dV = umma_desc_mn_sw128(sV), dP = umma_desc_sw128(sP);
      bool ok = true;
      // ---- load cursor: (work item, tile in item, global tile index) of the next K tile and of the next V tile
      // One item_of() per work item: the K cursor runs ahead of everything else this thread does, so it computes
      // the item's facts when it enters it and leaves them in a small shared-memory ring, keyed by the item's
      // ordinal; the V cursor, the Q loads and the main loop look them up (this thread used to evaluate item_of
      // four times per item, ~500 cycles each on a single lane).
      struct ItemLite { int n_kt, n0, seg1, hcol, row, q_first, pad0, pad1; };
      ItemLite* ring = reinterpret_cast<ItemLite*>(gen_base + (item_ring - smem_base));
      struct Cur { Walker wk; int it, g, ord; ItemLite f; };
      auto enter_k = [&](Cur& c) {                      // K cursor only: compute and publish
        if (c.wk.w < n_items) {
          const Item ci = item_of(sh, c.wk);
          c.f.n_kt = ci.ts.n_kt; c.f.n0 = ci.ts.n0; c.f.seg1 = ci.ts.seg1_begin; c.f.hcol = ci.h * HD;
          c.f.row = ci.row_base; c.f.q_first = ci.q_first; c.f.pad0 = c.f.pad1 = 0;
          ring[c.ord & 7] = c.f;
        }
      };
      auto enter_v = [&](Cur& c) { if (c.wk.w < n_items) c.f = ring[c.ord & 7]; };
      auto load_k = [&](Cur& c) {                       // tile c.g -> ring slot c.g % NS
        if (c.wk.w >= n_items) return;
        const int k0 = c.it < c.f.n0 ? c.it * KT : c.f.seg1 + (c.it - c.f.n0) * KT;
        const int sl = c.g % NS;
        mbar_expect_tx(bar_kfull + 8 * sl, KV_BYTES);
        tma_load_2d(sK + sl * KV_BYTES, &tmKV, bar_kfull + 8 * sl, D + c.f.hcol, c.f.row + k0);
        ++c.g;
        if (++c.it == c.f.n_kt) { c.it = 0; c.wk.next(sh); ++c.ord; enter_k(c); }
      };
      auto load_v = [&](Cur& c) {
        if (c.wk.w >= n_items) return;
        const int k0 = c.it < c.f.n0 ? c.it * KT : c.f.seg1 + (c.it - c.f.n0) * KT;
        const int sl = c.g % NS;
        mbar_expect_tx(bar_vfull + 8 * sl, KV_BYTES);
        tma_load_2d(sV + sl * KV_BYTES, &tmKV, bar_vfull + 8 * sl, 2 * D + c.f.hcol, c.f.row + k0);
        ++c.g;
        if (++c.it == c.f.n_kt) { c.it = 0; c.wk.next(sh); ++c.ord; enter_v(c); }
      };
      auto load_q = [&](const Walker& wq, int ord) {    // the K cursor has entered this item already
        if (wq.w >= n_items) return;
        const ItemLite q = ring[ord & 7];
        mbar_expect_tx(bar_qfull, Q_BYTES);
        tma_load_2d(sQ, &tmQ, bar_qfull, q.hcol, q.row + q.q_first);
      };
      Cur ck, cv;
      ck.wk.init(sh, vcta); ck.it = 0; ck.g = 0; ck.ord = 0; enter_k(ck);
      cv = ck;
      Walker wk, wnext;                  // this item / the next one (whose Q is loaded ahead)
      wk.init(sh, vcta);
      wnext = wk;
      load_q(wnext, 0);
      wnext.next(sh);
      for (int i = 0; i < NS; ++i) { load_k(ck); load_v(cv); }

      auto issue_s = [&](int g) {     // S(g) = Q K(g)^T
        const int sl = g % NS;
        if (!MMA_WAIT(bar_kfull + 8 * sl, (g / NS) & 1)) return false;
        tc_fence_after();
        const uint64_t dk = dK + (uint64_t)((sl * KV_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_ss(tmem_base + S_COL, dQ + 2 * k, dk + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
        tc_commit_1sm(bar_sfull);
        return true;
      };
      // The thread never waits for an MMA it has just issued.  The K slot of S(g+1) is refilled after PV(g) has been
      // issued, the V slot of PV(g-1) at the start of tile g -- by then those MMAs have long retired, so the waits
      // return at once.  (Waiting for S(g+1) and PV(g) right after issuing them, as this loop first did, made the
      // thread's own chain 2570 cycles per tile -- longer than the softmax chain it feeds; clock64 trace.)  A wait
      // may lag its barrier by one phase at most (parity waits), which this order guarantees.
      int gt = 0, wi = 0;
      bool pending_v = false;            // PV(g-1) has been issued and its V slot is not refilled yet
      for (; wk.w < n_items && ok; wk.next(sh), wnext.next(sh), ++wi) {
        const int n = ring[wi & 7].n_kt;
        if (!(ok = MMA_WAIT(bar_qfull, wi & 1))) break;
        if (gt > 0 && !(ok = MMA_WAIT(bar_sfree, (gt - 1) & 1))) break;   // softmax has read the previous item's last S
        tc_fence_after();
        if (!(ok = issue_s(gt))) break;
        if (n == 1) {     // the item's only S: once it has retired, K's slot and Q are free
          if (!(ok = MMA_WAIT(bar_sfull, gt & 1))) break;
          load_k(ck);
          load_q(wnext, wi + 1);
        }
        for (int it = 0; it < n && ok; ++it) {
          const int g = gt + it;
          TRACE(1, g, 0);
          if (it == 0 && n > 1) {   // S(g) of the item's first tile was issued above: free its K slot when it retires
            if (!(ok = MMA_WAIT(bar_sfull, g & 1))) break;
            load_k(ck);
          }
          if (it + 1 < n) {
            if (!(ok = MMA_WAIT(bar_sfree, g & 1))) break;          // softmax holds S(g) in registers
            tc_fence_after();
            TRACE(1, g, 1);
            if (!(ok = issue_s(g + 1))) break;
            TRACE(1, g, 2);
          }
          if (pending_v) {
            if (!(ok = MMA_WAIT(bar_pvdone, (g - 1) & 1))) break;     // PV(g-1) retired a tile ago: V's slot is free
            load_v(cv);                                                // V(g+1)
            pending_v = false;
          }
          const int sl = g % NS;
          if (!(ok = MMA_WAIT(bar_vfull + 8 * sl, (g / NS) & 1))) break;
          TRACE(1, g, 3);
          if (!(ok = MMA_WAIT(bar_pfull, g & 1))) break;            // P(g) is in smem, O has been rescaled
          tc_fence_after();
          TRACE(1, g, 4);
          const uint64_t dv = dV + (uint64_t)((sl * KV_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k)
            umma_ss(tmem_base + O_COL, dP + 2 * k, dv + (2048 >> 4) * k, idesc_pv, (it > 0 || k > 0) ? 1u : 0u);
          tc_commit_1sm(bar_pvdone);
          pending_v = true;
          TRACE(1, g, 5);
          if (it + 1 < n) {
            if (!(ok = MMA_WAIT(bar_sfull, (g + 1) & 1))) break;    // S(g+1), issued before PV(g), has retired
            load_k(ck);                                                // K(g+1+NS)
            if (it + 2 == n) load_q(wnext, wi + 1);                    // it was the item's last S: Q is free
            TRACE(1, g, 6);
          }
          TRACE(1, g, 7);
        }
        gt += n;
      }
    }
  } else {
    // ===================== softmax warps: one query row per thread =====================
    // Visibility of the 64 key columns of a tile for one query row is a contiguous column range [lo, hi):
    //   main keys: columns whose block <= qblock(row)  ->  [0, (qb+1)*main - k0)
    //   look-ahead keys: the rc copies owned by qblock(row)  ->  [T2 + qb*rc - k0, +rc)
    // so 32-column chunks are classified per warp as all-visible (no masking), none-visible (not loaded, no
    // exponentials: P = 0) or partial (one range compare per element).  Key padding (any mask, not only a
    // ragged tail) is detected per warp from the tile's 64 padding bytes and handled per element.
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;                 // row inside the query tile == TMEM lane
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t p_row = sP + (uint32_t)row * 128;     // this thread's row of the P tile / of the output staging tile
    const int T2 = sh.T2, main_ctx = sh.main_ctx, rc = sh.rc, nb = sh.nb;
    bool ok = true;
    int gt = 0;
    // Key padding.  prep_masks leaves one "any padded key" flag per 128 tokens (pad_blk); the flags of an
    // utterance are fetched one work item ahead and folded into a 32-bit mask, so the common case (no padding in a
    // key tile) costs one or two bit tests.  Bit j of the mask covers the tokens [j << flag_shift, (j + 1) <<
    // flag_shift): 128 tokens while the utterance has at most 32 flag blocks (M <= 4096), and 2^fs flag blocks
    // folded into one bit beyond that (lane j ORs its 2^fs flags), so that every token of an utterance of any
    // length maps to a bit below 32.  Only flagged tiles (or callers without pad_blk) read their 64 padding bytes
    // (2 per lane).
    const int pad_stride = (sh.M + 127) >> 7;
    const bool flags_usable = pad_blk != nullptr;
    const int fs = sh.flag_fold;                      // host: smallest fs with ceil(pad_stride / 2^fs) <= 32
    const int flag_shift = 7 + fs, flag_shift_it = 1 + fs;     // token -> mask bit; 64-key main tile index -> mask bit
    auto fetch_flags = [&](int b_) -> uint32_t {      // this lane's flags of utterance b_ (raw loads, consumed later)
      uint32_t v = 0;
      if (flags_usable) {
        if (fs == 0) {
          if (lane < pad_stride) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(pad_blk + (size_t)b_ * pad_stride + lane));
        } else {
          const int j0 = lane << fs, j1 = min(j0 + (1 << fs), pad_stride);
          for (int j = j0; j < j1; ++j) {
            uint32_t u;
            asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(u) : "l"(pad_blk + (size_t)b_ * pad_stride + j));
            v |= u;
          }
        }
      }
      return v;
    };
    // mask bit that covers token `tok` (callers without flags: the mask is all ones, "may have padding"; the fold
    // keeps every bit index below 32)
    auto flagged = [&](uint32_t mask, int tok) -> bool { return ((mask >> (tok >> flag_shift)) & 1u) != 0; };
    Walker wk, wnext;
    wk.init(sh, vcta);
    wnext = wk;
    wnext.next(sh);
    Item im;
    uint32_t flag_raw = 0;
    if (wk.w < n_items) { im = item_of(sh, wk); flag_raw = fetch_flags(im.b); }
    for (; wk.w < n_items && ok; wk.next(sh), wnext.next(sh)) {
      const int n_kt = im.ts.n_kt;
      const bool have_next = wnext.w < n_items;
      const uint32_t flag_mask = flags_usable ? __ballot_sync(0xffffffffu, flag_raw != 0) : 0xffffffffu;
      if (have_next) flag_raw = fetch_flags(fdiv(wnext.hb(sh), sh.magic_H));   // next item's utterance
      // rows past the end of the tile behave like the last valid row (their output is never stored); this keeps
      // the chunk classification uniform across the warp
      const int my_qb = row < im.q_count ? qblock_of(sh, im.q_first + row) : im.qb_hi;
      // leading main-key tiles that every row of this warp sees completely (rows of a warp are consecutive
      // tokens, lane 0 has the smallest block): they skip the range arithmetic and the masks altogether
      const int n_fullvis = min((__shfl_sync(0xffffffffu, my_qb, 0) + 1) * main_ctx, im.ts.seg0_end) / KT;
      float m_ref = -INFINITY;
      uint64_t l2a = pack2(0.f, 0.f), l2b = pack2(0.f, 0.f);   // row sum (four accumulators)
      // the P buffer doubles as this warp's output staging tile: the previous item's TMA store must have read it
      if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the thread that issues the stores
      __syncwarp();

      for (int it = 0; it < n_kt && ok; ++it) {
        const int g = gt + it;
        TRACE(0, g, 0);
        uint32_t kp_cur = 0;               // byte k: column 2*lane+k of this tile is padded
        bool has_pad = false;
        int lo = 0;
        uint32_t span = KT;
        bool all_vis[2] = {true, true}, none_vis[2] = {false, false};
        if (!(it < n_fullvis && ((flag_mask >> (it >> flag_shift_it)) & 1u) == 0)) {
          int k0, cnt; bool s1;
          im.ts.get(it, k0, cnt, s1);
          if (flagged(flag_mask, k0) || flagged(flag_mask, k0 + cnt - 1)) {
            const uint8_t* p_ = keypad + im.row_base + k0;
            const int last = cnt - 1, kr = cnt - 2 * lane;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              kp_cur |= (uint32_t)(kr > k && p_[min(2 * lane + k, last)] != 0) << (8 * k);
          }
          has_pad = __any_sync(0xffffffffu, kp_cur != 0);
          // ---- visible column range of this row.  lo and hi are non-decreasing in the lane index: the warp-wide
          //      classification needs lanes 0 and 31 only.
          int hi = 0;
          if (!s1) { hi = min(max((my_qb + 1) * main_ctx - k0, 0), cnt); }
          else if (my_qb <= nb - 1) { lo = min(max(T2 + my_qb * rc - k0, 0), cnt); hi = min(max(T2 + (my_qb + 1) * rc - k0, 0), cnt); }
          else { lo = hi = cnt; }             // past the last owner block: sees no look-ahead copy
          span = (uint32_t)(hi - lo);
          const int lo_min = __shfl_sync(0xffffffffu, lo, 0), lo_max = __shfl_sync(0xffffffffu, lo, 31);
          const int hi_min = __shfl_sync(0xffffffffu, hi, 0), hi_max = __shfl_sync(0xffffffffu, hi, 31);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int c0 = c * 32;
            all_vis[c] = !has_pad && lo_max <= c0 && c0 + 32 <= hi_min;
            none_vis[c] = hi_max <= c0 || lo_min >= c0 + 32;
          }
        }
        TRACE(0, g, 9);
        ok = SOFTMAX_WAIT(bar_sfull, g & 1);
        tc_fence_after();
        TRACE(0, g, 2);

        // ---- this row's scores -> registers (one TMEM read per score); S goes back to the MMA thread at once
        uint32_t r[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (!none_vis[c]) tmem_ld32(tlane + S_COL + c * 32, r[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(bar_sfree);
        TRACE(0, g, 1);

        // ---- mask (partial chunks) and row maximum
        uint32_t pb[2] = {0u, 0u};                  // pb[k] bit L = padding of column 2L + k
        if (has_pad) {
#pragma unroll
          for (int k = 0; k < 2; ++k) pb[k] = __ballot_sync(0xffffffffu, (kp_cur >> (8 * k)) & 0xffu);
        }
        float mx8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) mx8[k] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (none_vis[c]) continue;
          if (!all_vis[c]) {
            const int off = c * 32 - lo;
            if (!has_pad) {
#pragma unroll
              for (int j = 0; j < 32; ++j) r[c][j] = (uint32_t)(off + j) < span ? r[c][j] : 0xff800000u;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const bool padj = (pb[j & 1] >> (c * 16 + (j >> 1))) & 1u;
                r[c][j] = ((uint32_t)(off + j) < span && !padj) ? r[c][j] : 0xff800000u;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 16) {   // eight independent 3-input max chains
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mx8[k] = fmaxf(mx8[k], fmaxf(__uint_as_float(r[c][j + 2 * k]), __uint_as_float(r[c][j + 2 * k + 1])));
          }
        }
        const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                               fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        TRACE(0, g, 3);
        // ---- running maximum with lazy rescale (exact: the final normalisation uses the same reference)
        const float m_tile = mx * scale_log2;               // -inf stays -inf
        const bool grow = m_tile > m_ref + RESCALE_THRESHOLD;
        float alpha = 1.0f;
        if (grow) {
          alpha = ex2_approx(m_ref - m_tile);
          m_ref = m_tile;
          l2a = fmul2(l2a, pack2(alpha, alpha));
          l2b = fmul2(l2b, pack2(alpha, alpha));
        }
        if (it > 0) {
          ok = SOFTMAX_WAIT(bar_pvdone, (g - 1) & 1) && ok;    // PV(g-1) retired: the P buffer is free, O is stable
          tc_fence_after();
          TRACE(0, g, 7);
          if (__any_sync(0xffffffffu, grow)) {
            // rare (the first tiles of an item): 8 columns at a time, the 64 scores stay in registers meanwhile
#pragma unroll 1
            for (int c = 0; c < HD / 8; ++c) {
              uint32_t o[8];
              tmem_ld8(tlane + O_COL + c * 8, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
              tmem_st8(tlane + O_COL + c * 8, o);
            }
            tmem_st_wait();
          }
        }
        TRACE(0, g, 4);
        const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
        const uint64_t sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_use, -m_use);

        // ---- P = exp2(s * scale - m), row sum, bf16 P -> smem (K-major operand layout: this thread's row is 128
        //      contiguous bytes whose 16-byte chunks are XOR-swizzled with the row index)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pk[16];
          if (none_vis[c]) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float a0, a1, p0, p1;
              unpack2(ffma2(pack2(__uint_as_float(r[c][j]), __uint_as_float(r[c][j + 1])), sc2, nm2), a0, a1);
              // W2VS_ATTN_EMUL_MOD = n > 0 moves every n-th pair of exponentials to the FMA pipe (ex2_emul2).  Measured
              // at cfg3 (B200, per layer call): n = 0: 484 us, 8: 491, 5: 494, 4: 498, 3: 506 -- with three CTAs per
              // SM the kernel is bound by issued instructions, not by the MUFU pipe, so the default is off.
#ifndef W2VS_ATTN_EMUL_MOD
#define W2VS_ATTN_EMUL_MOD 0
#endif
              if (W2VS_ATTN_EMUL_MOD > 0 && (j >> 1) % (W2VS_ATTN_EMUL_MOD > 0 ? W2VS_ATTN_EMUL_MOD : 1) == W2VS_ATTN_EMUL_MOD - 1) {
                ex2_emul2(a0, a1, p0, p1);
              } else {
                p0 = ex2_approx(a0);
                p1 = ex2_approx(a1);
              }
              if (j & 2) l2b = fadd2(l2b, pack2(p0, p1)); else l2a = fadd2(l2a, pack2(p0, p1));
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(p_row + ((uint32_t)((c * 4 + q) ^ (row & 7)) << 4), make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]));
        }
        TRACE(0, g, 5);
        fence_async_smem();          // P (generic-proxy writes) -> visible to the tensor core's async-proxy reads
        tc_fence_before();           // orders the rescale's TMEM accesses before the arrive
        __syncwarp();
        if (lane == 0) mbar_arrive_local(bar_pfull);
        TRACE(0, g, 6);
      }
      gt += n_kt;

      // ---- epilogue: O / l -> bf16 -> ctx.  The next item is worked out while PV(last) is still in flight.
      float la, lb;
      unpack2(fadd2(l2a, l2b), la, lb);
      const float l = la + lb;
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      const int out_row0 = im.row_base + im.q_first + quarter * 32;      // first token row this warp writes
      const int out_col = im.h * HD;
      const int n_valid = min(max(im.q_count - quarter * 32, 0), 32);    // valid rows of this warp
      TRACE(0, gt - 1, 11);
      if (have_next) im = item_of(sh, wnext);
      ok = SOFTMAX_WAIT(bar_pvdone, (gt - 1) & 1) && ok;                    // last PV retired: O is final, P buffer is free
      tc_fence_after();
      TRACE(0, gt - 1, 12);
      if (n_valid == 32) {
        // full warp tile: rows -> 128B-swizzled staging (this warp's 32 rows of the P buffer) -> one TMA store
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tlane + O_COL + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const uint32_t a = p_row + ((uint32_t)((c * 4 + gq) ^ (row & 7)) << 4);
            sts128(a, make_uint4(pack_bf16x2(__uint_as_float(o[8 * gq + 0]) * inv, __uint_as_float(o[8 * gq + 1]) * inv),
                                 pack_bf16x2(__uint_as_float(o[8 * gq + 2]) * inv, __uint_as_float(o[8 * gq + 3]) * inv),
                                 pack_bf16x2(__uint_as_float(o[8 * gq + 4]) * inv, __uint_as_float(o[8 * gq + 5]) * inv),
                                 pack_bf16x2(__uint_as_float(o[8 * gq + 6]) * inv, __uint_as_float(o[8 * gq + 7]) * inv)));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d_(&tmCtx, sP + (uint32_t)quarter * 32 * 128, out_col, out_row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        // ragged warp tile (end of the main or of the look-ahead segment): guarded per-row stores
        bf16* dst = ctx + (size_t)(out_row0 + lane) * D + out_col;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tlane + O_COL + c * 32, o);
          tmem_ld_wait();
          if (lane < n_valid) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(o[8 * gq + 0]) * inv, __uint_as_float(o[8 * gq + 1]) * inv);
              v.y = pack_bf16x2(__uint_as_float(o[8 * gq + 2]) * inv, __uint_as_float(o[8 * gq + 3]) * inv);
              v.z = pack_bf16x2(__uint_as_float(o[8 * gq + 4]) * inv, __uint_as_float(o[8 * gq + 5]) * inv);
              v.w = pack_bf16x2(__uint_as_float(o[8 * gq + 6]) * inv, __uint_as_float(o[8 * gq + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + c * 32 + gq * 8) = v;
            }
          }
        }
      }
      tc_fence_before();   // orders these TMEM reads before this warp's next p_full arrive (next item overwrites O)
      TRACE(0, gt - 1, 13);
    }
    if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive its last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == N_SOFTMAX_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}
}  // namespace

w2vs_status_t launch_attention_tc(const AttnArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(a.dtype == W2VS_BF16, "tcgen05 attention takes bf16");
  W2VS_REQUIRE(a.D == a.heads * HD, "attention head_dim must be 64");
  W2VS_REQUIRE(a.n_step_q == 0, "tcgen05 attention implements the full-utterance (block mask) mode");
  W2VS_REQUIRE(((uintptr_t)a.qkv & 15) == 0 && ((uintptr_t)a.ctx & 15) == 0, "attention buffers must be 16-byte aligned");
  Shape sh;
  sh.T2 = a.T2; sh.main_ctx = a.main_ctx; sh.rc = a.rc; sh.rcd = a.rc > 0 ? a.rc : 1; sh.nb = a.T2 / a.main_ctx;
  sh.M = a.T2 + (a.rc > 0 ? sh.nb * a.rc : 0);
  sh.D = a.D; sh.H = a.heads; sh.B = a.B;
  sh.n_main_tiles = (a.T2 + QT - 1) / QT;
  sh.n_tiles = sh.n_main_tiles + (sh.M - a.T2 + QT - 1) / QT;
  alignas(64) CUtensorMap tmq, tmkv;   // same token buffer, boxes of 128 (Q) and 64 (K, V) rows x 64 columns
  W2VS_TRY(tc::make_map(&tmq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.qkv, (uint64_t)3 * a.D, (uint64_t)a.B * sh.M,
                        (uint64_t)3 * a.D, HD, QT, CU_TENSOR_MAP_SWIZZLE_128B));
  W2VS_TRY(tc::make_map(&tmkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.qkv, (uint64_t)3 * a.D, (uint64_t)a.B * sh.M,
                        (uint64_t)3 * a.D, HD, KT, CU_TENSOR_MAP_SWIZZLE_128B));
  static PerDeviceOnce attr_once;   // the attribute belongs to the current device's copy of the kernel
  bool& attr_done = attr_once.here();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_error("attn_tc smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    // three CTAs per SM need 3 x 65 KB: ask for the largest shared-memory carve-out (the default sizes it for one CTA)
    e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { set_error("attn_tc carve-out attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    attr_done = true;
  }
  const int64_t n_items = (int64_t)sh.n_tiles * a.heads * a.B;
  W2VS_REQUIRE(n_items < (1ll << 31), "attention problem too large");
  const int max_ctas = CTAS_PER_SM * tc::num_sms();     // persistent: three single-pipeline CTAs per SM
  const int grid = (int)(n_items < max_ctas ? n_items : max_ctas);
  sh.HB = a.heads * a.B;
  sh.n_vcta = grid;
  sh.flag_fold = 0;
  while (((((sh.M + 127) >> 7) + (1 << sh.flag_fold) - 1) >> sh.flag_fold) > 32) ++sh.flag_fold;
  // block of G pairs: the largest divisor of H*B that is at most 32 (the heads alone give 1, 2, 4, 8, 16)
  sh.G = 1;
  for (int d = 2; d <= 32; ++d) if (sh.HB % d == 0) sh.G = d;
  sh.S = sh.G * sh.n_tiles;
  sh.g_div = sh.n_vcta / sh.S;
  sh.g_mod = sh.n_vcta % sh.S;
  // magic multipliers: floor(x / d) == umulhi(x, ceil(2^32 / d)) holds for x < 2^32 / d
  auto magic = [](int d) { return (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
  W2VS_REQUIRE((int64_t)sh.M * a.main_ctx < (1ll << 32) && (int64_t)sh.M * sh.rcd < (1ll << 32) &&
               (int64_t)sh.HB * a.heads < (1ll << 32) && (int64_t)sh.S * sh.G < (1ll << 32),
               "attention shape out of range for the fast index arithmetic");
  sh.magic_main = a.main_ctx >= 2 ? magic(a.main_ctx) : 0u;
  sh.magic_rcd = sh.rcd >= 2 ? magic(sh.rcd) : 0u;   // d == 1 handled below
  sh.magic_H = a.heads >= 2 ? magic(a.heads) : 0u;
  sh.magic_G = sh.G >= 2 ? magic(sh.G) : 0u;
  const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  alignas(64) CUtensorMap tmc;   // ctx [B*M, D]: one warp's 32 rows x 64 columns per store
  W2VS_TRY(tc::make_map(&tmc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.ctx, (uint64_t)a.D, (uint64_t)a.B * sh.M,
                        (uint64_t)a.D, HD, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  attn_tc_kernel<<<grid, N_THREADS, SMEM_BYTES, st>>>(tmq, tmkv, tmc, a.keypad, a.pad_blk, (bf16*)a.ctx, sh,
                                                      (int)n_items, scale_log2);
  W2VS_CHECK_LAUNCH("attn_tc_kernel");
  return W2VS_OK;
}

w2vs_status_t debug_read_attn_tc_fault(int* out) {
  int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_attn_tc_fault, sizeof(int));
  if (e != cudaSuccess) { set_error("read g_attn_tc_fault: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  *out = v;
  return W2VS_OK;
}

}  // namespace w2vs
