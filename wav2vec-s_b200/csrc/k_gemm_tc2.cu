// bf16 GEMM on tcgen05 tensor cores, CTA-pair (cta_group::2) version:
//     C[M,N] = A[M,K] . W[N,K]^T + bias  (+GELU)  (+fp32 residual, in place)     fp32 accumulation in TMEM
//
// Every matrix product of the bf16 encoder runs here: the strided feature-extractor convs 1..6 (implicit
// GEMM over the channels-last activation), post_extract_proj, QKV / out_proj, fc1(+GELU) / fc2(+residual)
// (wav2vec2.py:725,568,950-973; modules/multihead_attention.py:162-194).
//
// One persistent cluster of two CTAs per SM pair; a cluster tile is 256 (M) x BN (N):
//   * each CTA TMA-loads its own 128 rows of A and its own BN/2 rows of W per 64-wide K stage (128B-swizzled
//     smem), signalling the leader CTA's "full" mbarrier (cp.async.bulk.tensor ... .cta_group::2);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 (M=256, N=BN, K=16): both tensor cores run, each
//     reads its own A rows and both halves of W, accumulators land in each CTA's own TMEM (rows 0-127 /
//     128-255); tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs;
//   * 8 epilogue warps per CTA (2 per TMEM lane quarter, each half of the columns; 16 for the GELU product) drain the accumulator in
//     32-column chunks: tcgen05.ld -> +bias (smem) -> [GELU] -> [+ residual chunk, TMA-loaded into the
//     staging slot] -> swizzled staging slot -> TMA store.  Two accumulator stages (2 x BN TMEM columns)
//     overlap the epilogue of tile i with the MMAs of tile i+1.
//
// Implicit-GEMM view of Conv1d(C_in -> C_out, k, stride s) on activations [rows, C_in]: output row r needs the
// k*C_in contiguous inputs starting at row r*s.  The A tensor map describes the matrix [rows, a_row_len =
// s*C_in]; K index kk >= a_row_len wraps to the next row: (x, y) = (kk % a_row_len, r + kk / a_row_len).
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
__device__ int g_tc2_fault = 0;  // set when a pipeline wait timed out (diagnostics)
}
#define W2VS_TC_FAULT_FLAG (&::w2vs::g_tc2_fault)
#include "tc_common.cuh"

namespace w2vs {

namespace {

constexpr int BM = 128, BK = 64;           // per-CTA rows; K per stage
constexpr int A_STAGE_BYTES = BM * BK * 2;
// Epilogue warps per CTA: EW = 8 (two per TMEM lane quarter, each half of the tile's columns) or 16 (four per
// quarter, a quarter of the columns each) -- a template parameter of the kernel.  Threads = 32 EW + 64.
// Warp roles.  The SMSP arbiter favours the highest warp id, and the TMA-producer / MMA-issuer threads sit on
// the critical path of every stage, so they take the two highest ids; the epilogue warps are 0..7.

// Columns per epilogue chunk: 32 for bf16 outputs; W2VS_GEMM_F32_CHUNK for fp32 outputs.  With 32 fp32 columns a
// staging slot is 4 KB and only four TMA ring stages fit next to two slots per warp; 16 columns (2 KB slots, five
// stages, still double buffered) was measured: the K = 4096 product gains what the single-slot variant below gains
// (16.8 -> 14.8 ms per step), but the K = 1024 product loses to the doubled per-chunk overhead (5.4 -> 6.1 ms), so
// the default is 32 plus W2VS_GEMM_LONGK_SINGLE_SLOT.
#ifndef W2VS_GEMM_F32_CHUNK
#define W2VS_GEMM_F32_CHUNK 32
#endif
// In-place residual (C += A.W^T + bias): true = the staged tile is added to C by the TMA reduction path
// (cp.reduce.async.bulk.tensor .add, fp32 adds in L2) -- no residual load, no second pass through smem;
// false = the residual chunk is TMA-loaded into the staging slot and added by the epilogue threads.
// Measured on B200 (cfg3, 24 layers): out_proj 5.5 ms -> 6.4 ms with the reduction path (fc2 unchanged): the
// L2 reduction units sustain less than the 4.3 TB/s the K=1024 product moves, so the default stays 0.
#ifndef W2VS_GEMM_TMA_REDUCE
#define W2VS_GEMM_TMA_REDUCE 0
#endif
constexpr bool kTmaReduce = W2VS_GEMM_TMA_REDUCE != 0;
constexpr int SMEM_LIMIT = 232448;

// SLOTS: staging slots per epilogue warp.  Two let the residual chunk c+1 be loaded while chunk c is processed; one
// frees 32 KB (fp32) for a fifth TMA ring stage -- used for the long-K in-place product (fc2, K = 4096), whose
// epilogue has 4x the MMA time of a K = 1024 tile to hide in and whose mainloop streams its A operand from HBM.
template <int BN, typename TC, int SLOTS = 2, int EW = 8> struct Cfg2 {
  static constexpr int kColsPerWarp = BN / (EW / 4);                     // columns of the tile one epilogue warp drains
  static constexpr int kBStageBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBStageBytes;
  static constexpr int kChunk = sizeof(TC) == 4 ? W2VS_GEMM_F32_CHUNK : 32;
  static constexpr int kSlotBytes = 32 * kChunk * (int)sizeof(TC);
  static constexpr int kStagingBytes = EW * SLOTS * kSlotBytes;
  static constexpr int kMiscBytes = EW * kColsPerWarp * 4 /*bias*/ + 768 /*barriers*/ + 1024 /*align*/;
  static constexpr int kStagesFit = (SMEM_LIMIT - kStagingBytes - kMiscBytes) / kStageBytes;
#ifndef W2VS_GEMM_MAX_STAGES
#define W2VS_GEMM_MAX_STAGES 8
#endif
  static constexpr int kStages = kStagesFit > W2VS_GEMM_MAX_STAGES ? W2VS_GEMM_MAX_STAGES : kStagesFit;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kMiscBytes;
  static constexpr int kChunksPerWarp = kColsPerWarp / kChunk;
  static_assert(kChunksPerWarp >= 1, "tile too narrow for this many epilogue warps");
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
  static_assert(kStages >= 3, "pipeline too shallow");
};

using namespace tc;

// RED (split-K for the in-place fp32 products of a few hundred rows, the incremental steps of a batch of streams):
// the tile index also walks `ksplit` K ranges, every range's partial tile goes into C through the TMA reduction path
// (fp32 adds at L2, so C += sum of the partials in whatever order they arrive; the bias rides on range 0).  A
// 384 x 1024 product has 32 cluster tiles for 74 clusters, each streaming its whole K through one SM pair; split
// four ways every SM pair has a tile and a quarter of the K loop.
template <int BN, typename TC, int SLOTS, int EW, bool RED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * EW + 64, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int has_residual,
                int M, int N, int K, int a_row_len, int gelu, int n_batch, int a_batch_rows, int w_batch_rows,
                int ksplit) {
  using C2 = Cfg2<BN, TC, SLOTS, EW>;
  constexpr bool kTmaReduce = w2vs::kTmaReduce || RED;
  constexpr int S = C2::kStages;
  constexpr bool kF32 = sizeof(TC) == 4;
  constexpr int N_EPI_WARPS = EW, PRODUCER_WARP = EW, MMA_WARP = EW + 1, CPW = C2::kColsPerWarp;
  constexpr int CHUNK = C2::kChunk;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + S * A_STAGE_BYTES;
  const uint32_t sStage = sB + S * C2::kBStageBytes;                 // epilogue staging slots (1024-aligned)
  const uint32_t sBias = sStage + C2::kStagingBytes;                 // [8 warps][BN/2] floats
  const uint32_t bars = sBias + N_EPI_WARPS * CPW * 4;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * S, bar_tfull = bars + 16 * S, bar_tempty = bar_tfull + 16;
  const uint32_t bar_res = bar_tempty + 16;                          // [8 warps][2 slots]
  const uint32_t tmem_slot = bar_res + 8 * 2 * N_EPI_WARPS;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to smem_base
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(gen_base + (sBias - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + 2 * BM - 1) / (2 * BM);
  const int tiles_pb = tiles_m * tiles_n;          // tiles per product; n_batch > 1: the tile index also walks the batch
  const int n_tiles = tiles_pb * (RED ? ksplit : n_batch);   // RED: the slow index is the K range, not the product
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per = RED ? (num_kb + ksplit - 1) / ksplit : num_kb;   // K blocks per range

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 2 * N_EPI_WARPS); }
    for (int i = 0; i < 2 * N_EPI_WARPS; ++i) mbar_init(bar_res + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C2::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();          // the peer's barriers are initialised and its TMEM is allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // launched as a programmatic dependent (GemmArgs::pdl) everything above overlaps the predecessor; a no-op otherwise.
  // (No griddepcontrol.launch_dependents here: releasing this kernel's successors early -- the step attention / the
  //  LayerNorm of an incremental step -- measured 2.36 -> 2.8 ms per 16-stream step.)
  // The weights do not depend on the predecessor: the producer requests the W half of the first ring stages before
  // it waits.
  if (warp != PRODUCER_WARP) pdl_wait();

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer (both CTAs) =====================
    // (single-thread roles are entered through elect_one(), see tc_common.cuh: behind `lane == 0` every TMA / MMA
    //  instruction is wrapped in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop)
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
      const uint32_t full_leader = mapa(bar_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      int pre = 0;     // leading ring stages (first uses, first tile) that are armed and whose W half is in flight
      if (cluster_id < n_tiles) {
        const int hi = cluster_id / tiles_pb, rem = cluster_id - hi * tiles_pb;
        const int n0 = (rem % tiles_n) * BN + (int)rank * (BN / 2) + (RED ? 0 : hi) * w_batch_rows;
        const int kb0 = RED ? hi * kb_per : 0, kb1 = RED ? min(kb0 + kb_per, num_kb) : num_kb;
        pre = min(S, kb1 - kb0);
        for (int i = 0; i < pre; ++i) {
          if (leader) mbar_expect_tx(bar_full + 8 * i, 2 * C2::kStageBytes);
          tma_load_2d_2sm(sB + i * C2::kBStageBytes, &tmB, full_leader + 8 * i, (kb0 + i) * BK, n0);
        }
      }
      pdl_wait();
      for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
        const int hi = tile / tiles_pb, rem = tile - hi * tiles_pb;
        const int bt = RED ? 0 : hi;
        const int m0 = (rem / tiles_n) * (2 * BM) + (int)rank * BM + bt * a_batch_rows;
        const int n0 = (rem % tiles_n) * BN + (int)rank * (BN / 2) + bt * w_batch_rows;
        const int kb0 = RED ? hi * kb_per : 0, kb1 = RED ? min(kb0 + kb_per, num_kb) : num_kb;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int kk = kb * BK;
          if (pre > 0) {
            --pre;
          } else {
            if (!(ok = mbar_wait(bar_empty + 8 * stage, phase ^ 1))) break;
            if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * C2::kStageBytes);
            tma_load_2d_2sm(sB + stage * C2::kBStageBytes, &tmB, full_leader + 8 * stage, kk, n0);
          }
          tma_load_2d_2sm(sA + stage * A_STAGE_BYTES, &tmA, full_leader + 8 * stage, kk % a_row_len, m0 + kk / a_row_len);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=256 (cta_group::2)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      bool ok = true;
      for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
        if (!(ok = mbar_wait_cluster(bar_tempty + 8 * as, aphase ^ 1))) break;   // peer CTA's epilogue arrives here
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(as * BN);
        const int kb0 = RED ? (tile / tiles_pb) * kb_per : 0, kb1 = RED ? min(kb0 + kb_per, num_kb) : num_kb;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!(ok = mbar_wait(bar_full + 8 * stage, phase))) break;
          tc_fence_after();
          const uint32_t a_addr = sA + stage * A_STAGE_BYTES, b_addr = sB + stage * C2::kBStageBytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_2sm(tmem_c, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                          (kb > kb0 || k > 0) ? 1u : 0u);
          tc_commit_2sm(bar_empty + 8 * stage);   // both CTAs may refill this stage once the MMAs retire
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        tc_commit_2sm(bar_tfull + 8 * as);        // accumulator complete in both CTAs
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int e = warp;                      // 0..7
    const int quarter = warp & 3;            // TMEM lanes this warp may touch: 32*quarter ..
    const int half = e >> 2;                 // column group of the tile (half with 8 epilogue warps, quarter with 16)
    const uint32_t slot0 = sStage + (uint32_t)e * SLOTS * C2::kSlotBytes;
    const uint32_t my_res_bar = bar_res + 16 * e;
    const uint32_t tempty_leader = mapa(bar_tempty, 0);
    uint32_t res_phase = 0;                  // bit s = parity of residual slot s
    int as = 0;
    uint32_t aphase = 0;
    bool ok = true;
    for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
      const int hi = tile / tiles_pb, rem = tile - hi * tiles_pb;
      const int bt = RED ? 0 : hi;
      const int m0 = (rem / tiles_n) * (2 * BM) + (int)rank * BM;
      const int n0 = (rem % tiles_n) * BN;
      const int row0 = m0 + quarter * 32;
      const int colw = n0 + half * CPW;                // first column of this warp
      // bias of this warp's columns -> its private smem strip
      float* bs = bias_s + e * CPW;
      __syncwarp();
#pragma unroll
      for (int i = lane; i < CPW; i += 32) bs[i] = (bias != nullptr && colw + i < N && !(RED && hi > 0)) ? bias[bt * N + colw + i] : 0.f;
      // residual chunk 0 prefetch (overlaps the wait for the accumulator)
      if (SLOTS == 2 && !kTmaReduce && has_residual && elect_one()) {
        bulk_wait_read<0>();                           // earlier stores from slot 0/1 have been read out
        mbar_expect_tx(my_res_bar, C2::kSlotBytes);
        tma_load_2d(slot0, &tmC, my_res_bar, colw, row0);
      }
      __syncwarp();                                    // bias strip visible to the whole warp
      ok = mbar_wait(bar_tfull + 8 * as, aphase);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * CPW);
#pragma unroll 1
      for (int c = 0; c < C2::kChunksPerWarp; ++c) {
        const int sl = c & (SLOTS - 1);
        const uint32_t slot = slot0 + sl * C2::kSlotBytes;
        if (SLOTS == 1) {
          // single slot: wait until the previous store has read it, then fetch this chunk's residual into it (the
          // accumulator load and the bias arithmetic below run while it is in flight)
          if (elect_one()) {
            bulk_wait_read<0>();
            if (!kTmaReduce && has_residual) {
              mbar_expect_tx(my_res_bar, C2::kSlotBytes);
              tma_load_2d(slot, &tmC, my_res_bar, colw + c * CHUNK, row0);
            }
          }
        } else if (!kTmaReduce && has_residual) {
          if (c + 1 < C2::kChunksPerWarp && elect_one()) {
            bulk_wait_read<0>();                       // the store that used the other slot has drained it
            mbar_expect_tx(my_res_bar + 8 * (sl ^ 1), C2::kSlotBytes);
            tma_load_2d(slot0 + (sl ^ 1) * C2::kSlotBytes, &tmC, my_res_bar + 8 * (sl ^ 1), colw + (c + 1) * CHUNK, row0);
          }
        } else if (elect_one()) {
          bulk_wait_read<1>();                         // the store issued two chunks ago (same slot) is done
        }
        uint32_t r[CHUNK];
        tmem_ld_chunk<CHUNK>(taddr + c * CHUNK, r);
        tmem_ld_wait();
        float v[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; j += 2) {
          const float2 bb = *reinterpret_cast<const float2*>(bs + c * CHUNK + j);   // smem broadcast
          unpack2(fadd2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pack2(bb.x, bb.y)), v[j], v[j + 1]);
          if (gelu) gelu2<TC>(v[j], v[j + 1]);
        }
        if (!kTmaReduce && has_residual) {
          ok = mbar_wait(my_res_bar + 8 * sl, (res_phase >> sl) & 1u);
          res_phase ^= 1u << sl;
        }
        __syncwarp();                                  // lane 0's wait_group / everyone's mbarrier wait done
        if (kF32) {
          // CHUNK 32: 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row `lane` lives at chunk j ^ (lane & 7)
          // CHUNK 16:  64-byte rows, SWIZZLE_64B:  ... at chunk j ^ ((lane >> 1) & 3)
          const uint32_t rowaddr = slot + lane * (CHUNK * 4);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            const uint32_t a = rowaddr + ((uint32_t)(j ^ (CHUNK == 32 ? (lane & 7) : ((lane >> 1) & 3))) << 4);
            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (!kTmaReduce && has_residual) {
              const uint4 rr = lds128(a);
              o.x += __uint_as_float(rr.x); o.y += __uint_as_float(rr.y);
              o.z += __uint_as_float(rr.z); o.w += __uint_as_float(rr.w);
            }
            sts128(a, make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
          }
        } else {
          // 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row `lane` lives at chunk j ^ ((lane >> 1) & 3)
          const uint32_t rowaddr = slot + lane * 64;
#pragma unroll
          for (int j = 0; j < CHUNK / 8; ++j) {
            const uint32_t a = rowaddr + ((uint32_t)(j ^ ((lane >> 1) & 3)) << 4);
            sts128(a, make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
          }
        }
        fence_async_smem();                            // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (elect_one()) {
          if (kTmaReduce && has_residual) tma_reduce_add_2d(&tmC, slot, colw + c * CHUNK, row0);
          else if (n_batch > 1) tma_store_3d(&tmC, slot, colw + c * CHUNK, bt, row0);   // {column, product, row}: clipped at N
          else tma_store_2d(&tmC, slot, colw + c * CHUNK, row0);
          bulk_commit();
        }
      }
      // this warp has read its part of the accumulator: release the TMEM stage to the leader's MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (elect_one()) bulk_wait_read<0>();              // smem must stay valid until the last store has read it
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();          // both CTAs are done with both TMEMs
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C2::kTmemCols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
template <int BN, typename TC, int SLOTS = 2, int EW = 8, bool RED = false>
w2vs_status_t launch_bn(const GemmArgs& g, cudaStream_t st, int ksplit = 1) {
  using C2 = Cfg2<BN, TC, SLOTS, EW>;
  alignas(64) CUtensorMap tmA, tmB, tmC;
  const int64_t a_row_len = g.lda;
  const uint64_t a_inner = (uint64_t)(g.K <= a_row_len ? g.K : a_row_len);
  W2VS_TRY(make_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.A, a_inner, (uint64_t)g.a_rows, (uint64_t)a_row_len,
                    BK, BM, CU_TENSOR_MAP_SWIZZLE_128B));
  const int n_batch = g.batch > 1 ? g.batch : 1;
  W2VS_TRY(make_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.W, (uint64_t)g.K,
                    (uint64_t)((n_batch - 1) * g.w_batch_rows + g.N), (uint64_t)g.K, BK, BN / 2,
                    CU_TENSOR_MAP_SWIZZLE_128B));
  if (n_batch > 1) {
    // one product per batch index: the third map dimension clips every store at the product's own N columns
    if (sizeof(TC) == 4)
      W2VS_TRY(make_map3(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.C, (uint64_t)g.N, (uint64_t)n_batch, (uint64_t)g.M,
                         (uint64_t)g.c_batch_stride, (uint64_t)g.ldc, C2::kChunk, 32,
                         C2::kChunk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
    else
      W2VS_TRY(make_map3(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.C, (uint64_t)g.N, (uint64_t)n_batch, (uint64_t)g.M,
                         (uint64_t)g.c_batch_stride, (uint64_t)g.ldc, C2::kChunk, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  } else if (sizeof(TC) == 4)
    W2VS_TRY(make_map(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.C, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc,
                      C2::kChunk, 32, C2::kChunk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
  else
    W2VS_TRY(make_map(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.C, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc,
                      C2::kChunk, 32, CU_TENSOR_MAP_SWIZZLE_64B));
  static PerDeviceOnce attr_once;   // the attribute belongs to the current device's copy of the kernel
  bool& attr_done = attr_once.here();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN, TC, SLOTS, EW, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C2::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc2 smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    attr_done = true;
  }
  const int tiles = (int)(ceil_div64(g.M, 2 * BM) * ceil_div64(g.N, BN)) * (RED ? ksplit : n_batch);
  const int max_clusters = num_sms() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  // plain stream launch: as a programmatic dependent (PDL) this kernel's 200 KB CTAs cannot become resident early
  // anyway, and 16-stream incremental steps measured 7 % slower with it
  if (g.pdl && g_pdl_on)
    launch_pdl(gemm_tc2_kernel<BN, TC, SLOTS, EW, RED>, dim3(2 * clusters), dim3(32 * EW + 64), (size_t)C2::kSmemBytes, st,
               tmA, tmB, tmC, g.bias, g.residual != nullptr ? 1 : 0, g.M, g.N, g.K, (int)a_row_len,
               (g.flags & W2VS_EPI_GELU) ? 1 : 0, n_batch, (int)g.a_batch_rows, (int)g.w_batch_rows, ksplit);
  else
  gemm_tc2_kernel<BN, TC, SLOTS, EW, RED><<<2 * clusters, 32 * EW + 64, C2::kSmemBytes, st>>>(
      tmA, tmB, tmC, g.bias, g.residual != nullptr ? 1 : 0, g.M, g.N, g.K, (int)a_row_len,
      (g.flags & W2VS_EPI_GELU) ? 1 : 0, n_batch, (int)g.a_batch_rows, (int)g.w_batch_rows, ksplit);
  if (g_prof_on) {
    char name[96];
    char bat[24] = "";
    if (n_batch > 1) snprintf(bat, sizeof(bat), ",batch=%d", n_batch);
    if (RED) snprintf(bat, sizeof(bat), ",ksplit=%d", ksplit);
    snprintf(name, sizeof(name), "gemm_tc2_kernel[M=%d,N=%d,K=%d,lda=%lld,%s%s%s%s]", g.M, g.N, g.K, (long long)g.lda,
             sizeof(TC) == 4 ? "f32" : "bf16", g.residual ? ",res" : "", (g.flags & W2VS_EPI_GELU) ? ",gelu" : "", bat);
    W2VS_CHECK_LAUNCH(name);
  } else {
    W2VS_CHECK_LAUNCH("gemm_tc2_kernel");
  }
  return W2VS_OK;
}

template <typename TC>
w2vs_status_t launch_typed(const GemmArgs& g, cudaStream_t st) {
  // Widest tile that still gives every SM pair a tile: with a few hundred rows (incremental steps of a batch of
  // streams) 256-wide tiles would leave most of the GPU idle while a handful of clusters stream all of W.
  const int64_t m_tiles = ceil_div64(g.M, 2 * BM);
  const int clusters = num_sms() / 2;
#ifndef W2VS_GEMM_LONGK_SINGLE_SLOT
#define W2VS_GEMM_LONGK_SINGLE_SLOT 1
#endif
#ifndef W2VS_GEMM_SINGLE_SLOT_MIN_K
#define W2VS_GEMM_SINGLE_SLOT_MIN_K 2048     // measured with 1024: the K = 1024 product goes 5.5 -> 6.0 ms per step
#endif
  if (W2VS_GEMM_LONGK_SINGLE_SLOT && sizeof(TC) == 4 && g.K >= W2VS_GEMM_SINGLE_SLOT_MIN_K && g.N % 256 == 0 && m_tiles * (g.N / 256) >= clusters)
    return launch_bn<256, TC, 1>(g, st);          // fp32 output, long K: five ring stages instead of four
#ifndef W2VS_GEMM_BF16_EPI_WARPS
#define W2VS_GEMM_BF16_EPI_WARPS 16
#endif
  // bf16 output with GELU (fc1), full-width tiles: 16 epilogue warps (four per TMEM lane quarter) with one staging
  // slot each -- the accumulator drain (bias, GELU, pack, store) paces this product and more warps hide its
  // latencies: 16.7 -> 15.8 ms per step (same-box A/B); the bias-only QKV product is 2 % slower with 16, so it
  // keeps 8, and so does the fp32 + residual K = 1024 product (5.5 -> 5.75 ms with 16 warps and one slot each).
  if (W2VS_GEMM_BF16_EPI_WARPS == 16 && sizeof(TC) == 2 && (g.flags & W2VS_EPI_GELU) && g.N % 256 == 0 &&
      m_tiles * (g.N / 256) >= clusters)
    return launch_bn<256, TC, 1, 16>(g, st);
#ifndef W2VS_GEMM_WIDE_FILL_NUM
#define W2VS_GEMM_WIDE_FILL_NUM 2      // quarters of the clusters a wider tile must still fill (4 = the round-1 rule; 16 streams: 4 / 3 / 2 -> 2.40 / 2.36 ms, then 1.90 / 1.81)
#endif
#ifndef W2VS_GEMM_SPLITK
#define W2VS_GEMM_SPLITK 1
#endif
  // in-place fp32 product with fewer 64-wide tiles than half the SM pairs: split K (see RED above).  Tile width 64 or
  // 128, whichever leaves the shorter K range per cluster (the chain of dependent TMA stages is what these products
  // cost): out_proj / fc2 of 16 streams run as 16 tiles of 128 columns x 4 ranges.
  if (W2VS_GEMM_SPLITK && (g.flags & W2VS_EPI_SPLITK) && sizeof(TC) == 4 && g.residual != nullptr && g.batch <= 1 && g.N % 64 == 0 && g.K % BK == 0 &&
      g.K <= g.lda && m_tiles * (g.N / 64) * 2 <= clusters) {
    const int num_kb = g.K / BK;
    auto ranges = [&](int64_t tiles) {
      int ks = (int)(clusters / tiles);                   // ranges that still give every cluster at most one tile
      while (ks > 1 && (num_kb / ks < 4 || (ks - 1) * ((num_kb + ks - 1) / ks) >= num_kb)) --ks;   // >= 4 K blocks each, none empty
      return ks;
    };
    const int ks64 = ranges(m_tiles * (g.N / 64));
    const int ks128 = g.N % 128 == 0 ? ranges(m_tiles * (g.N / 128)) : 1;
    if (ks128 > 1 && (num_kb + ks128 - 1) / ks128 < (num_kb + ks64 - 1) / ks64) return launch_bn<128, TC, 2, 8, true>(g, st, ks128);
    if (ks64 > 1) return launch_bn<64, TC, 2, 8, true>(g, st, ks64);
  }
  if (g.N <= 64) return launch_bn<64, TC>(g, st);   // one group of the positional conv (N = D / groups = 48 or 64)
  // (half of the clusters busy for one wave beat 1.3 - 1.7 waves of tiles half as wide: of 16 streams, fc1, 384 x 4096,
  //  has 64 tiles of 128 columns for 74 clusters, QKV 48)
  const int64_t enough = (int64_t)clusters * W2VS_GEMM_WIDE_FILL_NUM / 4;
  if (g.N % 256 == 0 && (m_tiles * (g.N / 256) >= enough || g.N % 128 != 0)) return launch_bn<256, TC>(g, st);
  if (g.N % 128 == 0 && (m_tiles * (g.N / 128) >= enough || g.N % 64 != 0)) return launch_bn<128, TC>(g, st);
  if (g.N % 64 == 0) return launch_bn<64, TC>(g, st);
  return launch_bn<256, TC>(g, st);
}

}  // namespace

w2vs_status_t launch_gemm_tc2(const GemmArgs& g, cudaStream_t st) {
  W2VS_REQUIRE(g.dtype_ab == W2VS_BF16, "tcgen05 GEMM takes bf16 operands");
  W2VS_REQUIRE(g.K % 8 == 0 && g.N % 8 == 0, "GEMM needs K % 8 == 0 and N % 8 == 0");
  W2VS_REQUIRE(g.lda % 8 == 0 && g.ldc % 8 == 0, "GEMM leading dims must be multiples of 8");
  W2VS_REQUIRE(g.K <= g.lda || g.lda % BK == 0, "wrapped (conv) A rows need lda % 64 == 0");
  W2VS_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0 && ((uintptr_t)g.C & 15) == 0,
               "GEMM operands must be 16-byte aligned");
  W2VS_REQUIRE(g.residual == nullptr || (g.dtype_c == W2VS_F32 && g.residual == (const float*)g.C),
               "tcgen05 GEMM adds the residual in place (residual == C, fp32)");
  if (g.batch > 1)
    W2VS_REQUIRE(g.residual == nullptr && g.c_batch_stride * (g.dtype_c == W2VS_F32 ? 4 : 2) % 16 == 0 &&
                     g.c_batch_stride >= g.N && g.c_batch_stride <= g.ldc && g.w_batch_rows >= g.N &&
                     (int64_t)g.batch * g.a_batch_rows < (1ll << 31),
                 "batched GEMM: no residual, products are 16-byte aligned column blocks of C");
  if (g.M <= 0) return W2VS_OK;
  return g.dtype_c == W2VS_F32 ? launch_typed<float>(g, st) : launch_typed<bf16>(g, st);
}

w2vs_status_t debug_read_tc2_fault(int* out) {
  int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc2_fault, sizeof(int));
  if (e != cudaSuccess) { set_error("read g_tc2_fault: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  *out = v;
  return W2VS_OK;
}

}  // namespace w2vs
