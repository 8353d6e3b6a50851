// One decision step of the incremental encoder as ONE persistent cooperative kernel (bf16 models).
//
// What it replaces: the chain embed -> 24 x [LN, QKV, attention, out_proj, LN, fc1, fc2] -> final LN of stream.cu's
// block_step (the reference computes the same rows by re-encoding the whole prefix, rain/simul/transducer_agent.py:
// 138-167; layer arithmetic wav2vec2.py:921-978, attention modules/multihead_attention.py:162-194).  A step is at
// most 32 tokens, so its arithmetic is negligible; its cost as ~170 dependent launches was launch latency (7.5 us
// per link) against an HBM floor of 0.1 ms for reading the 613 MB of weights once.  Here the whole step is one
// grid of one CTA per SM:
//   * phases separated by grid barriers (one 64-bit arrival counter in global memory, self-resetting):
//       per layer  [LN + QKV -> q, K/V cache]  |  [attention over the cache, keys split over CTAs]  |
//                  [out_proj + residual]  |  [LN + fc1 + GELU]  |  [fc2 + residual],   then the final LayerNorm;
//   * every matrix product is split over the CTAs by output columns (units of 8 weight rows, CTA c owns units
//     c, c + G, ...), so that each CTA streams a fixed, contiguous share of every weight matrix;
//   * a producer cursor walks that share for ALL layers ahead of time: 8 x KC slabs of W go global -> shared with
//     cp.async.bulk into a ring of stages (full/empty mbarriers).  Weights do not depend on activations, so the HBM
//     stream never stops at a phase boundary.  The cursor belongs to thread 0 and is advanced without blocking
//     ("pump") wherever that thread has nothing better to do: while it spins in a grid barrier, after every stage
//     its warp has consumed, between attention items.  (A dedicated ninth producer warp would cap the kernel at
//     168 registers per thread -- warps are allocated in fours -- and the product phases want more.);
//   * 8 consumer warps split K eight ways (mma.sync.m16n8k16, bf16, fp32 accumulate; the k index inside a 16-step
//     is permuted identically for A and W so that both are read with 8-byte accesses), partial sums are reduced
//     through shared memory in a fixed order (deterministic), then bias / GELU / residual;
//   * LayerNorm is computed by every CTA for all rows (fp32 statistics from the fp32 residual stream, bf16 operand
//     rows in shared memory); post-LN models keep the un-normalised sum in the residual buffer and apply the
//     pending LayerNorm when the value is next read (operand rows, residual term, final output);
//   * attention: one (stream, head, key split) item per group of 4 warps, K/V tiles from the cache via cp.async,
//     flash-style mma.sync, split results merged by the group that finishes a (stream, head) last.
// Numerically the step matches the multi-kernel path up to summation order (same bf16 roundings: operand rows,
// q/k/v, P, context, FFN hidden; fp32 residual stream and statistics).
#include <cuda.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"
#include "layout.h"

namespace w2vs {
__device__ int g_fused_fault = 0;   // set when a barrier / pipeline wait timed out (diagnostics)
// phase timestamps (globaltimer ns) of the last launch, written by thread 0 of the first and the last CTA:
// [cta 0 | cta G-1][layer][event]; events: 0 layer start, 1 QKV done, 2 barrier, 3 attention done, 4 barrier,
// 5 out_proj done, 6 barrier, 7 fc1 done, 8 barrier, 9 fc2 done, 10 barrier, 11 clock64 at layer start.  Read by
// tools/stream_trace.py through w2vs_debug_fused_trace (one store per phase: free).
__device__ unsigned long long g_fused_trace[2][64][12];
}
#define W2VS_TC_FAULT_FLAG (&::w2vs::g_fused_fault)
#include "tc_common.cuh"

namespace w2vs {
namespace {
using namespace tc;

constexpr int FS_CW = 8;                        // warps: all of them consume; thread 0 also drives the copies
constexpr int FS_THREADS = 32 * FS_CW;
constexpr int FS_ROWS = 32;                     // token rows of a step (two m16 tiles)
constexpr int FS_UMAX = 4;                      // 8-column units of one product a CTA may own
constexpr int FS_MAX_STAGES = 12;
constexpr int FS_SCHED = 32;                    // slabs of one layer a CTA may own
constexpr int FS_KT = 64;                       // attention key tile
constexpr int FS_ATT_GROUP_BYTES = 5 * 8192;    // Q + 2 K + 2 V tiles of 64 x 64 bf16
constexpr int FS_MAX_SPLITS = 16;
constexpr int FS_SMEM_LIMIT = 232448;
constexpr unsigned long long FS_TIMEOUT_NS = 2000000000ull;

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// non-blocking probe (mbarrier.try_wait may suspend the thread for a system-dependent time when the phase is not
// complete -- measured ~1-2 us here -- which is the last thing the producer cursor should do)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// ---- attention tile helpers (64 x 64 bf16 tiles, 16-byte chunks XOR-swizzled by the row) ----
__device__ __forceinline__ int sw(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

struct FusedArgs {
  const uint8_t* W;                    // packed weights
  unsigned long long wqkv, bqkv, wo, bo, ln1_w, ln1_b, w1, b1, w2s, b2, ln2_w, ln2_b;   // layer 0 offsets (bytes)
  unsigned long long layer_stride, enc_ln_w, enc_ln_b, sin_table;
  int layers, D, F, H, pre_ln;
  int B, ntok, n_main, f0;             // tokens per stream in this step, frames emitted, first frame index
  const float* feats; long long feat_rows;     // projected frames [B][feat_rows][D] fp32
  float* R;                            // residual stream [B * ntok][D] fp32
  bf16* q; bf16* ctx; bf16* h;         // [B * ntok][D], [.][D], FFN hidden chunk-major [F / D][32][D]
  bf16* kv; long long kv_layer_elems, kv_rows;   // cache [layers][B][kv_rows][2D]
  float* partials; unsigned* counters; // attention split states [B*H][splits][32][66]; zeroed counters [B*H]
  bf16* out;                           // [n_main][B][D]
  unsigned long long* bar;             // [0] arrivals, [1] departures (both zero between launches)
  int n_stages, n_splits;
  float scale_log2;
};

// Shared memory map: byte offsets from a 1024-aligned base.
//   ring   weight slabs, n_stages x (8 rows x D bf16)
//   abuf   the A operand of the running product: 32 rows x D bf16 (written by the LayerNorm or by a bulk copy)
//   part   per-warp partial sums [8 warps][4 units][32 rows][8] fp32          (abuf + part: attention tiles)
//   lnp    LayerNorm parameters of the two LN phases: 2 x [gamma D | beta D] fp32 (bulk copies, a layer ahead)
//   sched  this CTA's slab list of one layer (byte offsets from the layer's weights), stats [32][2], barriers, flags
struct Smem { uint32_t ring, abuf, part, lnp, sched, stats, bars, flags; int stage_bytes; };
// barriers: full[12] @0, empty[12] @96, A operand buffers @192 / @216, LayerNorm parameters [2] @200 / @208

__device__ __forceinline__ const float* lw(const FusedArgs& a, int layer, unsigned long long off) {
  return reinterpret_cast<const float*>(a.W + (size_t)layer * a.layer_stride + off);
}

// Consumer and producer cursors of the weight ring.  The producer side lives in thread 0: slabs are requested in
// consumption order (layer, product, unit, K chunk) from the per-CTA list `sched`, never blocking.
// One out-of-line copy of the mbarrier wait: its time-out logic is ~40 instructions per inlined site, and the code
// size of this kernel is what bounds its speed (see the layer loop).
#ifndef W2VS_FS_SPIN_TEST
#define W2VS_FS_SPIN_TEST 1
#endif
__device__ __noinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
#if W2VS_FS_SPIN_TEST
  // spin on the non-blocking probe: mbarrier.try_wait parks the thread for a hardware-chosen time when the phase is
  // not complete yet, and every wait in this kernel sits on the critical path of a phase
  if (mbar_test(bar, parity)) return true;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (!mbar_test(bar, parity)) {
    if ((++spins & 0xfff) == 0 && global_ns() - t0 > FS_TIMEOUT_NS) { atomicExch(&g_fused_fault, 1); return false; }
  }
  return true;
#else
  return mbar_wait(bar, parity);
#endif
}
struct Ring { int stage; uint32_t phase; };
struct Prod { int e, l, n_entries, stage; uint32_t phase; };

__device__ __forceinline__ void pump(const FusedArgs& a, const Smem& sm, uint32_t sb, const uint8_t* smem_gen, Prod& pr,
                                     int budget) {
  const unsigned long long* sched = reinterpret_cast<const unsigned long long*>(smem_gen + sm.sched);
  while (budget-- > 0 && pr.l < a.layers) {
    if (!mbar_test(sb + sm.bars + 96 + 8 * pr.stage, pr.phase ^ 1)) return;    // not yet released by all 8 warps
    const uint32_t full = sb + sm.bars + 8 * pr.stage;
    mbar_expect_tx(full, (uint32_t)sm.stage_bytes);
    // every slab is 8 rows x D bf16, contiguous in global memory (8 whole rows of W, or the slab-ordered fc2 copy):
    // ONE bulk copy -- issuing a copy costs the issuing thread ~80 ns, whatever its size
    bulk_g2s(sb + sm.ring + (uint32_t)pr.stage * sm.stage_bytes, a.W + (size_t)pr.l * a.layer_stride + sched[pr.e],
             (uint32_t)sm.stage_bytes, full);
    if (++pr.stage == a.n_stages) { pr.stage = 0; pr.phase ^= 1; }
    if (++pr.e == pr.n_entries) { pr.e = 0; ++pr.l; }
  }
}

// Grid-wide barrier for the 256 threads of every CTA.  Arrival counter in global memory, monotonically increasing
// inside a launch (target = barriers so far x CTAs); the last CTA to leave the kernel resets it.  Thread 0 uses the
// time between its arrival and the release to refill the weight ring.
__device__ __noinline__ bool grid_barrier(const FusedArgs& a, const Smem& sm, uint32_t sb, const uint8_t* smem_gen,
                                             Prod& pr, unsigned long long target, volatile int* s_abort) {
  named_bar(1, 32 * FS_CW);
  if (threadIdx.x == 0) {
    // release at gpu scope: the CTA's writes of this phase (ordered before this thread by the bar.sync above) are
    // visible to whoever observes the arrival
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(a.bar), "l"(1ull) : "memory");
    unsigned long long v;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.bar) : "memory");
      if (v >= target) break;
      pump(a, sm, sb, smem_gen, pr, 1);       // the wait is the producer's time: one slab per probe
      if ((++spins & 0x3ff) == 0) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > FS_TIMEOUT_NS) { atomicExch(&g_fused_fault, 1); *s_abort = 1; break; }
      }
    }
    pump(a, sm, sb, smem_gen, pr, FS_MAX_STAGES);     // whatever is still free: the next phase's slabs must be under way
  }
  named_bar(1, 32 * FS_CW);
  return *s_abort == 0;
}

// mean / rstd of one fp32 row held as float4 v[8] per lane
__device__ __forceinline__ void row_stats(const float4 (&v)[8], int nv, int D, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) if (j < nv) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
    }
  rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
}

// ---- LN phase prologue: operand rows A = bf16(LN(R)) into abuf, row statistics into `stats` ----
// Warp w normalises four rows; the next row's loads are in flight while a row is normalised.  gamma / beta come from
// shared memory (`lnp`, bulk-copied a layer ahead): as global loads they were sixteen dependent L2 round trips per
// row.  Which rows a warp takes is rotated by the CTA index, so that the CTAs of the grid, which all read the same
// 96 KB, do not ask the same L2 lines at the same instant.
__device__ __forceinline__ void ln_rows_to_smem(const FusedArgs& a, const Smem& sm, uint8_t* smem_gen, int warp, int lane,
                                                const float* lnp) {
  const int Mt = a.B * a.ntok, D = a.D, nv = D >> 7;
  float* stats = reinterpret_cast<float*>(smem_gen + sm.stats);
  constexpr int RPW = FS_ROWS / FS_CW;
  const int cta = blockIdx.x;
  const int wrot = (warp + cta) & (FS_CW - 1), irot = (cta >> 3) & (RPW - 1);
  float4 vn[8];
  {
    const int row = wrot + FS_CW * irot;
    if (row < Mt) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nv) vn[j] = __ldcg(reinterpret_cast<const float4*>(a.R + (size_t)row * D + 4 * lane + 128 * j));
    }
  }
#pragma unroll 1
  for (int it = 0; it < RPW; ++it) {
    const int row = wrot + FS_CW * ((it + irot) & (RPW - 1));
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = vn[j];
    if (it + 1 < RPW) {
      const int rown = wrot + FS_CW * ((it + 1 + irot) & (RPW - 1));
      if (rown < Mt) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nv) vn[j] = __ldcg(reinterpret_cast<const float4*>(a.R + (size_t)rown * D + 4 * lane + 128 * j));
      }
    }
    uint8_t* dst = smem_gen + sm.abuf + (size_t)row * (2 * D);
    if (row < Mt) {
      float mean, rstd;
      row_stats(v, nv, D, mean, rstd);
      if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) {
          const float4 gm = *reinterpret_cast<const float4*>(lnp + 4 * lane + 128 * j);
          const float4 bt = *reinterpret_cast<const float4*>(lnp + D + 4 * lane + 128 * j);
          uint2 u;
          u.x = pack_bf16x2((v[j].x - mean) * rstd * gm.x + bt.x, (v[j].y - mean) * rstd * gm.y + bt.y);
          u.y = pack_bf16x2((v[j].z - mean) * rstd * gm.z + bt.z, (v[j].w - mean) * rstd * gm.w + bt.w);
          *reinterpret_cast<uint2*>(dst + (4 * lane + 128 * j) * 2) = u;
        }
    } else {
      for (int j = 0; j < nv; ++j) *reinterpret_cast<uint2*>(dst + (4 * lane + 128 * j) * 2) = make_uint2(0u, 0u);
    }
  }
}

enum { EPI_QKV = 0, EPI_RESID = 1, EPI_GELU = 2 };

// One matrix product of a phase for this CTA's units.  The A operand sits in abuf (32 rows x D bf16): written by the
// LayerNorm (a_glob == nullptr), or bulk-copied from global memory by thread 0 -- chunk 0 was requested by the caller
// right after the grid barrier, chunk j + 1 of the long-K product (fc2; the FFN hidden is stored chunk-major,
// [K / D][32][D], so that every chunk is one dense block) as soon as all warps hold chunk j in registers.  One
// instance serves all four products (rolled loops over units and K chunks: code size bounds this kernel).
template <int MT>
__device__ __forceinline__ bool gemm_phase(const FusedArgs& a, const Smem& sm, uint32_t sb, uint8_t* smem_gen, Ring& rs,
                                           Prod& pr, uint32_t& aphase, int layer, int N, int K, const bf16* a_glob, int epi,
                                           const float* bias, const float* res_lnp, int a_bufs, int warp, int lane) {
  const int G = gridDim.x, cta = blockIdx.x;
  const int D = a.D, KC = D, Mt = a.B * a.ntok;
  const int kw = KC / FS_CW, nsteps = kw >> 4, nch = K / KC;
  const int g = lane >> 2, q = lane & 3;
  const int units_total = N >> 3;
  const int n_units = units_total > cta ? (units_total - 1 - cta) / G + 1 : 0;
  float* part = reinterpret_cast<float*>(smem_gen + sm.part);
  const float* stats = reinterpret_cast<const float*>(smem_gen + sm.stats);
  const uint32_t abar = sb + sm.bars + 192;        // A operand arrival: buffer 0; buffer 1 at + 24
  const uint32_t a_chunk_bytes = (uint32_t)Mt * KC * 2;
  const int tid = threadIdx.x;
  if (tid == 0 && a_glob != nullptr && n_units > 0) {
    // the A operand comes from global memory (attention context / FFN hidden): bulk copies into abuf -- every warp
    // of this CTA is past its last use of that memory (they all arrived at the grid barrier before this phase).
    // Long-K product: the first a_bufs chunks now, chunk j + a_bufs once every warp holds chunk j in registers.
    fence_async_smem();
    for (int j = 0; j < a_bufs && j < nch; ++j) {
      mbar_expect_tx(abar + 24 * j, a_chunk_bytes);
      bulk_g2s(sb + sm.abuf + j * a_chunk_bytes, a_glob + (size_t)j * FS_ROWS * KC, a_chunk_bytes, abar + 24 * j);
    }
  }

  // the epilogue's residual and bias values do not depend on the product: request them first
  const bool epi_thread = tid < n_units * Mt;
  const int ei = epi_thread ? tid / Mt : 0, er = epi_thread ? tid - ei * Mt : 0;
  const int en0 = (cta + ei * G) * 8;
  float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, bb0 = o0, bb1 = o0;
  if (epi_thread) {
    bb0 = *reinterpret_cast<const float4*>(bias + en0);
    bb1 = *reinterpret_cast<const float4*>(bias + en0 + 4);
    if (epi == EPI_RESID) {
      const float* xr = a.R + (size_t)er * D + en0;
      o0 = __ldcg(reinterpret_cast<const float4*>(xr));
      o1 = __ldcg(reinterpret_cast<const float4*>(xr + 4));
    }
  }

  uint2 af[MT][2][8];       // A fragments of one K chunk: [m tile][row half][k16 step] x 8 bytes
  bool ok = true, need_a = true;
#pragma unroll 1
  for (int i = 0; i < n_units; ++i) {
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
#pragma unroll 1
    for (int j = 0; j < nch; ++j) {
      if (need_a) {
        need_a = nch > 1;      // K == D: the same fragments serve every unit
        const int ab = a_glob != nullptr && a_bufs == 2 ? (j & 1) : 0;      // which A buffer holds chunk j
        if (a_glob != nullptr) { ok = wait_bar(abar + 24 * ab, (aphase >> ab) & 1u) && ok; aphase ^= 1u << ab; }
        const uint32_t base = sb + sm.abuf + (a_glob != nullptr ? ab * a_chunk_bytes : 0u) + (uint32_t)g * (2 * D) +
                              (uint32_t)(warp * kw + 4 * q) * 2;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int s = 0; s < 8; ++s)
              if (s < nsteps) af[mt][hf][s] = lds64(base + (uint32_t)(mt * 16 + 8 * hf) * (2 * D) + s * 32);
        if (nch > 1) {
          // every warp holds chunk j in registers: thread 0 may refill that buffer with chunk j + a_bufs
          named_bar(1, 32 * FS_CW);
          if (tid == 0 && j + a_bufs < nch) {
            mbar_expect_tx(abar + 24 * ab, a_chunk_bytes);
            bulk_g2s(sb + sm.abuf + ab * a_chunk_bytes, a_glob + (size_t)(j + a_bufs) * FS_ROWS * KC, a_chunk_bytes, abar + 24 * ab);
          }
        }
      }
      ok = wait_bar(sb + sm.bars + 8 * rs.stage, rs.phase) && ok;
      const uint32_t wrow = sb + sm.ring + (uint32_t)rs.stage * sm.stage_bytes + (uint32_t)g * (2 * D) +
                            (uint32_t)(warp * kw + 4 * q) * 2;
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < nsteps) {
          const uint2 bw = lds64(wrow + s * 32);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            mma_16816(acc[mt], af[mt][0][s].x, af[mt][1][s].x, af[mt][0][s].y, af[mt][1][s].y, bw.x, bw.y);
        }
      // the MMAs have consumed the registers, so every lane's loads from the stage have completed: hand the
      // stage back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(sb + sm.bars + 96 + 8 * rs.stage);
      if (++rs.stage == a.n_stages) { rs.stage = 0; rs.phase ^= 1; }
      if (tid == 0) pump(a, sm, sb, smem_gen, pr, 1);       // usually refills the stage released one step earlier
    }
    // partial sums of this warp's K slice -> shared memory
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float* p0 = part + (((size_t)warp * FS_UMAX + i) * FS_ROWS + mt * 16 + g) * 8 + 2 * q;
      *reinterpret_cast<float2*>(p0) = make_float2(acc[mt][0], acc[mt][1]);
      *reinterpret_cast<float2*>(p0 + 64) = make_float2(acc[mt][2], acc[mt][3]);
    }
  }
  named_bar(1, 32 * FS_CW);
  // ---- fixed-order reduction + epilogue: one thread per (unit, row), 8 consecutive columns
  if (epi_thread) {
    const int i = ei, r = er, n0 = en0;
    float v[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
    for (int w = 0; w < FS_CW; ++w) {
      const float* p = part + (((size_t)w * FS_UMAX + i) * FS_ROWS + r) * 8;
      const float4 x0 = *reinterpret_cast<const float4*>(p), x1 = *reinterpret_cast<const float4*>(p + 4);
      v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w; v[4] += x1.x; v[5] += x1.y; v[6] += x1.z; v[7] += x1.w;
    }
    if (epi == EPI_QKV) {
      bf16* dst;
      if (n0 < D) dst = a.q + (size_t)r * D + n0;
      else {
        const int b = r / a.ntok, t = r - b * a.ntok;
        dst = a.kv + (size_t)layer * a.kv_layer_elems + ((size_t)b * a.kv_rows + a.f0 + t) * (2 * D) + (n0 - D);
      }
      store8(dst, v);
    } else if (epi == EPI_GELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = gelu_tanh(v[e]);
      const int jc = n0 / KC;
      store8(a.h + ((size_t)jc * FS_ROWS + r) * KC + (n0 - jc * KC), v);
    } else {
      float o[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      if (res_lnp != nullptr) {     // post-LN: the residual term is the pending LayerNorm of the stored sum
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float gg[8], bt[8];
        load8(res_lnp + n0, gg);
        load8(res_lnp + D + n0, bt);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (o[e] - mean) * rstd * gg[e] + bt[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += o[e];
      store8(a.R + (size_t)r * D + n0, v);
    }
  }
  return ok;
}

// ---- attention over the K/V cache: items (stream, head, key split) on groups of 4 warps ----
__device__ __forceinline__ void load_tile(bf16* tile, const bf16* src, long long row_stride, int first_row, int n_rows,
                                          int gtid) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = gtid + 128 * r;
    const int row = c >> 3, chunk = c & 7;
    const bool ok = row < n_rows;
    const bf16* gp = src + (size_t)(first_row + (ok ? row : 0)) * row_stride + chunk * 8;
    cp_async16(smem_u32(tile + sw(row, chunk)), gp, ok);
  }
}

__device__ __forceinline__ void attention_phase(const FusedArgs& a, const Smem& sm, uint8_t* smem_gen, int layer,
                                                int warp, int lane) {
  const int G = gridDim.x, cta = blockIdx.x;
  const int gi = warp >> 2, wi = warp & 3, gtid = threadIdx.x & 127;
  const int mt = wi & 1, par = wi >> 1;
  const int g = lane >> 2, t4 = lane & 3;
  const int D = a.D, H = a.H, S = a.n_splits, ntok = a.ntok;
  const int n_keys = a.f0 + ntok, n_kt = (n_keys + FS_KT - 1) / FS_KT, per = (n_kt + S - 1) / S;
  const long long krs = 2ll * D;
  bf16* Qs = reinterpret_cast<bf16*>(smem_gen + sm.abuf + (size_t)gi * FS_ATT_GROUP_BYTES);
  bf16* Ks = Qs + 4096;          // [2][64 x 64]
  bf16* Vs = Ks + 2 * 4096;      // [2][64 x 64]
  const bf16* kv_l = a.kv + (size_t)layer * a.kv_layer_elems;
  const float sl2 = a.scale_log2;
  const bool active = mt * 16 < ntok;      // this warp's 16 query rows exist
  for (int item = cta * 2 + gi; item < a.B * H * S; item += 2 * G) {
    const int sp = item % S, bh = item / S, h = bh % H, b = bh / H;
    const int t_begin = min(sp * per, n_kt), t_end = min(t_begin + per, n_kt);
    const bf16* qbase = a.q + (size_t)b * ntok * D + (size_t)h * 64;
    const bf16* kbase = kv_l + (size_t)b * a.kv_rows * krs + (size_t)h * 64;
    const bf16* vbase = kbase + D;
    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int t0 = t_begin; t0 < t_end; t0 += 2) {
      if (t0 == t_begin) load_tile(Qs, qbase, D, 0, ntok, gtid);
#pragma unroll
      for (int p = 0; p < 2; ++p)
        if (t0 + p < t_end) {
          const int k0 = (t0 + p) * FS_KT, cnt = min(FS_KT, n_keys - k0);
          load_tile(Ks + p * 4096, kbase, krs, k0, cnt, gtid);
          load_tile(Vs + p * 4096, vbase, krs, k0, cnt, gtid);
        }
      cp_async_commit();
      cp_async_wait_all();
      named_bar(2 + gi, 128);
      if (t0 == t_begin) {
        const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ldsm_x4(smem_u32(Qs + sw(row, kk * 2 + (lane >> 4))), qf[kk]);
      }
      if (active && t0 + par < t_end) {
        const bf16* Kt = Ks + par * 4096;
        const bf16* Vt = Vs + par * 4096;
        const int cnt = min(FS_KT, n_keys - (t0 + par) * FS_KT);
        const int mi = lane >> 3;
        // 16 keys at a time in a rolled loop (S = Q K^T for 16 keys, online softmax, O += P V): a fraction of the
        // code of the fully unrolled 64-key step, which this kernel cannot afford
#pragma unroll 1
        for (int sub = 0; sub * 16 < cnt; ++sub) {
          float s[2][4];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) { s[jj][0] = s[jj][1] = s[jj][2] = s[jj][3] = 0.f; }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t kf[4];
            ldsm_x4(smem_u32(Kt + sw(sub * 16 + (lane & 7) + (mi >> 1) * 8, kk * 2 + (mi & 1))), kf);
            mma_16816(s[0], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], kf[0], kf[1]);
            mma_16816(s[1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], kf[2], kf[3]);
          }
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const bool vis = sub * 16 + jj * 8 + 2 * t4 + e < cnt;      // keys past the end of the cache (tile tail)
              s[jj][e] = vis ? s[jj][e] : -INFINITY;
              s[jj][2 + e] = vis ? s[jj][2 + e] : -INFINITY;
              mx0 = fmaxf(mx0, s[jj][e]);
              mx1 = fmaxf(mx1, s[jj][2 + e]);
            }
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
          const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
          const float ms0 = mn0 == -INFINITY ? 0.f : mn0 * sl2;
          const float ms1 = mn1 == -INFINITY ? 0.f : mn1 * sl2;
          const float a0 = exp2f(m0 * sl2 - ms0), a1 = exp2f(m1 * sl2 - ms1);
          m0 = mn0; m1 = mn1;
          uint32_t pf[4];
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const float p0 = exp2f(fmaf(s[jj][0], sl2, -ms0)), p1 = exp2f(fmaf(s[jj][1], sl2, -ms0));
            const float p2 = exp2f(fmaf(s[jj][2], sl2, -ms1)), p3 = exp2f(fmaf(s[jj][3], sl2, -ms1));
            sum0 += p0 + p1;
            sum1 += p2 + p3;
            pf[2 * jj] = pack_bf16x2(p0, p1);          // row g,     keys 8 jj + 2 t4 ..
            pf[2 * jj + 1] = pack_bf16x2(p2, p3);      // row g + 8
          }
          l0 = l0 * a0 + sum0;
          l1 = l1 * a1 + sum1;
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }
#pragma unroll
          for (int jp = 0; jp < 4; ++jp) {
            uint32_t vf[4];
            ldsm_x4_trans(smem_u32(Vt + sw(sub * 16 + (lane & 7) + (mi & 1) * 8, jp * 2 + (mi >> 1))), vf);
            mma_16816(o[2 * jp], pf[0], pf[1], pf[2], pf[3], vf[0], vf[1]);
            mma_16816(o[2 * jp + 1], pf[0], pf[1], pf[2], pf[3], vf[2], vf[3]);
          }
        }
      }
      named_bar(2 + gi, 128);       // both tiles consumed before the next pass (or the merge scratch) overwrites them
    }
    // ---- row sums across the quad, then the odd-tile warps hand their state to the even-tile warps
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float* scr = reinterpret_cast<float*>(Ks) + (size_t)mt * 16 * 66;     // [2 m-tiles][16 rows][66]
    if (par == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<float2*>(scr + g * 66 + j * 8 + 2 * t4) = make_float2(o[j][0], o[j][1]);
        *reinterpret_cast<float2*>(scr + (g + 8) * 66 + j * 8 + 2 * t4) = make_float2(o[j][2], o[j][3]);
      }
      if (t4 == 0) {
        *reinterpret_cast<float2*>(scr + g * 66 + 64) = make_float2(m0, l0);
        *reinterpret_cast<float2*>(scr + (g + 8) * 66 + 64) = make_float2(m1, l1);
      }
    }
    named_bar(2 + gi, 128);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    if (par == 0) {
      const float2 ml0 = *reinterpret_cast<const float2*>(scr + g * 66 + 64);
      const float2 ml1 = *reinterpret_cast<const float2*>(scr + (g + 8) * 66 + 64);
      const float M0 = fmaxf(m0, ml0.x), M1 = fmaxf(m1, ml1.x);
      const float wa0 = m0 == -INFINITY ? 0.f : exp2f((m0 - M0) * sl2), wb0 = ml0.x == -INFINITY ? 0.f : exp2f((ml0.x - M0) * sl2);
      const float wa1 = m1 == -INFINITY ? 0.f : exp2f((m1 - M1) * sl2), wb1 = ml1.x == -INFINITY ? 0.f : exp2f((ml1.x - M1) * sl2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 x0 = *reinterpret_cast<const float2*>(scr + g * 66 + j * 8 + 2 * t4);
        const float2 x1 = *reinterpret_cast<const float2*>(scr + (g + 8) * 66 + j * 8 + 2 * t4);
        o[j][0] = o[j][0] * wa0 + x0.x * wb0; o[j][1] = o[j][1] * wa0 + x0.y * wb0;
        o[j][2] = o[j][2] * wa1 + x1.x * wb1; o[j][3] = o[j][3] * wa1 + x1.y * wb1;
      }
      l0 = l0 * wa0 + ml0.y * wb0; l1 = l1 * wa1 + ml1.y * wb1;
      m0 = M0; m1 = M1;
      if (S > 1) {
        float* P = a.partials + ((size_t)bh * S + sp) * (size_t)(FS_ROWS * 66);
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
      } else {
        const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
        bf16* c0 = a.ctx + ((size_t)b * ntok + r0) * D + (size_t)h * 64 + 2 * t4;
        bf16* c1 = a.ctx + ((size_t)b * ntok + r1) * D + (size_t)h * 64 + 2 * t4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (r0 < ntok) *reinterpret_cast<uint32_t*>(c0 + j * 8) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
          if (r1 < ntok) *reinterpret_cast<uint32_t*>(c1 + j * 8) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
        }
      }
    }
    if (S > 1) {
      // every split of this (stream, head) has to be in global memory before anyone folds them: a head-level
      // barrier on a counter that only grows inside a launch (target = splits x layers so far), then the S groups
      // share the fold -- group `sp` takes the (row, 8-column chunk) pairs congruent to sp.  (Letting the group that
      // arrives last fold all pairs alone put 5 us on the critical path of every layer.)  The groups of a head sit
      // on consecutive work-item slots, all of them running, so the wait cannot deadlock.
      named_bar(2 + gi, 128);
      if (gtid == 0) {
        atomicAdd(a.counters + bh, 1u);
        const unsigned target = (unsigned)S * (unsigned)(layer + 1);
        unsigned v, spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.counters + bh) : "memory");
          if (v >= target) break;
          if ((++spins & 0x3ff) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > FS_TIMEOUT_NS) { atomicExch(&g_fused_fault, 1); break; }
          }
        }
      }
      named_bar(2 + gi, 128);
      const float* P0 = a.partials + (size_t)bh * S * (size_t)(FS_ROWS * 66);
      for (int idx = sp + S * gtid; idx < ntok * 8; idx += S * 128) {
        const int row = idx >> 3, chunk = idx & 7;
        float mrun = -INFINITY, lsum = 0.f;
        float accv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int s0 = 0; s0 < S; s0 += 8) {       // batches of 8 splits: all loads of a batch in flight together
          float2 ml[8];
          float4 ov[8][2];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (s0 + k < S) {
              const float* pr2 = P0 + (size_t)(s0 + k) * FS_ROWS * 66 + row * 66;       // 8-byte aligned rows
              ml[k] = __ldcg(reinterpret_cast<const float2*>(pr2 + 64));
              const float2 x0 = __ldcg(reinterpret_cast<const float2*>(pr2 + chunk * 8)), x1 = __ldcg(reinterpret_cast<const float2*>(pr2 + chunk * 8 + 2));
              const float2 x2 = __ldcg(reinterpret_cast<const float2*>(pr2 + chunk * 8 + 4)), x3 = __ldcg(reinterpret_cast<const float2*>(pr2 + chunk * 8 + 6));
              ov[k][0] = make_float4(x0.x, x0.y, x1.x, x1.y);
              ov[k][1] = make_float4(x2.x, x2.y, x3.x, x3.y);
            }
          float mb = mrun;
#pragma unroll
          for (int k = 0; k < 8; ++k) if (s0 + k < S) mb = fmaxf(mb, ml[k].x);
          if (mb == -INFINITY) continue;
          const float resc = mrun == -INFINITY ? 0.f : exp2f((mrun - mb) * sl2);
          lsum *= resc;
#pragma unroll
          for (int e = 0; e < 8; ++e) accv[e] *= resc;
          mrun = mb;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (s0 + k < S && ml[k].x != -INFINITY) {
              const float w = exp2f((ml[k].x - mb) * sl2);
              lsum = fmaf(ml[k].y, w, lsum);
              accv[0] = fmaf(ov[k][0].x, w, accv[0]); accv[1] = fmaf(ov[k][0].y, w, accv[1]);
              accv[2] = fmaf(ov[k][0].z, w, accv[2]); accv[3] = fmaf(ov[k][0].w, w, accv[3]);
              accv[4] = fmaf(ov[k][1].x, w, accv[4]); accv[5] = fmaf(ov[k][1].y, w, accv[5]);
              accv[6] = fmaf(ov[k][1].z, w, accv[6]); accv[7] = fmaf(ov[k][1].w, w, accv[7]);
            }
        }
        const float inv = lsum > 0.f ? 1.0f / lsum : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) accv[e] *= inv;
        store8(a.ctx + ((size_t)b * ntok + row) * D + (size_t)h * 64 + chunk * 8, accv);
      }
      named_bar(2 + gi, 128);     // the scratch tiles are reused by the group's next item
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(FS_THREADS, 1)
stream_fused_kernel(const __grid_constant__ FusedArgs a, const __grid_constant__ Smem sm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;      // shared-window address of the map's base
  uint8_t* smem_gen = smem_raw + (sb - smem_u32(smem_raw));       // the same place as a generic pointer
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  volatile int* s_abort = reinterpret_cast<volatile int*>(smem_gen + sm.flags);
  const int D = a.D, Mt = a.B * a.ntok, nv = D >> 7;
  const uint32_t lnbar = sb + sm.bars + 200;
  const uint32_t ln_bytes = 8u * D;      // gamma | beta, fp32

  Prod pr{0, 0, 0, 0, 0u};
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.n_stages; ++i) { mbar_init(sb + sm.bars + 8 * i, 1); mbar_init(sb + sm.bars + 96 + 8 * i, FS_CW); }
    mbar_init(sb + sm.bars + 192, 1);
    mbar_init(sb + sm.bars + 216, 1);
    mbar_init(lnbar, 1);
    mbar_init(lnbar + 8, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // this CTA's slabs of one layer, in consumption order: product, unit, K chunk
    unsigned long long* sched = reinterpret_cast<unsigned long long*>(smem_gen + sm.sched);
    const int nchunk = a.F / D;
    int n = 0;
    for (int p = 0; p < 4; ++p) {
      const unsigned long long off = p == 0 ? a.wqkv : (p == 1 ? a.wo : (p == 2 ? a.w1 : a.w2s));
      const int units = (p == 0 ? 3 * D : (p == 2 ? a.F : D)) >> 3, nch = p == 3 ? nchunk : 1;
      for (int u = cta; u < units; u += G)
        for (int j = 0; j < nch; ++j) sched[n++] = off + ((unsigned long long)u * nch + j) * sm.stage_bytes;
    }
    pr.n_entries = n;
    if (n == 0) pr.l = a.layers;
    fence_async_smem();
    pump(a, sm, sb, smem_gen, pr, FS_MAX_STAGES);       // the weight stream starts before anything else
    // LayerNorm parameters of layer 0's two LN phases (gamma and beta are adjacent in the packed weights)
    mbar_expect_tx(lnbar, ln_bytes);
    bulk_g2s(sb + sm.lnp, a.pre_ln ? (const void*)lw(a, 0, a.ln1_w) : (const void*)(a.W + a.enc_ln_w), ln_bytes, lnbar);
    mbar_expect_tx(lnbar + 8, ln_bytes);
    bulk_g2s(sb + sm.lnp + ln_bytes, a.pre_ln ? lw(a, 0, a.ln2_w) : lw(a, 0, a.ln1_w), ln_bytes, lnbar + 8);
  }
  __syncthreads();

  unsigned long long nbar = 0;
  Ring rs{0, 0};
  uint32_t aphase = 0;
  // two A-operand buffers for the long-K product when two chunks of Mt rows fit into abuf + part (part is written
  // only after the last chunk has been read)
  const int a_bufs = 2u * Mt * D * 2u <= sm.lnp - sm.abuf ? 2 : 1;
  bool ok = true;
  const int tr = threadIdx.x == 0 ? (cta == 0 ? 0 : (cta == G - 1 ? 1 : -1)) : -1;
#define FS_TRACE(l_, ev_) do { if (tr >= 0 && (l_) < 64) g_fused_trace[tr][l_][ev_] = global_ns(); } while (0)
  // ---- embed: R = projected frame + sinusoidal position (absolute index frame + 2), one warp per row
  for (int row = cta * FS_CW + warp; row < Mt; row += G * FS_CW) {
    const int b = row / a.ntok, t = row - b * a.ntok;
    const float* fr = a.feats + ((size_t)b * a.feat_rows + a.f0 + t) * D + 4 * lane;
    const float* ps = reinterpret_cast<const float*>(a.W + a.sin_table) + (size_t)(a.f0 + t + 2) * D + 4 * lane;
    for (int j = 0; j < nv; ++j) {
      const float4 x = __ldcg(reinterpret_cast<const float4*>(fr + 128 * j));
      const float4 p = *reinterpret_cast<const float4*>(ps + 128 * j);
      *reinterpret_cast<float4*>(a.R + (size_t)row * D + 4 * lane + 128 * j) = make_float4(x.x + p.x, x.y + p.y, x.z + p.z, x.w + p.w);
    }
  }
  ok = grid_barrier(a, sm, sb, smem_gen, pr, ++nbar * G, s_abort);

#pragma unroll 1
  for (int l = 0; l < a.layers && ok; ++l) {
    // The four products of a layer run through ONE inlined copy of the LayerNorm / product / barrier code (a rolled
    // loop): every instruction of this kernel executes once per layer, so its code must stay small.
    FS_TRACE(l, 0);
    if (tr >= 0 && l < 64) g_fused_trace[tr][l][11] = (unsigned long long)clock64();
#pragma unroll 1
    for (int ph = 0; ph < 4 && ok; ++ph) {
      // ph 0: LN + QKV -> q, K/V cache | 1: out_proj + residual | 2: LN + fc1 + GELU | 3: fc2 + residual
      // LayerNorm parameters: buffer 0 serves ph 0 (and the post-LN residual of ph 1), buffer 1 ph 2 (and ph 3)
      const float* lnp = reinterpret_cast<const float*>(smem_gen + sm.lnp + (ph >> 1) * ln_bytes);
      if ((ph & 1) == 0) {
        ok = wait_bar(lnbar + 8 * (ph >> 1), l & 1) && ok;
        ln_rows_to_smem(a, sm, smem_gen, warp, lane, lnp);
        named_bar(1, 32 * FS_CW);
      }
      const int N = ph == 0 ? 3 * D : (ph == 2 ? a.F : D), K = ph == 3 ? a.F : D;
      const bf16* ag = ph == 1 ? a.ctx : (ph == 3 ? a.h : nullptr);
      const int epi = ph == 0 ? EPI_QKV : (ph == 2 ? EPI_GELU : EPI_RESID);
      const float* bias = lw(a, l, ph == 0 ? a.bqkv : (ph == 1 ? a.bo : (ph == 2 ? a.b1 : a.b2)));
      ok = gemm_phase<MT>(a, sm, sb, smem_gen, rs, pr, aphase, l, N, K, ag, epi, bias, ((ph & 1) && !a.pre_ln) ? lnp : nullptr,
                          a_bufs, warp, lane) && ok;
      FS_TRACE(l, ph == 0 ? 1 : 3 + 2 * ph);
      if (threadIdx.x == 0 && (ph & 1) && l + 1 < a.layers) {
        // this LayerNorm-parameter buffer is free now: request the next layer's set.  Pre-LN: ln1 / ln2 of layer
        // l + 1; post-LN: the LayerNorm pending on the residual sum, i.e. ln2 of layer l / ln1 of layer l + 1.
        const float* src = ph == 1 ? (a.pre_ln ? lw(a, l + 1, a.ln1_w) : lw(a, l, a.ln2_w))
                                   : (a.pre_ln ? lw(a, l + 1, a.ln2_w) : lw(a, l + 1, a.ln1_w));
        fence_async_smem();
        mbar_expect_tx(lnbar + 8 * (ph >> 1), ln_bytes);
        bulk_g2s(sb + sm.lnp + (ph >> 1) * ln_bytes, src, ln_bytes, lnbar + 8 * (ph >> 1));
      }
      ok = grid_barrier(a, sm, sb, smem_gen, pr, ++nbar * G, s_abort) && ok;
      FS_TRACE(l, ph == 0 ? 2 : 4 + 2 * ph);
      if (ph == 0 && ok) {
        // attention over the cache (this step's K/V rows were appended by the QKV epilogue)
        attention_phase(a, sm, smem_gen, l, warp, lane);
        FS_TRACE(l, 3);
        ok = grid_barrier(a, sm, sb, smem_gen, pr, ++nbar * G, s_abort);
        FS_TRACE(l, 4);
      }
    }
  }
  if (ok) {
    // ---- final LayerNorm of the emitted frames: encoder.layer_norm (pre-LN) or the pending final_layer_norm
    const float* gf = a.pre_ln ? reinterpret_cast<const float*>(a.W + a.enc_ln_w) : lw(a, a.layers - 1, a.ln2_w);
    const float* bf = a.pre_ln ? reinterpret_cast<const float*>(a.W + a.enc_ln_b) : lw(a, a.layers - 1, a.ln2_b);
    for (int o = cta * FS_CW + warp; o < a.B * a.n_main; o += G * FS_CW) {
      const int b = o / a.n_main, t = o - b * a.n_main;
      const float* xr = a.R + ((size_t)b * a.ntok + t) * D + 4 * lane;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nv) v[j] = __ldcg(reinterpret_cast<const float4*>(xr + 128 * j));
      float mean, rstd;
      row_stats(v, nv, D, mean, rstd);
      bf16* dst = a.out + ((size_t)t * a.B + b) * D + 4 * lane;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) {
          const float4 gg = *reinterpret_cast<const float4*>(gf + 4 * lane + 128 * j);
          const float4 bt = *reinterpret_cast<const float4*>(bf + 4 * lane + 128 * j);
          uint2 u;
          u.x = pack_bf16x2((v[j].x - mean) * rstd * gg.x + bt.x, (v[j].y - mean) * rstd * gg.y + bt.y);
          u.y = pack_bf16x2((v[j].z - mean) * rstd * gg.z + bt.z, (v[j].w - mean) * rstd * gg.w + bt.w);
          *reinterpret_cast<uint2*>(dst + 128 * j) = u;
        }
    }
  }
  // ---- leave: the last CTA out resets the barrier words for the next launch
  named_bar(1, 32 * FS_CW);
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.bar + 1, 1ull) == (unsigned long long)G - 1) {
      a.bar[0] = 0ull;
      a.bar[1] = 0ull;
      for (int i = 0; i < a.B * a.H; ++i) a.counters[i] = 0u;      // head-level barriers of the attention phase
      __threadfence();
    }
  }
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------
bool stream_fused_applicable(const w2vs_config* cfg, int B, int ntok) {
  const int D = cfg->embed_dim, F = cfg->ffn_dim, G = num_sms();
  const int max_units = ((3 * D > F ? 3 * D : F) / 8 + G - 1) / G;
  return cfg->dtype == W2VS_BF16 && D % 128 == 0 && D <= 1024 && F % D == 0 && cfg->heads * 64 == D && ntok >= 1 &&
         B * ntok <= FS_ROWS && max_units <= FS_UMAX && cfg->pos_type == W2VS_POS_SIN;
}

w2vs_status_t launch_stream_fused(const StreamFusedArgs& h, cudaStream_t st) {
  const w2vs_config* cfg = h.cfg;
  W2VS_REQUIRE(stream_fused_applicable(cfg, h.B, h.ntok), "fused incremental step: configuration not supported");
  const int D = cfg->embed_dim, F = cfg->ffn_dim, G = num_sms();
  Smem sm;
  sm.stage_bytes = 8 * 2 * D;               // a slab lands exactly as it lies in global memory: one bulk copy
  const int abuf_bytes = FS_ROWS * 2 * D, part_bytes = FS_CW * FS_UMAX * FS_ROWS * 8 * 4;
  const int uni_bytes = (int)align_up((size_t)(abuf_bytes + part_bytes > 2 * FS_ATT_GROUP_BYTES ? abuf_bytes + part_bytes
                                                                                                : 2 * FS_ATT_GROUP_BYTES), 1024);
  const int lnp_bytes = 2 * 8 * D;          // two [gamma | beta] fp32 sets
  const int tail_bytes = 1024;              // slab list (256) + stats (256) + barriers (256) + flags (64)
  int n_stages = (FS_SMEM_LIMIT - 1024 - uni_bytes - lnp_bytes - tail_bytes) / sm.stage_bytes;
  if (n_stages > FS_MAX_STAGES) n_stages = FS_MAX_STAGES;
  // the ring is topped up at every grid barrier, so it has to hold what one product phase of a CTA consumes
  const int u_qkv = (3 * D / 8 + G - 1) / G, u_fc1 = (F / 8 + G - 1) / G, u_fc2 = ((D / 8 + G - 1) / G) * (F / D);
  const int need = u_qkv > u_fc1 ? (u_qkv > u_fc2 ? u_qkv : u_fc2) : (u_fc1 > u_fc2 ? u_fc1 : u_fc2);
  W2VS_REQUIRE(n_stages >= need && n_stages >= 2, "fused incremental step: shared memory too small for the weight ring");
  W2VS_REQUIRE(u_qkv + (D / 8 + G - 1) / G + u_fc1 + u_fc2 <= FS_SCHED, "fused incremental step: too few SMs for this model");
  const int ring_bytes = (int)align_up((size_t)n_stages * sm.stage_bytes, 1024);
  sm.ring = 0;
  sm.abuf = ring_bytes;
  sm.part = sm.abuf + abuf_bytes;
  sm.lnp = sm.abuf + uni_bytes;
  sm.sched = sm.lnp + lnp_bytes;
  sm.stats = sm.sched + 256;
  sm.bars = sm.stats + 256;
  sm.flags = sm.bars + 256;
  const size_t smem_bytes = (size_t)ring_bytes + uni_bytes + lnp_bytes + tail_bytes + 1024;

  FusedArgs a{};
  a.W = reinterpret_cast<const uint8_t*>(h.W);
  const LayerW& l0 = h.wl->layer0;
  a.wqkv = l0.wqkv; a.bqkv = l0.bqkv; a.wo = l0.wo; a.bo = l0.bo; a.ln1_w = l0.ln1_w; a.ln1_b = l0.ln1_b;
  a.w1 = l0.w1; a.b1 = l0.b1; a.w2s = l0.w2s; a.b2 = l0.b2; a.ln2_w = l0.ln2_w; a.ln2_b = l0.ln2_b;
  a.layer_stride = h.wl->layer_stride; a.enc_ln_w = h.wl->enc_ln_w; a.enc_ln_b = h.wl->enc_ln_b; a.sin_table = h.wl->sin_table;
  a.layers = cfg->layers; a.D = D; a.F = F; a.H = cfg->heads; a.pre_ln = cfg->layer_norm_first != 0;
  a.B = h.B; a.ntok = h.ntok; a.n_main = h.n_main; a.f0 = h.f0;
  a.feats = h.feats; a.feat_rows = h.feat_rows;
  a.R = h.R; a.q = (bf16*)h.q; a.ctx = (bf16*)h.ctx; a.h = (bf16*)h.h;
  a.kv = (bf16*)h.kv; a.kv_layer_elems = h.kv_layer_elems; a.kv_rows = h.kv_rows;
  a.partials = h.partials; a.counters = h.counters; a.out = (bf16*)h.out; a.bar = h.bar;
  a.n_stages = n_stages;
  const int n_kt = (h.f0 + h.ntok + FS_KT - 1) / FS_KT;
  int splits = (2 * G) / (h.B * cfg->heads);
  if (splits > n_kt) splits = n_kt;
  if (splits > FS_MAX_SPLITS) splits = FS_MAX_SPLITS;
  if (splits > h.max_splits) splits = h.max_splits;
  if (splits < 1) splits = 1;
  a.n_splits = splits;
  a.scale_log2 = 0.125f * 1.4426950408889634f;

  const int mt = (h.B * h.ntok + 15) / 16;
  void (*kern)(FusedArgs, Smem) = mt <= 1 ? stream_fused_kernel<1> : stream_fused_kernel<2>;
  static PerDeviceOnce once[2];
  bool& done = once[mt <= 1 ? 0 : 1].here();
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM_LIMIT);
    if (e != cudaSuccess) { set_error("stream_fused smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    done = true;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)G); lc.blockDim = dim3(FS_THREADS); lc.dynamicSmemBytes = smem_bytes; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: the grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, kern, a, sm);
  if (e != cudaSuccess) { set_error("stream_fused_kernel launch: %s", cudaGetErrorString(e)); cudaGetLastError(); return W2VS_CUDA_ERROR; }
  W2VS_CHECK_LAUNCH("stream_fused_kernel");
  return W2VS_OK;
}

w2vs_status_t debug_read_fused_trace(unsigned long long* out, int n) {
  if (n > 2 * 64 * 12) n = 2 * 64 * 12;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_fused_trace, (size_t)n * 8);
  if (e != cudaSuccess) { set_error("read g_fused_trace: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  return W2VS_OK;
}

w2vs_status_t debug_read_fused_fault(int* out) {
  int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_fused_fault, sizeof(int));
  if (e != cudaSuccess) { set_error("read g_fused_fault: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  *out = v;
  return W2VS_OK;
}

}  // namespace w2vs
