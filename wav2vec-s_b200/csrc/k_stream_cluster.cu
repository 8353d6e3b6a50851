// One decision step of the incremental encoder (one stream, pre-LN bf16 models) as ONE persistent kernel of
// thread-block clusters: 2 H clusters of 4 CTAs (two per attention head), two phases and two grid barriers per layer.
//
// What it replaces: the chain embed -> layers x [LN, QKV, attention, out_proj, LN, fc1, fc2] -> final LN of
// stream.cu's block_step (the reference computes the same rows by re-encoding the whole prefix,
// rain/simul/transducer_agent.py:138-167; layer arithmetic wav2vec2.py:921-978, attention
// modules/multihead_attention.py:162-194).  A step is at most 32 tokens: its arithmetic is nothing, the 613 MB of
// weights take 0.1 ms to read, and everything else is latency -- of dependent launches in the chain (7.5 us x 193),
// of grid barriers and of L2 -> SM broadcasts in the first persistent kernel (k_stream_fused.cu: five barriers per
// layer, and every operand that all 148 CTAs need whole costs its size x 148 of L2 bandwidth).  This kernel removes
// the broadcasts and three of the five barriers per layer by never letting a CTA need a whole operand:
//   phase A (cluster = head h):    x slice -> LN -> q/k/v of head h -> attention over the cache -> out_proj partial
//   phase C (cluster = F/H hidden): x slice -> LN -> fc1 + GELU for the cluster's hidden units -> fc2 partial
// Inside a cluster, CTA j of 8
//   1. loads ITS 1/8 of the feature dimension of the fp32 residual rows (K slice j), computes partial LayerNorm
//      statistics (mean, centred sum of squares), exchanges them through distributed shared memory (Chan's
//      combination: exact, one exchange), normalises its slice to bf16;
//   2. multiplies the slice with its K slice of the cluster's weight rows (mma.sync, fp32 accumulate) and
//      reduce-scatters the partial sums through DSMEM: CTA d receives the columns it owns from its 7 peers and adds
//      them in a fixed order;
//   3. applies bias (+ GELU / bf16 rounding) and all-gathers the small result (q/k/v of the head: 32 x 192; FFN
//      hidden of the cluster: 32 x F/H) into every CTA of the cluster; phase A then runs the attention of head h with
//      the keys split over the 32 warp groups of the cluster (K / V fragments straight from the cache in global
//      memory into registers, flash-style partial states merged through shared memory and DSMEM);
//   4. multiplies the gathered rows with its 1/8 of the OUTPUT columns of the second weight matrix (out_proj rows
//      restricted to head h's 64 inputs, fc2 rows restricted to the cluster's hidden units) and adds the partial
//      result into the next residual buffer with fp32 reductions at L2 (red.global.add.v2.f32; 16 to 32 additions per
//      element and phase in whatever order they arrive: the fp32 sums differ in the last bit from run to run, and
//      after 24 layers of bf16 roundings two runs over the same audio differ by up to two bf16 steps at the output.
//      Accumulating in 64-bit fixed point (red.global.add.u64) makes the step bit-reproducible -- built, tested and
//      measured at 30.6 instead of 24.2 us per layer (64-bit reductions and the loads behind them are that much
//      slower at L2), so it is not used; stream_step_impl = 1, the operator chain, gives the same bits every time).
// The residual stream rotates through three buffers: phase p reads X[p % 3], accumulates into X[(p + 1) % 3] (the
// rows' old values + bias are added by the CTA that holds them anyway) and zeroes X[(p + 2) % 3] for the phase after.
// Weights: a second copy in the packed blob (LayerW::wc) holds, per layer and CTA, the four operand pieces exactly as
// they lie in shared memory (rows padded by 16 bytes against bank conflicts), so a piece is a few bulk copies; they
// are requested one piece ahead (two 68 KB slots + the out_proj slot) and never stop at a barrier.
// Numerically the step matches the operator chain up to summation order (same bf16 roundings: operand rows, q/k/v,
// P, context, FFN hidden; fp32 statistics and residual stream).
#include <cuda.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"
#include "layout.h"

namespace w2vs {
__device__ int g_cluster_fault = 0;                       // a barrier / pipeline wait timed out (diagnostics)
__device__ unsigned long long g_cluster_trace[64][32];    // globaltimer stamps of CTA 0, [layer][event]
}
#define W2VS_TC_FAULT_FLAG (&::w2vs::g_cluster_fault)
#include "tc_common.cuh"

namespace w2vs {
namespace {
using namespace tc;

constexpr int CL = kStreamClusterSize;        // CTAs per cluster
constexpr int CW = 8, CT = 32 * CW;           // warps / threads per CTA
constexpr int ROWS = 32;                      // token rows of a step (two m16 tiles)
constexpr unsigned long long CL_TIMEOUT_NS = 2000000000ull;
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int MRG_BLK = 16 * 68 + 32;         // floats per merge block: [16][68] O + [16][2] (m, l)
constexpr int MRG_P = 68;                     // floats per row of a merge block (64 + 4: rows land in different banks)
static_assert(CL == 4, "warp -> destination CTA maps below assume four CTAs per cluster");

struct ClArgs {
  const uint8_t* W;                    // packed weights
  unsigned long long wc, bqkv, bo, ln1_w, ln1_b, b1, b2, ln2_w, ln2_b;   // layer-0 byte offsets
  unsigned long long layer_stride, enc_ln_w, enc_ln_b, sin_table;
  int layers, ntok, n_main, f0;        // tokens of this step, frames emitted, first frame index
  const float* feats;                  // projected frames [feat_rows][D] fp32 (stream 0)
  float* X;                            // residual stream, 3 x [32][D] fp32
  bf16* kv; long long kv_layer_elems;  // cache [layers][kv_rows][2D]
  bf16* out;                           // [n_main][D]
  unsigned long long* bar;             // [0..3] arrivals per cluster rank, [4] departures (all zero between launches)
  float scale_log2;
  int trace;                           // write the phase timestamps (tools/cluster_trace.py); off by default
};

template <int D, int F, int H>
struct CK {
  static constexpr int NC = 2 * H;               // clusters: (head, m tile) in phase A, F / NC hidden units in phase C
  static constexpr int KS = D / CL;              // K slice of the normalised operand per CTA
  static constexpr int NV4 = cmax(1, KS / 128);  // float4 per lane and row of the slice
  static constexpr int HC = F / NC;              // FFN hidden units per cluster
  static constexpr int HR = HC / CL;             // ... owned by one CTA after the reduce-scatter
  static constexpr int NO = D / CL;              // output columns per CTA in the second products
  static constexpr int PQ = KS * 2 + 16;         // row pitch (bytes) of K-slice operands
  static constexpr int PO = 64 * 2 + 16;         // context / out_proj rows
  static constexpr int PH = HC * 2 + 16;         // FFN hidden / fc2 rows
  static constexpr int PG = 192 * 2 + 16;        // gathered q|k|v rows
  // weight pieces of one (layer, CTA), in the order of use: [q|k rows] [v rows + out_proj rows] [fc1 rows] [fc2 rows]
  static constexpr int QK_B = 128 * PQ, V_B = 64 * PQ, WO_B = NO * PO, P1_B = V_B + WO_B, W1_B = HC * PQ, W2_B = NO * PH;
  static constexpr int CTA_B = QK_B + P1_B + W1_B + W2_B;
  static constexpr int BIG = (cmax(cmax(QK_B, P1_B), cmax(W1_B, W2_B)) + 127) / 128 * 128;
  static constexpr int NT1C = HC / 64;           // n tiles per warp, fc1
  static constexpr int NT2 = cmax(1, NO / 64);   // n tiles per warp, second products
  // shared memory map (bytes from a 128-aligned base)
  static constexpr int S_SLOT0 = 0, S_SLOT1 = BIG, S_STAT = 2 * BIG;
  static constexpr int S_A = S_STAT + CL * ROWS * 8;
  static constexpr int A_B = cmax(ROWS * PQ, 3072);                    // operand slice; later bf16 staging rows
  static constexpr int S_SCR = S_A + A_B;
  static constexpr int SCR_B = cmax(CL * ROWS * 48 * 4, CL * ROWS * HR * 4);
  static constexpr int S_G = S_SCR + SCR_B;
  static constexpr int G_B = cmax(ROWS * PG, ROWS * PH);
  static constexpr int S_CTX = S_G + G_B;
  static constexpr int S_MRG = S_CTX + 16 * PO;                       // attention merge: 4 x ([16][68] O + [16][2] (m, l)) fp32
  static constexpr int S_FIN = S_MRG + 4 * MRG_BLK * 4;            // [32 rows][8]: rstd, m_j - mean (j < 4) of the folded LayerNorm
  static constexpr int S_BAR = S_FIN + ROWS * 32;
  static constexpr int S_END = S_BAR + 128;
  // phase A scratch after the reduce: rx2 [4][16][16], ml_rx [4][16][2]   (fp32, written by the peers)
  static constexpr int X_RX2 = 8192, X_ML = 12288;
  // behind the CTA blocks of a layer: the vectors that fold the LayerNorm into the first products (fp32)
  //   sq [4][3D], bq [3D]: s_j[n] = sum over K slice j of gamma_k W[n][k], b[n] = bias[n] + sum_k beta_k W[n][k]   (q|k|v rows)
  //   s1 [4][F],  b1 [F]:  the same for fc1
  static constexpr size_t FOLD_OFF = (size_t)CL * NC * CTA_B;
  static constexpr size_t F_SQ = FOLD_OFF, F_BQ = F_SQ + (size_t)CL * 3 * D * 4, F_S1 = F_BQ + (size_t)3 * D * 4,
                          F_B1 = F_S1 + (size_t)CL * F * 4, LAYER_B = F_B1 + (size_t)F * 4;
  static_assert(D % 128 == 0 && D == 64 * H && HC % 64 == 0 && KS % 16 == 0, "model shape");
  static_assert(S_END + 128 <= 232448, "shared memory");
};

// ---- small PTX helpers ----
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
// Remote shared-memory stores that count their bytes on an mbarrier of the DESTINATION CTA (st.async): the receiver
// waits on its own barrier for the bytes it expects -- no fence on the sender, no cluster-wide barrier.  (With plain
// st.shared::cluster + barrier.cluster the release fence of every barrier cost ~0.6 us: 18 % of the kernel's samples.)
__device__ __forceinline__ void st_async_v2f(uint32_t addr, float a, float b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1,%2}, [%3];"
               ::"r"(addr), "f"(a), "f"(b), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.u32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// four independent warp sums in lock step (shuffles are convergent operations the compiler keeps in program order:
// four separate warp_sum calls are twenty dependent shuffle latencies, this is five)
__device__ __forceinline__ void warp_sum4(float (&v)[4]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = __shfl_xor_sync(0xffffffffu, v[i], o);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += t[i];
  }
}
__device__ __forceinline__ float ex2w(float m, float M, float sl2) { return m == -INFINITY ? 0.f : ex2_approx((m - M) * sl2); }

// Grid barrier (all CTAs are co-resident: cooperative launch).  One arrival counter PER CLUSTER RANK in global memory,
// monotonically increasing inside a launch; the last CTA to leave the kernel resets them.  Between two phases a CTA
// of rank j only depends on the rank-j CTAs of the other clusters (they own the same quarter of the feature
// dimension: the columns its second product adds into are the K slice its next first product reads), so it waits for
// its own rank's counter only -- a quarter of the arrivals on the line it polls, and no waiting for stragglers of the
// other ranks; `all` = wait for every rank (after the embedding and before the final LayerNorm, which touch whole rows).
__device__ __forceinline__ bool grid_barrier(unsigned long long* bar, int rank, unsigned long long target, bool all,
                                             unsigned long long* stamps = nullptr) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (stamps) stamps[0] = global_ns();
    // release at gpu scope: this CTA's writes and reductions of the phase (ordered before this thread by the
    // bar.sync above) are visible to whoever observes the arrival
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar + rank), "l"(1ull) : "memory");
    if (stamps) stamps[1] = global_ns();
    unsigned long long v, t0 = 0;
    unsigned spins = 0;
    int ok = 1;
    for (int k = 0; k < (all ? CL : 1) && ok; ++k) {
      const unsigned long long* p = bar + (all ? k : rank);
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= target) break;
        if ((++spins & 0x3ff) == 0) {
          const unsigned long long now = global_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > CL_TIMEOUT_NS) { atomicExch(&g_cluster_fault, 1); ok = 0; break; }
        }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// acc[mt][nt] += A[16 MT x 16 KSTEPS] . B_nt[8 x 16 KSTEPS]^T   (A, B row-major bf16 in shared memory, k contiguous;
// bb[nt] = shared address of the first of the 8 weight rows of n tile nt)
template <int NTW, int KSTEPS, int PA, int PB>
__device__ __forceinline__ void mma_block(uint32_t a_base, const uint32_t (&bb)[NTW], float (&acc)[2][NTW][4], int lane, bool two_mt) {
  const uint32_t a_lane = a_base + (uint32_t)(lane & 15) * PA + (uint32_t)(lane >> 4) * 16;
  const uint32_t b_lane = (uint32_t)(lane & 7) * PB + (uint32_t)((lane >> 3) & 1) * 16;
  // the fragments of k step ks + 1 are requested before the MMAs of step ks are issued (the ldmatrix asm statements
  // are volatile and stay in program order: without the look-ahead every step would wait for its own loads)
  uint32_t af[2][2][4], bf[2][NTW][2];
  auto load = [&](int ks, int buf) {
    ldsm_x4(a_lane + ks * 32, af[buf][0]);
    if (two_mt) ldsm_x4(a_lane + 16 * PA + ks * 32, af[buf][1]);
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) ldsm_x2(bb[nt] + b_lane + ks * 32, bf[buf][nt]);
  };
  load(0, 0);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const int cur = ks & 1;
    if (ks + 1 < KSTEPS) load(ks + 1, cur ^ 1);
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      mma_16816(acc[0][nt], af[cur][0][0], af[cur][0][1], af[cur][0][2], af[cur][0][3], bf[cur][nt][0], bf[cur][nt][1]);
      if (two_mt) mma_16816(acc[1][nt], af[cur][1][0], af[cur][1][1], af[cur][1][2], af[cur][1][3], bf[cur][nt][0], bf[cur][nt][1]);
    }
  }
}

template <int D, int F, int H>
__global__ void __launch_bounds__(CT, 1)
stream_cluster_kernel(const __grid_constant__ ClArgs a) {
  using K = CK<D, F, H>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;        // identical in every CTA: DSMEM offsets line up
  uint8_t* sm = smem_raw + (sb - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int G = gridDim.x, gid = blockIdx.x;
  const int rank = (int)cluster_ctarank(), cl = gid / CL;
  const int Mt = a.ntok;
  const bool two_mt = Mt > 16;
  const float sl2 = a.scale_log2;
  const uint32_t bar_slot = sb + K::S_BAR;
  // shared-window base of every CTA of the cluster (mapa reads a special register: once, not per store)
  const uint32_t rb0 = mapa(sb, 0), rb1 = mapa(sb, 1), rb2 = mapa(sb, 2), rb3 = mapa(sb, 3);
  auto rbase = [&](int d) { return d == 0 ? rb0 : (d == 1 ? rb1 : (d == 2 ? rb2 : rb3)); };
  const uint8_t* wc_cta = a.W + a.wc + (size_t)gid * K::CTA_B;     // + layer * layer_stride

  // ---- weight pieces: piece i = 4 layer + {0 q|k, 1 v + out_proj, 2 fc1, 3 fc2} lives in slot i & 1; thread 0
  //      requests piece i + 2 as soon as piece i has been consumed ----
  auto issue_piece = [&](int i) {
    const int l = i >> 2, k = i & 3;
    if (l >= a.layers) return;
    const uint8_t* src = wc_cta + (size_t)l * a.layer_stride +
                         (k == 0 ? 0 : (k == 1 ? K::QK_B : (k == 2 ? K::QK_B + K::P1_B : K::QK_B + K::P1_B + K::W1_B)));
    const uint32_t bytes = k == 0 ? K::QK_B : (k == 1 ? K::P1_B : (k == 2 ? K::W1_B : K::W2_B));
    const uint32_t dst = sb + ((i & 1) ? K::S_SLOT1 : K::S_SLOT0), bar = bar_slot + 8 * (i & 1);
    fence_async_smem();
    mbar_expect_tx(bar, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) bulk_g2s(dst + o, src + o, min(32768u, bytes - o), bar);
  };
  // piece i is the (i >> 1)-th use of its slot
  auto wait_piece = [&](int i) { return mbar_wait(bar_slot + 8 * (i & 1), (uint32_t)(i >> 1) & 1u); };
  // L2 prefetch of the parameter vectors phase `p` reads through plain loads: LayerNorm weight / bias and the second
  // product's bias over this CTA's K slice, the first product's bias over the columns this CTA reduces
  auto prefetch_params = [&](int p) {
    const int l = p >> 1;
    if (l >= a.layers) return;
    const uint8_t* Wl = a.W + (size_t)l * a.layer_stride;
    const bool pa_ = (p & 1) == 0;
    constexpr int LPV = (K::KS * 4 + 127) / 128;          // 128-byte lines per vector slice
    if (tid < 2 * LPV) {                                  // LayerNorm weight and the second product's bias over the K slice
      const int v = tid / LPV, i = tid - v * LPV;
      const unsigned long long off = pa_ ? (v == 0 ? a.ln1_w : a.bo) : (v == 0 ? a.ln2_w : a.b2);
      prefetch_l2(Wl + off + 4 * (rank * K::KS) + 128 * i);
    } else if (tid >= 64 && tid < 64 + 5 * 6) {           // fold vectors (4 x s_j, b) over the columns this CTA reduces
      const int v = (tid - 64) / 6, i = (tid - 64) % 6;
      if (pa_) {        // q|k|v rows of the head: three runs of 64 floats (two lines each)
        const size_t base = v < 4 ? K::F_SQ + (size_t)v * 3 * D * 4 : K::F_BQ;
        prefetch_l2(Wl + a.wc + base + 4 * ((i >> 1) * D + (cl >> 1) * 64) + 128 * (i & 1));
      } else if (i < 2) {
        const size_t base = v < 4 ? K::F_S1 + (size_t)v * F * 4 : K::F_B1;
        prefetch_l2(Wl + a.wc + base + 4 * (cl * K::HC + rank * K::HR) + 128 * i);
      }
    }
  };
  // exchange barriers (byte-counting, one use per layer): phase A statistics, partial q|k|v, gathered q|k|v, attention
  // partials, gathered context; phase C statistics, partial hidden units, gathered hidden units
  constexpr uint32_t XB_ASTAT = 16, XB_ARX1 = 24, XB_AG = 32, XB_ARX2 = 40, XB_ACTX = 48, XB_CSTAT = 56, XB_CRX1 = 64, XB_CG = 72;
  if (tid == 0) {
    mbar_init(bar_slot, 1); mbar_init(bar_slot + 8, 1);
    for (uint32_t o = XB_ASTAT; o <= XB_CG; o += 8) mbar_init(bar_slot + o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // (bulk copies are issued inside an elect_one() region, see tc_common.cuh: behind `tid == 0` every UBLKCP sits in an
  //  ELECT + R2UR.BROADCAST + BRA.U.ANY loop)
  __syncwarp();
  if (warp == 0 && elect_one()) { issue_piece(0); issue_piece(1); }
  // gather targets start as zeros (rows past the step's tokens are multiplied, never stored)
  for (int i = tid; i < (K::G_B + 16 * K::PO) / 16; i += CT) reinterpret_cast<uint4*>(sm + K::S_G)[i] = make_uint4(0u, 0u, 0u, 0u);

  const bool tr = a.trace != 0 && gid == 0 && tid == 0;
#define CL_TRACE(l_, ev_) do { if (tr && (l_) < 64) g_cluster_trace[l_][ev_] = global_ns(); } while (0)

  // ---- embed: X[0] = projected frame + sinusoidal position (absolute index frame + 2); X[1] = 0 ----
  {
    const int e = gid * CT + tid;          // 32 D == 256 G: one element per thread
    const int r = e / D, c = e - r * D;
    float v = 0.f;
    if (r < Mt)
      v = __ldcg(a.feats + (size_t)(a.f0 + r) * D + c) +
          reinterpret_cast<const float*>(a.W + a.sin_table)[(size_t)(a.f0 + r + 2) * D + c];
    a.X[e] = v;
    a.X[(size_t)ROWS * D + e] = 0.f;
  }
  prefetch_params(0);
  cluster_sync();                          // mbarrier inits and zeroed buffers before any remote traffic
  unsigned long long nbar = 0;
  bool ok = grid_barrier(a.bar, rank, ++nbar * K::NC, true);

#pragma unroll 1
  for (int ph = 0; ph < 2 * a.layers && ok; ++ph) {
    const int l = ph >> 1;
    const bool pa = (ph & 1) == 0;         // phase A (attention) / phase C (FFN)
    const float* Xin = a.X + (size_t)(ph % 3) * ROWS * D;
    float* Xout = a.X + (size_t)((ph + 1) % 3) * ROWS * D;
    const uint8_t* Wl = a.W + (size_t)l * a.layer_stride;
    CL_TRACE(l, pa ? 0 : 8);
    const bool a_on = !pa || 16 * (cl & 1) < Mt;      // phase A: does this cluster's query tile exist in this step?
    const uint32_t lpar = (uint32_t)l & 1u;           // every exchange barrier completes one phase per layer
    if (tid == 0) {
      if (pa) {
        mbar_expect_tx(bar_slot + XB_ASTAT, CL * ROWS * 8);
        if (a_on) {
          mbar_expect_tx(bar_slot + XB_ARX1, (uint32_t)(CL * Mt * 48 * 4));
          mbar_expect_tx(bar_slot + XB_AG, (uint32_t)(CL * Mt * 96));
          mbar_expect_tx(bar_slot + XB_ARX2, CL * 16 * (64 + 8));
          mbar_expect_tx(bar_slot + XB_ACTX, CL * 16 * 32);
        }
      } else {
        mbar_expect_tx(bar_slot + XB_CSTAT, CL * ROWS * 8);
        mbar_expect_tx(bar_slot + XB_CRX1, (uint32_t)(CL * Mt * K::HR * 4));
        mbar_expect_tx(bar_slot + XB_CG, (uint32_t)(CL * Mt * K::HR * 2));
      }
    }
    // ================= operand rows of the first product: the K slice, centred and scaled =================
    // LayerNorm is folded into the product (see the reduce stages): the operand is bf16((x - m_j) gamma) with m_j the
    // row's mean over THIS slice -- known after one warp sum, no exchange -- so the product starts an exchange and a
    // normalisation pass earlier; centring on the slice mean keeps the rounding as fine as that of the normalised
    // row (rounding x gamma itself would lose |mean| / std in accuracy).
    float4 x[4][K::NV4], gm[K::NV4];
#pragma unroll
    for (int j = 0; j < K::NV4; ++j) {
      const int c = 4 * lane + 128 * j;
      gm[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < K::KS) gm[j] = *reinterpret_cast<const float4*>(Wl + (pa ? a.ln1_w : a.ln2_w) + 4 * (rank * K::KS + c));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * warp + i;
        x[i][j] = (c < K::KS && r < Mt) ? __ldcg(reinterpret_cast<const float4*>(Xin + (size_t)r * D + rank * K::KS + c))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // The residual term of the second product (x + bias): the rows are dealt to the clusters (row r -> cluster r % NC),
    // a cluster's CTAs split the columns: 64 float4 per CTA.  Loaded now, added at the end of the phase.
    float4 res_x = make_float4(0.f, 0.f, 0.f, 0.f);
    const int res_row = cl + K::NC * (tid / (K::NO / 4)), res_col = rank * K::NO + 4 * (tid % (K::NO / 4));
    const bool res_on = tid < 64 && res_row < Mt;
    if (res_on) {
      const float4 xr = __ldcg(reinterpret_cast<const float4*>(Xin + (size_t)res_row * D + res_col));
      const float4 b = *reinterpret_cast<const float4*>(Wl + (pa ? a.bo : a.b2) + 4 * res_col);
      res_x = make_float4(xr.x + b.x, xr.y + b.y, xr.z + b.z, xr.w + b.w);
    }
    // the small parameter vectors of the NEXT phase (first touched there: a DRAM miss on its critical path otherwise)
    prefetch_params(ph + 1);
    if (pa && 16 * (cl & 1) < Mt) {
      // this warp's past K / V rows of head cl / 2 -> L2 (one 128-byte line per lane: 16 keys x {K, V} per step)
      const uint8_t* kv_h = reinterpret_cast<const uint8_t*>(a.kv + (size_t)l * a.kv_layer_elems + (cl >> 1) * 64);
      for (int s = rank * CW + warp; 16 * s < a.f0; s += CL * CW) {
        const int key = min(16 * s + (lane & 15), a.f0 - 1);
        prefetch_l2(kv_h + (size_t)key * (4 * D) + (lane >> 4) * (2 * D));
      }
    }

    float mean_i[4], m2_i[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < K::NV4; ++j) s += (x[i][j].x + x[i][j].y) + (x[i][j].z + x[i][j].w);
      mean_i[i] = s;
    }
    warp_sum4(mean_i);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * warp + i;
      const float mean = mean_i[i] * (1.0f / K::KS);
      float d2 = 0.f;
#pragma unroll
      for (int j = 0; j < K::NV4; ++j)
        if (4 * lane + 128 * j < K::KS) {
          const float d0 = x[i][j].x - mean, d1 = x[i][j].y - mean, d2_ = x[i][j].z - mean, d3 = x[i][j].w - mean;
          d2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2_, d2_, fmaf(d3, d3, d2))));
          uint2 u = make_uint2(0u, 0u);
          if (r < Mt) {
            u.x = pack_bf16x2(d0 * gm[j].x, d1 * gm[j].y);
            u.y = pack_bf16x2(d2_ * gm[j].z, d3 * gm[j].w);
          }
          *reinterpret_cast<uint2*>(sm + K::S_A + r * K::PQ + 2 * (4 * lane + 128 * j)) = u;
        }
      mean_i[i] = mean;
      m2_i[i] = d2;
    }
    CL_TRACE(l, pa ? 18 : 20);
    // the statistics of the slice (mean, centred sum of squares) for the reduce stage of every CTA of the cluster
    warp_sum4(m2_i);
    if (lane < CL) {
      const uint32_t dst = rbase(lane) + K::S_STAT + (uint32_t)((rank * ROWS + 4 * warp) * 8);
      const uint32_t mb = rbase(lane) + K::S_BAR + (pa ? XB_ASTAT : XB_CSTAT);
#pragma unroll
      for (int i = 0; i < 4; ++i) st_async_v2f(dst + i * 8, mean_i[i], m2_i[i], mb);
    }
    CL_TRACE(l, pa ? 19 : 21);
    __syncthreads();
    CL_TRACE(l, pa ? 1 : 9);
    // The folded LayerNorm, applied where the partial products are reduced:
    //   y[r][n] = rstd_r (sum_j P_j[r][n] + sum_j (m_j[r] - mean_r) s_j[n]) + b[n]
    // with P_j the product of slice j's operand rows, (m_j, M2_j) the slice statistics every CTA has received, mean and
    // variance of the row by Chan's combination, s_j / b the vectors packed with the weights.  One thread per row
    // leaves rstd and the four offsets in shared memory.
    auto fold_stats = [&]() {
      if (tid < ROWS) {
        const float2* st = reinterpret_cast<const float2*>(sm + K::S_STAT);
        float mj[CL], mean = 0.f, m2 = 0.f;
#pragma unroll
        for (int j = 0; j < CL; ++j) { mj[j] = st[j * ROWS + tid].x; mean += mj[j]; }
        mean *= 1.0f / CL;
#pragma unroll
        for (int j = 0; j < CL; ++j) { const float dv = mj[j] - mean; m2 += st[j * ROWS + tid].y + (float)K::KS * dv * dv; }
        float* fin = reinterpret_cast<float*>(sm + K::S_FIN) + tid * 8;
        fin[0] = 1.0f / sqrtf(m2 * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int j = 0; j < CL; ++j) fin[1 + j] = mj[j] - mean;
      }
      __syncthreads();
    };

    if (pa) {
      // ================= q | k | v of head cl / 2 (all rows): K split over the CTAs, reduce-scatter =================
      const int head = cl >> 1, mtile = cl & 1;
      // K / V fragments of one attention step (16 keys) of this warp: key-split group grp takes steps grp, grp + 32, ...
      struct KVF { uint4 ka[2], kc[2], vv[4]; int valid; };
      const int grp = rank * CW + warp;
      const int n_past = (a.f0 + 15) >> 4, n_steps = n_past + ((Mt + 15) >> 4);
      const uint8_t* kv_h = reinterpret_cast<const uint8_t*>(a.kv + (size_t)l * a.kv_layer_elems + head * 64);
      auto load_step = [&](int s, KVF& f) {
        const uint8_t* kb; size_t pitch; int voff;
        if (s < n_past) { kb = kv_h + (size_t)(16 * s) * (4 * D); pitch = 4 * D; f.valid = min(16, a.f0 - 16 * s); voff = 2 * D; }
        else { const int t0 = 16 * (s - n_past); kb = sm + K::S_G + t0 * K::PG + 128; pitch = K::PG; f.valid = min(16, Mt - t0); voff = 128; }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint8_t* p = kb + (size_t)min(8 * t + g, f.valid - 1) * pitch + 32 * q;
          f.ka[t] = *reinterpret_cast<const uint4*>(p);
          f.kc[t] = *reinterpret_cast<const uint4*>(p + 16);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = 2 * q + (e & 1) + 8 * (e >> 1);
          f.vv[e] = *reinterpret_cast<const uint4*>(kb + voff + (size_t)min(key, f.valid - 1) * pitch + 16 * g);
        }
      };
      KVF cur, nxt;
      float2 fsj[CL], fb = make_float2(0.f, 0.f);
      ok = wait_piece(4 * l) && ok;
      ok = wait_piece(4 * l + 1) && ok;
      CL_TRACE(l, 15);
      if (a_on) {
        // (eight warps x three n tiles: the product is bound by the issue rate of mma.sync -- 768 HMMA.16816 per CTA at one
        //  per ~12 cycles and SM sub-partition = 1.15 us measured; four warps x six n tiles, which re-read the operand
        //  tile from shared memory half as often, take 1.6 us)
        float acc[2][3][4] = {};
        uint32_t bb[3];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const int n = 3 * warp + nt;           // n tile of the head's 192 columns: q | k in slot 0, v in slot 1
          bb[nt] = n < 16 ? sb + K::S_SLOT0 + (uint32_t)(n * 8) * K::PQ : sb + K::S_SLOT1 + (uint32_t)((n - 16) * 8) * K::PQ;
        }
        mma_block<3, K::KS / 16, K::PQ, K::PQ>(sb + K::S_A, bb, acc, lane, two_mt);
        CL_TRACE(l, 24);
        const uint32_t dst = rbase((warp >> 1)) + K::S_SCR + (uint32_t)(rank * ROWS * 48 * 4);
        const uint32_t mb = rbase((warp >> 1)) + K::S_BAR + XB_ARX1;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            const int r0 = mt * 16 + g, c = (warp & 1) * 24 + nt * 8 + 2 * q;
            if (r0 < Mt) st_async_v2f(dst + (uint32_t)((r0 * 48 + c) * 4), acc[mt][nt][0], acc[mt][nt][1], mb);
            if (r0 + 8 < Mt) st_async_v2f(dst + (uint32_t)(((r0 + 8) * 48 + c) * 4), acc[mt][nt][2], acc[mt][nt][3], mb);
          }
        // the first attention step of this warp, if it reads the cache (rows of earlier decision steps): in flight
        // during the reduce and gather stages
        CL_TRACE(l, 25);
        if (grp < n_past) load_step(grp, cur);
        // the reduce stage below: thread = (column pair cp of the 24 this CTA owns, rows rg, rg + 10, rg + 20); its fold
        // vector entries are requested now and arrive while the partial sums are still in flight
        if (tid < 240) {
          const int cp = tid % 24, col = 48 * rank + 2 * cp, n = (col >> 6) * D + head * 64 + (col & 63);
          const float* sq = reinterpret_cast<const float*>(Wl + a.wc + K::F_SQ);
#pragma unroll
          for (int s = 0; s < CL; ++s) fsj[s] = *reinterpret_cast<const float2*>(sq + (size_t)s * 3 * D + n);
          fb = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(Wl + a.wc + K::F_BQ) + n);
        }
        __syncthreads();       // all warps are past their reads of the operand slice: its memory becomes the staging rows
        ok = mbar_wait(bar_slot + XB_ARX1, lpar) && ok;
      }
      CL_TRACE(l, 2);
      if (!a_on && warp == 0 && elect_one()) issue_piece(4 * l + 2);
      if (a_on) {
        // ---- reduce, bias, bf16, all-gather; K / V rows -> cache (by the cluster of the head's first query tile) ----
        uint32_t* stage = reinterpret_cast<uint32_t*>(sm + K::S_A);          // [32][24] bf16 pairs
        const float* rx = reinterpret_cast<const float*>(sm + K::S_SCR);
        ok = mbar_wait(bar_slot + XB_ASTAT, lpar) && ok;
        fold_stats();
        if (tid < 240) {
          const int cp = tid % 24;
          for (int r = tid / 24; r < Mt; r += 10) {
            const float* fin = reinterpret_cast<const float*>(sm + K::S_FIN) + r * 8;
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int s = 0; s < CL; ++s) {
              const float2 p = *reinterpret_cast<const float2*>(rx + (s * ROWS + r) * 48 + 2 * cp);
              v0 += p.x + fin[1 + s] * fsj[s].x; v1 += p.y + fin[1 + s] * fsj[s].y;
            }
            stage[r * 24 + cp] = pack_bf16x2(fmaf(fin[0], v0, fb.x), fmaf(fin[0], v1, fb.y));
          }
        }
        __syncthreads();
        // every warp of this CTA is past its reads of the q|k rows: request fc1's rows into that slot
        if (warp == 0 && elect_one()) issue_piece(4 * l + 2);
        for (int idx = tid; idx < Mt * 6 * CL; idx += CT) {
          const int d = idx & (CL - 1), rc_ = idx >> 2, r = rc_ / 6, ch = rc_ - 6 * r;
          const uint4 v = *reinterpret_cast<const uint4*>(sm + K::S_A + r * 96 + ch * 16);
          st_async_v4(rbase(d) + K::S_G + (uint32_t)(r * K::PG + rank * 96 + ch * 16), v, rbase(d) + K::S_BAR + XB_AG);
        }
        ok = mbar_wait(bar_slot + XB_AG, lpar) && ok;
      }
      CL_TRACE(l, 3);

      if (a_on) {
        // ================= attention: 16 query rows, keys split over the 32 warps of the cluster =================
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        {
          const uint8_t* qr = sm + K::S_G + (mtile * 16 + g) * K::PG + 32 * q;
          const uint4 q0a = *reinterpret_cast<const uint4*>(qr), q0b = *reinterpret_cast<const uint4*>(qr + 16);
          const uint4 q1a = *reinterpret_cast<const uint4*>(qr + 8 * K::PG), q1b = *reinterpret_cast<const uint4*>(qr + 8 * K::PG + 16);
          // steps on this step's own keys (the gathered k | v rows in shared memory) are loaded now
          if (grp < n_steps && grp >= n_past) load_step(grp, cur);
#pragma unroll 1
          for (int s = grp; s < n_steps; s += CL * CW) {
            // one step ahead: the next fragments are in flight while this step is computed
            if (s + CL * CW < n_steps) load_step(s + CL * CW, nxt);
            const int valid = cur.valid;
            // S = Q K^T for 16 keys (the d index is permuted identically for Q and K: lane q owns d = 16 q .. 16 q + 15)
            float sc[2][4];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f;
              mma_16816(sc[t], q0a.x, q1a.x, q0a.y, q1a.y, cur.ka[t].x, cur.ka[t].y);
              mma_16816(sc[t], q0a.z, q1a.z, q0a.w, q1a.w, cur.ka[t].z, cur.ka[t].w);
              mma_16816(sc[t], q0b.x, q1b.x, q0b.y, q1b.y, cur.kc[t].x, cur.kc[t].y);
              mma_16816(sc[t], q0b.z, q1b.z, q0b.w, q1b.w, cur.kc[t].z, cur.kc[t].w);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const bool vis = 8 * t + 2 * q + e < valid;
                sc[t][e] = vis ? sc[t][e] : -INFINITY;
                sc[t][2 + e] = vis ? sc[t][2 + e] : -INFINITY;
                mx0 = fmaxf(mx0, sc[t][e]);
                mx1 = fmaxf(mx1, sc[t][2 + e]);
              }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: a step has at least one key
            const float a0 = ex2w(m0, mn0, sl2), a1 = ex2w(m1, mn1, sl2);
            m0 = mn0; m1 = mn1;
            const float ms0 = mn0 * sl2, ms1 = mn1 * sl2;
            uint32_t pf[4];
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const float p0 = ex2_approx(fmaf(sc[t][0], sl2, -ms0)), p1 = ex2_approx(fmaf(sc[t][1], sl2, -ms0));
              const float p2 = ex2_approx(fmaf(sc[t][2], sl2, -ms1)), p3 = ex2_approx(fmaf(sc[t][3], sl2, -ms1));
              sum0 += p0 + p1;
              sum1 += p2 + p3;
              pf[2 * t] = pack_bf16x2(p0, p1);          // row g,     keys 8 t + 2 q ..
              pf[2 * t + 1] = pack_bf16x2(p2, p3);      // row g + 8
            }
            l0 = l0 * a0 + sum0;
            l1 = l1 * a1 + sum1;
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }
            // O += P V: accumulator column n of tile j is d = 8 n + j (lane group g supplies d = 8 g .. 8 g + 7)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
              const uint32_t v0 = (&cur.vv[0].x)[j >> 1], v1 = (&cur.vv[1].x)[j >> 1], v2 = (&cur.vv[2].x)[j >> 1], v3 = (&cur.vv[3].x)[j >> 1];
              mma_16816(o[j], pf[0], pf[1], pf[2], pf[3], prmt(v0, v1, sel), prmt(v2, v3, sel));
            }
            cur = nxt;
          }
          l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
          l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
          l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
          l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        }
        // ---- merge the eight warps of this CTA, then the four CTAs ----
        // Warps 4-7 hand their flash state (m, l, O) to warps 0-3 through shared memory, which fold it into their own
        // and publish the result; the sending threads below fold those four states while they read them.  (Shared-
        // memory float atomics would be compare-and-swap loops: 10 us per layer when eight warps hit the same rows.)
        float* scr = reinterpret_cast<float*>(sm + K::S_SCR);
        {
          float* mb = reinterpret_cast<float*>(sm + K::S_MRG) + (warp & 3) * MRG_BLK;     // [16][68] O, then [16][2] (m, l)
          float* mlb = mb + 16 * MRG_P;
          auto put = [&]() {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float* r0p = mb + g * MRG_P + 16 * q + 8 * hf;
              *reinterpret_cast<float4*>(r0p) = make_float4(o[0][hf], o[1][hf], o[2][hf], o[3][hf]);
              *reinterpret_cast<float4*>(r0p + 4) = make_float4(o[4][hf], o[5][hf], o[6][hf], o[7][hf]);
              float* r1p = r0p + 8 * MRG_P;
              *reinterpret_cast<float4*>(r1p) = make_float4(o[0][2 + hf], o[1][2 + hf], o[2][2 + hf], o[3][2 + hf]);
              *reinterpret_cast<float4*>(r1p + 4) = make_float4(o[4][2 + hf], o[5][2 + hf], o[6][2 + hf], o[7][2 + hf]);
            }
            if (q == 0) {
              *reinterpret_cast<float2*>(mlb + 2 * g) = make_float2(m0, l0);
              *reinterpret_cast<float2*>(mlb + 2 * (g + 8)) = make_float2(m1, l1);
            }
          };
          if (warp >= 4) put();
          __syncthreads();
          if (warp < 4) {
            const float2 p0 = *reinterpret_cast<const float2*>(mlb + 2 * g), p1 = *reinterpret_cast<const float2*>(mlb + 2 * (g + 8));
            const float M0 = fmaxf(m0, p0.x), M1 = fmaxf(m1, p1.x);
            const float wa0 = ex2w(m0, M0, sl2), wb0 = ex2w(p0.x, M0, sl2), wa1 = ex2w(m1, M1, sl2), wb1 = ex2w(p1.x, M1, sl2);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const float* r0p = mb + g * MRG_P + 16 * q + 8 * hf;
              const float4 x0 = *reinterpret_cast<const float4*>(r0p), x1 = *reinterpret_cast<const float4*>(r0p + 4);
              const float4 y0 = *reinterpret_cast<const float4*>(r0p + 8 * MRG_P), y1 = *reinterpret_cast<const float4*>(r0p + 8 * MRG_P + 4);
              const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w}, ys[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                o[j][hf] = o[j][hf] * wa0 + xs[j] * wb0;
                o[j][2 + hf] = o[j][2 + hf] * wa1 + ys[j] * wb1;
              }
            }
            l0 = l0 * wa0 + p0.y * wb0; l1 = l1 * wa1 + p1.y * wb1;
            m0 = M0; m1 = M1;
            __syncwarp();
            put();
          }
          __syncthreads();
        }
        CL_TRACE(l, 4);
        {
          // CTA d merges the context columns [16 d, 16 d + 16): one 16-byte store per thread (row, destination, quarter)
          const int r = tid >> 4, d = (tid >> 2) & 3, qq = tid & 3;
          const float* mb = reinterpret_cast<const float*>(sm + K::S_MRG);
          float mv[4], lv[4], M = -INFINITY;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float2 v = *reinterpret_cast<const float2*>(mb + b * MRG_BLK + 16 * MRG_P + 2 * r);
            mv[b] = v.x; lv[b] = v.y;
            M = fmaxf(M, v.x);
          }
          float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float lsum = 0.f;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float w = ex2w(mv[b], M, sl2);
            const float4 v = *reinterpret_cast<const float4*>(mb + b * MRG_BLK + r * MRG_P + 16 * d + 4 * qq);
            acc4.x = fmaf(v.x, w, acc4.x); acc4.y = fmaf(v.y, w, acc4.y); acc4.z = fmaf(v.z, w, acc4.z); acc4.w = fmaf(v.w, w, acc4.w);
            lsum = fmaf(lv[b], w, lsum);
          }
          st_async_v4(rbase(d) + K::S_SCR + K::X_RX2 + (uint32_t)(((rank * 16 + r) * 16 + 4 * qq) * 4),
                      make_uint4(__float_as_uint(acc4.x), __float_as_uint(acc4.y), __float_as_uint(acc4.z), __float_as_uint(acc4.w)),
                      rbase(d) + K::S_BAR + XB_ARX2);
          if (qq == 0) st_async_v2f(rbase(d) + K::S_SCR + K::X_ML + (uint32_t)((rank * 16 + r) * 8), M, lsum, rbase(d) + K::S_BAR + XB_ARX2);
        }
        ok = mbar_wait(bar_slot + XB_ARX2, lpar) && ok;
        CL_TRACE(l, 5);
        {
          const int r = tid >> 4, cc = tid & 15;
          const float2* ml = reinterpret_cast<const float2*>(sm + K::S_SCR + K::X_ML);
          const float* rx2 = reinterpret_cast<const float*>(sm + K::S_SCR + K::X_RX2);
          float M = -INFINITY;
#pragma unroll
          for (int s = 0; s < CL; ++s) M = fmaxf(M, ml[s * 16 + r].x);
          float num = 0.f, den = 0.f;
#pragma unroll
          for (int s = 0; s < CL; ++s) {
            const float2 v = ml[s * 16 + r];
            const float w = ex2w(v.x, M, sl2);
            den = fmaf(v.y, w, den);
            num = fmaf(rx2[(s * 16 + r) * 16 + cc], w, num);
          }
          reinterpret_cast<bf16*>(sm + K::S_A)[r * 16 + cc] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
        }
        __syncthreads();
        if (tid < 16 * 2 * CL) {
          const int r = tid >> 3, ch = (tid >> 2) & 1, d = tid & 3;
          st_async_v4(rbase(d) + K::S_CTX + (uint32_t)(r * K::PO + rank * 32 + ch * 16),
                      *reinterpret_cast<const uint4*>(sm + K::S_A + r * 32 + ch * 16), rbase(d) + K::S_BAR + XB_ACTX);
        }
        ok = mbar_wait(bar_slot + XB_ACTX, lpar) && ok;
        CL_TRACE(l, 6);
        // ================= out_proj restricted to the head: this CTA's D/4 output columns, 16 rows =================
        if (warp * K::NT2 * 8 < K::NO) {
          float acc[2][K::NT2][4] = {};
          uint32_t bb[K::NT2];
#pragma unroll
          for (int nt = 0; nt < K::NT2; ++nt) bb[nt] = sb + K::S_SLOT1 + K::V_B + (uint32_t)((warp * K::NT2 + nt) * 8) * K::PO;
          mma_block<K::NT2, 4, K::PO, K::PO>(sb + K::S_CTX, bb, acc, lane, false);
#pragma unroll
          for (int nt = 0; nt < K::NT2; ++nt) {
            const int r0 = mtile * 16 + g, col = rank * K::NO + (warp * K::NT2 + nt) * 8 + 2 * q;
            if (r0 < Mt) red_add_v2(Xout + (size_t)r0 * D + col, acc[0][nt][0], acc[0][nt][1]);
            if (r0 + 8 < Mt) red_add_v2(Xout + (size_t)(r0 + 8) * D + col, acc[0][nt][2], acc[0][nt][3]);
          }
        }
        if (mtile == 0) {
          // this step's K / V rows of the head -> cache (every CTA holds the gathered rows: each writes a quarter)
          bf16* kv_l = a.kv + (size_t)l * a.kv_layer_elems;
          for (int idx = rank * CT + tid; idx < Mt * 16; idx += CL * CT) {
            const int r = idx >> 4, ch = idx & 15;            // 16-byte chunk ch of [k | v] (8 + 8 chunks)
            *reinterpret_cast<uint4*>(kv_l + (size_t)(a.f0 + r) * (2 * D) + (size_t)(ch >> 3) * D + head * 64 + (ch & 7) * 8) =
                *reinterpret_cast<const uint4*>(sm + K::S_G + r * K::PG + 128 + ch * 16);
          }
        }
      }
      __syncthreads();
      if (warp == 0 && elect_one()) issue_piece(4 * l + 3);
      CL_TRACE(l, 7);
    } else {
      // ================= fc1 for the cluster's hidden units: K split over the CTAs, reduce-scatter =================
      ok = wait_piece(4 * l + 2) && ok;
      CL_TRACE(l, 16);
      {
        float acc[2][K::NT1C][4] = {};
        uint32_t bb[K::NT1C];
#pragma unroll
        for (int nt = 0; nt < K::NT1C; ++nt) bb[nt] = sb + K::S_SLOT0 + (uint32_t)((warp * K::NT1C + nt) * 8) * K::PQ;
        mma_block<K::NT1C, K::KS / 16, K::PQ, K::PQ>(sb + K::S_A, bb, acc, lane, two_mt);
        const uint32_t dst = rbase((warp >> 1)) + K::S_SCR + (uint32_t)(rank * ROWS * K::HR * 4);
        const uint32_t mb = rbase((warp >> 1)) + K::S_BAR + XB_CRX1;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < K::NT1C; ++nt) {
            const int r0 = mt * 16 + g, c = ((warp & 1) * K::NT1C + nt) * 8 + 2 * q;
            if (r0 < Mt) st_async_v2f(dst + (uint32_t)((r0 * K::HR + c) * 4), acc[mt][nt][0], acc[mt][nt][1], mb);
            if (r0 + 8 < Mt) st_async_v2f(dst + (uint32_t)(((r0 + 8) * K::HR + c) * 4), acc[mt][nt][2], acc[mt][nt][3], mb);
          }
      }
      // the reduce stage below: thread = (column pair cp of the HR / 2 this CTA owns, rows rg, rg + 256 / HP, ...); its
      // fold vector entries are requested now
      constexpr int HP = K::HR / 2, RSTEP = CT / HP;
      float2 fsj[CL], fb;
      {
        const int cp = tid % HP;
        const float* s1 = reinterpret_cast<const float*>(Wl + a.wc + K::F_S1) + cl * K::HC + rank * K::HR;
#pragma unroll
        for (int s = 0; s < CL; ++s) fsj[s] = *reinterpret_cast<const float2*>(s1 + (size_t)s * F + 2 * cp);
        fb = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(Wl + a.wc + K::F_B1) + cl * K::HC + rank * K::HR + 2 * cp);
      }
      __syncthreads();         // all warps are past their reads of the operand slice: its memory becomes the staging rows
      ok = mbar_wait(bar_slot + XB_CRX1, lpar) && ok;
      CL_TRACE(l, 10);
      // ---- reduce, bias, GELU, bf16, all-gather ----
      {
        uint32_t* stage = reinterpret_cast<uint32_t*>(sm + K::S_A);          // [32][HR / 2] bf16 pairs
        const float* rx = reinterpret_cast<const float*>(sm + K::S_SCR);
        ok = mbar_wait(bar_slot + XB_CSTAT, lpar) && ok;
        fold_stats();
        {
          const int cp = tid % HP;
          for (int r = tid / HP; r < Mt; r += RSTEP) {
            const float* fin = reinterpret_cast<const float*>(sm + K::S_FIN) + r * 8;
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int s = 0; s < CL; ++s) {
              const float2 p = *reinterpret_cast<const float2*>(rx + (s * ROWS + r) * K::HR + 2 * cp);
              v0 += p.x + fin[1 + s] * fsj[s].x; v1 += p.y + fin[1 + s] * fsj[s].y;
            }
            stage[r * HP + cp] = pack_bf16x2(gelu_tanh(fmaf(fin[0], v0, fb.x)), gelu_tanh(fmaf(fin[0], v1, fb.y)));
          }
        }
        __syncthreads();
        if (warp == 0 && elect_one()) issue_piece(4 * l + 4);      // every warp is past fc1's rows: next layer's q|k rows into that slot
        constexpr int CH = K::HR * 2 / 16;       // 16-byte chunks per row
        for (int idx = tid; idx < Mt * CH * CL; idx += CT) {
          const int d = idx & (CL - 1), rc_ = idx >> 2, r = rc_ / CH, ch = rc_ - CH * r;
          const uint4 v = *reinterpret_cast<const uint4*>(sm + K::S_A + r * (K::HR * 2) + ch * 16);
          st_async_v4(rbase(d) + K::S_G + (uint32_t)(r * K::PH + rank * (K::HR * 2) + ch * 16), v, rbase(d) + K::S_BAR + XB_CG);
        }
      }
      ok = mbar_wait(bar_slot + XB_CG, lpar) && ok;
      CL_TRACE(l, 11);
      // ================= fc2 restricted to the cluster's hidden units: this CTA's D/4 output columns =================
      ok = wait_piece(4 * l + 3) && ok;
      CL_TRACE(l, 17);
      if (warp * K::NT2 * 8 < K::NO) {
        float acc[2][K::NT2][4] = {};
        uint32_t bb[K::NT2];
#pragma unroll
        for (int nt = 0; nt < K::NT2; ++nt) bb[nt] = sb + K::S_SLOT1 + (uint32_t)((warp * K::NT2 + nt) * 8) * K::PH;
        mma_block<K::NT2, K::HC / 16, K::PH, K::PH>(sb + K::S_G, bb, acc, lane, two_mt);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < K::NT2; ++nt) {
            const int r0 = mt * 16 + g, col = rank * K::NO + (warp * K::NT2 + nt) * 8 + 2 * q;
            if (r0 < Mt) red_add_v2(Xout + (size_t)r0 * D + col, acc[mt][nt][0], acc[mt][nt][1]);
            if (r0 + 8 < Mt) red_add_v2(Xout + (size_t)(r0 + 8) * D + col, acc[mt][nt][2], acc[mt][nt][3]);
          }
      }
      __syncthreads();
      if (warp == 0 && elect_one()) issue_piece(4 * l + 5);      // next layer's v + out_proj rows
      CL_TRACE(l, 12);
    }
    // zero the buffer the NEXT phase accumulates into (it was last read in the previous phase).  Global writes sit at
    // the end of a phase: before a cluster barrier they would hold up its release fence for an L2 round trip.
    // (this CTA's share of that buffer lies in its rank's columns, like everything else it writes: see grid_barrier)
    if (tid < 64) *reinterpret_cast<float4*>(a.X + (size_t)((ph + 2) % 3) * ROWS * D + (size_t)res_row * D + res_col) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res_on) red_add_v4(Xout + (size_t)res_row * D + res_col, res_x);
    ok = grid_barrier(a.bar, rank, ++nbar * K::NC, ph + 1 == 2 * a.layers, tr && !pa && l < 64 ? &g_cluster_trace[l][22] : nullptr) && ok;
    CL_TRACE(l, pa ? 13 : 14);
  }

  if (ok) {
    // ---- final LayerNorm of the emitted frames (encoder.layer_norm), one warp per row ----
    const float* Xf = a.X + (size_t)((2 * a.layers) % 3) * ROWS * D;
    const float* gf = reinterpret_cast<const float*>(a.W + a.enc_ln_w);
    const float* bf = reinterpret_cast<const float*>(a.W + a.enc_ln_b);
    constexpr int NV = D / 128;
    for (int r = gid * CW + warp; r < a.n_main; r += G * CW) {
      float4 v[NV];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        v[j] = __ldcg(reinterpret_cast<const float4*>(Xf + (size_t)r * D + 4 * lane + 128 * j));
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
      const float mean = warp_sum(s) * (1.0f / D);
      float qq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
        qq = fmaf(d0, d0, qq); qq = fmaf(d1, d1, qq); qq = fmaf(d2, d2, qq); qq = fmaf(d3, d3, qq);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(qq) * (1.0f / D) + 1e-5f);
      bf16* dst = a.out + (size_t)r * D + 4 * lane;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 gg = *reinterpret_cast<const float4*>(gf + 4 * lane + 128 * j);
        const float4 bb = *reinterpret_cast<const float4*>(bf + 4 * lane + 128 * j);
        uint2 u;
        u.x = pack_bf16x2((v[j].x - mean) * rstd * gg.x + bb.x, (v[j].y - mean) * rstd * gg.y + bb.y);
        u.y = pack_bf16x2((v[j].z - mean) * rstd * gg.z + bb.z, (v[j].w - mean) * rstd * gg.w + bb.w);
        *reinterpret_cast<uint2*>(dst + 128 * j) = u;
      }
    }
  }
  // ---- leave: no CTA may exit while a peer can still address its shared memory; the last CTA out resets the barrier
  cluster_sync();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(a.bar + CL, 1ull) == (unsigned long long)G - 1) {
      for (int k = 0; k <= CL; ++k) a.bar[k] = 0ull;
      __threadfence();
    }
  }
}

// ---- weight pieces of one (layer, CTA): pack_cluster_kernel writes them as they lie in shared memory ----
template <int D, int F, int H>
__global__ void __launch_bounds__(256)
pack_cluster_kernel(const bf16* __restrict__ wqkv, const bf16* __restrict__ wo, const bf16* __restrict__ w1,
                    const bf16* __restrict__ w2, uint8_t* __restrict__ dst) {
  using K = CK<D, F, H>;
  const int gid = blockIdx.x, cl = gid / CL, rank = gid % CL, head = cl >> 1;
  bf16* out = reinterpret_cast<bf16*>(dst + (size_t)gid * K::CTA_B);
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (int i = threadIdx.x; i < K::CTA_B / 2; i += blockDim.x) {
    int e = i;
    bf16 v = zero;
    if (e < (K::QK_B + K::V_B) / 2) {            // q | k | v rows of the head (contiguous: the v rows open piece 1)
      const int n = e / (K::PQ / 2), k = e % (K::PQ / 2);
      if (k < K::KS) v = wqkv[((size_t)(n >> 6) * D + head * 64 + (n & 63)) * D + rank * K::KS + k];
    } else if ((e -= (K::QK_B + K::V_B) / 2) < K::WO_B / 2) {
      const int n = e / (K::PO / 2), k = e % (K::PO / 2);
      if (k < 64) v = wo[((size_t)rank * K::NO + n) * D + head * 64 + k];
    } else if ((e -= K::WO_B / 2) < K::W1_B / 2) {
      const int n = e / (K::PQ / 2), k = e % (K::PQ / 2);
      if (k < K::KS) v = w1[((size_t)cl * K::HC + n) * D + rank * K::KS + k];
    } else {
      e -= K::W1_B / 2;
      const int n = e / (K::PH / 2), k = e % (K::PH / 2);
      if (k < K::HC) v = w2[((size_t)rank * K::NO + n) * F + cl * K::HC + k];
    }
    out[i] = v;
  }
}

// fold vectors of one layer: one warp per weight row n (q|k|v rows, then fc1 rows)
template <int D, int F, int H>
__global__ void __launch_bounds__(256)
pack_fold_kernel(const bf16* __restrict__ wqkv, const bf16* __restrict__ w1, const float* __restrict__ ln1_w,
                 const float* __restrict__ ln1_b, const float* __restrict__ bqkv, const float* __restrict__ ln2_w,
                 const float* __restrict__ ln2_b, const float* __restrict__ b1, uint8_t* __restrict__ dst) {
  using K = CK<D, F, H>;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= 3 * D + F) return;
  const bool qkv = n < 3 * D;
  const int m = qkv ? n : n - 3 * D;
  const bf16* w = (qkv ? wqkv : w1) + (size_t)m * D;
  const float* gm = qkv ? ln1_w : ln2_w;
  const float* bt = qkv ? ln1_b : ln2_b;
  float sj[CL] = {}, sb = 0.f;
#pragma unroll
  for (int j = 0; j < CL; ++j)
    for (int k = j * K::KS + lane; k < (j + 1) * K::KS; k += 32) {
      const float wv = __bfloat162float(w[k]);
      sj[j] = fmaf(gm[k], wv, sj[j]);
      sb = fmaf(bt[k], wv, sb);
    }
#pragma unroll
  for (int j = 0; j < CL; ++j) sj[j] = warp_sum(sj[j]);
  sb = warp_sum(sb);
  if (lane == 0) {
    float* sv = reinterpret_cast<float*>(dst + (qkv ? K::F_SQ : K::F_S1));
    float* bv = reinterpret_cast<float*>(dst + (qkv ? K::F_BQ : K::F_B1));
    const int N = qkv ? 3 * D : F;
#pragma unroll
    for (int j = 0; j < CL; ++j) sv[(size_t)j * N + m] = sj[j];
    bv[m] = (qkv ? bqkv[m] : b1[m]) + sb;
  }
}

template <int D, int F, int H>
void launch_config(cudaLaunchConfig_t& lc, cudaLaunchAttribute (&attr)[2], cudaStream_t st) {
  using K = CK<D, F, H>;
  lc = cudaLaunchConfig_t{};
  lc.gridDim = dim3((unsigned)(CL * K::NC)); lc.blockDim = dim3(CT); lc.dynamicSmemBytes = K::S_END + 128; lc.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: the grid barriers cannot deadlock
  attr[1].val.cooperative = 1;
  lc.attrs = attr; lc.numAttrs = 2;
}

// Clusters of this kernel the current device holds at a time (the grid barriers need all 2 H of them resident; how
// many fit depends on how the SMs of the part are spread over its GPCs).  Cached per device; -1 on a CUDA error.
template <int D, int F, int H>
int resident_clusters() {
  using K = CK<D, F, H>;
  static int cached[kMaxDevices];     // 0 = not asked yet
  int& n = cached[current_device()];
  if (n == 0) {
    auto kern = stream_cluster_kernel<D, F, H>;
    cudaLaunchConfig_t lc; cudaLaunchAttribute attr[2];
    launch_config<D, F, H>(lc, attr, nullptr);
    int v = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::S_END + 128) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&v, kern, &lc) != cudaSuccess) { cudaGetLastError(); v = -1; }
    n = v == 0 ? -1 : v;
  }
  return n;
}

template <int D, int F, int H>
w2vs_status_t launch_t(const ClArgs& a, cudaStream_t st) {
  using K = CK<D, F, H>;
  const int nres = resident_clusters<D, F, H>();
  if (nres < K::NC) {
    set_error("unsupported: this device holds %d clusters of %d CTAs at a time, the cluster step kernel needs %d", nres, CL, K::NC);
    return W2VS_UNSUPPORTED;
  }
  cudaLaunchConfig_t lc; cudaLaunchAttribute attr[2];
  launch_config<D, F, H>(lc, attr, st);
  cudaError_t e = cudaLaunchKernelEx(&lc, stream_cluster_kernel<D, F, H>, a);
  if (e != cudaSuccess) { set_error("stream_cluster_kernel launch: %s", cudaGetErrorString(e)); cudaGetLastError(); return W2VS_CUDA_ERROR; }
  W2VS_CHECK_LAUNCH("stream_cluster_kernel");
  return W2VS_OK;
}

template <int D, int F, int H>
w2vs_status_t pack_t(const ClusterPackArgs& p, cudaStream_t st) {
  pack_cluster_kernel<D, F, H><<<CL * CK<D, F, H>::NC, 256, 0, st>>>((const bf16*)p.wqkv, (const bf16*)p.wo, (const bf16*)p.w1, (const bf16*)p.w2, (uint8_t*)p.dst);
  W2VS_CHECK_LAUNCH("pack_cluster_kernel");
  pack_fold_kernel<D, F, H><<<(3 * D + F + 7) / 8, 256, 0, st>>>((const bf16*)p.wqkv, (const bf16*)p.w1, p.ln1_w, p.ln1_b, p.bqkv, p.ln2_w, p.ln2_b, p.b1, (uint8_t*)p.dst);
  W2VS_CHECK_LAUNCH("pack_fold_kernel");
  return W2VS_OK;
}

}  // namespace

static int g_cluster_trace_on = 0;    // process-wide debug switch (w2vs_debug_cluster_trace with n < 0 toggles it)

// The shapes this kernel is instantiated for: the released large model and the tiny model of the parity tests.
#define W2VS_CLUSTER_SHAPES(X) X(1024, 4096, 16) X(128, 256, 2)

size_t stream_cluster_layer_bytes(const w2vs_config* cfg) {
  if (!stream_cluster_model(cfg)) return 0;
#define X(D_, F_, H_) if (cfg->embed_dim == D_ && cfg->ffn_dim == F_ && cfg->heads == H_) return CK<D_, F_, H_>::LAYER_B;
  W2VS_CLUSTER_SHAPES(X)
#undef X
  return 0;
}

bool stream_cluster_applicable(const w2vs_config* cfg, int B, int ntok) {
  if (stream_cluster_layer_bytes(cfg) == 0 || B != 1 || ntok < 1 || ntok > ROWS) return false;
#define X(D_, F_, H_) if (cfg->embed_dim == D_ && cfg->ffn_dim == F_ && cfg->heads == H_) return resident_clusters<D_, F_, H_>() >= CK<D_, F_, H_>::NC;
  W2VS_CLUSTER_SHAPES(X)
#undef X
  return false;
}

w2vs_status_t launch_pack_cluster(const w2vs_config* cfg, const ClusterPackArgs& p, cudaStream_t st) {
#define X(D_, F_, H_) if (cfg->embed_dim == D_ && cfg->ffn_dim == F_ && cfg->heads == H_) return pack_t<D_, F_, H_>(p, st);
  W2VS_CLUSTER_SHAPES(X)
#undef X
  set_error("unsupported: no cluster step kernel for this model shape");
  return W2VS_UNSUPPORTED;
}

w2vs_status_t launch_stream_cluster(const StreamFusedArgs& h, cudaStream_t st) {
  const w2vs_config* cfg = h.cfg;
  W2VS_REQUIRE(stream_cluster_applicable(cfg, h.B, h.ntok), "cluster incremental step: configuration not supported");
  ClArgs a{};
  a.W = reinterpret_cast<const uint8_t*>(h.W);
  const LayerW& l0 = h.wl->layer0;
  a.wc = l0.wc; a.bqkv = l0.bqkv; a.bo = l0.bo; a.ln1_w = l0.ln1_w; a.ln1_b = l0.ln1_b;
  a.b1 = l0.b1; a.b2 = l0.b2; a.ln2_w = l0.ln2_w; a.ln2_b = l0.ln2_b;
  a.layer_stride = h.wl->layer_stride; a.enc_ln_w = h.wl->enc_ln_w; a.enc_ln_b = h.wl->enc_ln_b; a.sin_table = h.wl->sin_table;
  a.layers = cfg->layers; a.ntok = h.ntok; a.n_main = h.n_main; a.f0 = h.f0;
  a.feats = h.feats; a.X = h.R;
  a.kv = (bf16*)h.kv; a.kv_layer_elems = h.kv_layer_elems;
  a.out = (bf16*)h.out; a.bar = h.bar;
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  a.trace = g_cluster_trace_on;
#define X(D_, F_, H_) if (cfg->embed_dim == D_ && cfg->ffn_dim == F_ && cfg->heads == H_) return launch_t<D_, F_, H_>(a, st);
  W2VS_CLUSTER_SHAPES(X)
#undef X
  set_error("unsupported: no cluster step kernel for this model shape");
  return W2VS_UNSUPPORTED;
}

void debug_cluster_trace_enable(int on) { g_cluster_trace_on = on; }

w2vs_status_t debug_read_cluster_trace(unsigned long long* out, int n) {
  if (n > 64 * 32) n = 64 * 32;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_cluster_trace, (size_t)n * 8);
  if (e != cudaSuccess) { set_error("read g_cluster_trace: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  return W2VS_OK;
}

w2vs_status_t debug_read_cluster_fault(int* out) {
  int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_cluster_fault, sizeof(int));
  if (e != cudaSuccess) { set_error("read g_cluster_fault: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  *out = v;
  return W2VS_OK;
}

}  // namespace w2vs
