// One conv block of the feature extractor on the new rows of a decision step (incremental mode, one stream, bf16):
//     carry + new rows  ->  Conv1d(C, C, k, stride s)  ->  + bias  ->  LayerNorm(C)  ->  GELU  ->  bf16 rows
// as ONE launch (wav2vec2.py:703-746: the blocks after the first; the streaming schedule of stream.cu).  The operator
// chain spends three launches per block on this -- a copy that puts the carried rows in front of the new ones, the
// tcgen05 GEMM (15 us for a few hundred rows: 24 dependent TMA stages and no programmatic launch) and the
// LayerNorm + GELU pass -- 130 us per decision step for arithmetic that takes a few microseconds.  Here:
//   * a cluster of 8 CTAs owns 32 output rows; CTA j computes the C/8 output channels [j C/8, (j+1) C/8) of those
//     rows: the input rows it needs (2 x 31 + k of them, read from the carry buffer or from the new rows -- the
//     concatenation is never materialised) sit in shared memory, its C/8 weight rows stream through a three-slot
//     cp.async ring in K chunks of 256 (four of them requested and the rest prefetched into L2 before
//     griddepcontrol.wait: they do not depend on the predecessor, and the layer kernel of the previous step has pushed
//     them out of L2);
//   * mma.sync.m16n8k16 over the im2col view (output row r, tap j, channel ci = input row s r + j, channel ci);
//   * LayerNorm over the C channels of a row spans the 8 CTAs: per-CTA mean and centred sum of squares, exchanged
//     through distributed shared memory (st.async + byte-counting mbarrier) and combined with Chan's formula;
//     normalisation and GELU on the fp32 accumulators (the chain rounds the pre-norm activation to bf16 first);
//   * the rows the next step has to see again (n_in - s n_out of them) are copied to the other carry buffer.
// Measured (large model, one stream, globaltimer stamps of CTA 0): 8.4 us from block to block -- input rows 1.6, the
// product 4.6 (k = 3; 48 ns per k step: shared-memory reads of the operand fragments), statistics and exchange 1.0,
// store and exit 1.1 -- against ~18 us for the three launches of the chain; the blocks of a step are launched as
// programmatic dependents and sit in griddepcontrol.wait with their first weight chunks loaded.
#include <cuda.h>
#include <math.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {
__device__ int g_conv_step_fault = 0;
}
#define W2VS_TC_FAULT_FLAG (&::w2vs::g_conv_step_fault)
#include "tc_common.cuh"

namespace w2vs {
namespace {
using namespace tc;

constexpr int CS_CL = 8, CS_WARPS = 8, CS_THREADS = 32 * CS_WARPS, CS_RT = 32, CS_SLOTS = 4;

template <int C>
struct CSK {
  static constexpr int CPC = C / CS_CL;                  // output channels per CTA
  static constexpr int NT = CPC / 8;                     // n tiles per CTA (one per warp)
  static constexpr int KC = C >= 256 ? 256 : C;          // K chunk (divides C: a chunk lies inside one tap)
  static constexpr int PX = C * 2 + 16, PW = KC * 2 + 16;
  static constexpr int XROWS = 2 * (CS_RT - 1) + 3;      // input rows of a tile for stride 2, k <= 3
  // the input rows lie in two regions, even and odd (relative to the tile's first row): output row r, tap j reads
  // input row 2 r + j = row r + j / 2 of region j % 2, so consecutive output rows are one pitch apart (16 mod 128
  // bytes: ldmatrix without bank conflicts; with one region the stride would be 32 mod 128, a two-way conflict)
  static constexpr int XREG = (XROWS + 1) / 2 * PX;
  static constexpr int NTW = NT >= 4 ? NT / 4 : 1;       // n tiles per warp: warp = (m tile, group of NTW n tiles)
  static constexpr int S_X = 0, S_W = (2 * XREG + 127) / 128 * 128, W_B = (CPC * PW + 127) / 128 * 128;
  static constexpr int S_PART = S_W + CS_SLOTS * W_B;    // [2][4 column groups (8 slots)][32 rows] fp32
  static constexpr int S_RX = S_PART + 2 * CS_WARPS * CS_RT * 4;     // [8 CTAs][32 rows][2] fp32, written by the peers
  static constexpr int S_FIN = S_RX + CS_CL * CS_RT * 8;             // [32][2] mean, rstd
  static constexpr int S_BAR = S_FIN + CS_RT * 8;
  static constexpr int S_END = S_BAR + 16;
  static_assert(C % 64 == 0 && NT >= 1 && NT <= CS_WARPS, "channels");
  static_assert(S_END + 128 <= 232448, "shared memory");
};

struct CsArgs {
  const bf16* carry; int n_carry;      // rows carried from the previous step (start of the valid rows)
  const bf16* fresh; int n_fresh;      // this step's new input rows
  const bf16* W;                       // [C][k * C] K-major, tap-major inside K
  const float *bias, *gamma, *beta;    // bias may be NULL
  bf16* out; int n_out;                // [n_out][C]
  bf16* carry_out;                     // rows [s * n_out, n_carry + n_fresh) of the concatenation
  int k, s;
};

__device__ __forceinline__ void cs_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cs_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cs_ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void cs_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cs_st_async_v2f(uint32_t addr, float a, float b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1,%2}, [%3];"
               ::"r"(addr), "f"(a), "f"(b), "r"(mbar) : "memory");
}

template <int C>
__global__ void __launch_bounds__(CS_THREADS, 1)
conv_step_kernel(const __grid_constant__ CsArgs a) {
  using K = CSK<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sm = smem_raw + (sb - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int rank = (int)cluster_ctarank(), tile = blockIdx.x / CS_CL;
  const int r_begin = tile * CS_RT, rows = min(CS_RT, a.n_out - r_begin);
  const int n_in = a.n_carry + a.n_fresh, Ktot = a.k * C, nchunks = Ktot / K::KC;
  const uint32_t bar = sb + K::S_BAR;
  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, CS_CL * CS_RT * 8);
  }
  // ---- weight rows of this CTA, K chunk c -> ring slot c % 3 (16-byte cp.async: the rows are k*C*2 bytes apart) ----
  const bf16* wrow0 = a.W + (size_t)(rank * K::CPC) * Ktot;
  auto issue_chunk = [&](int c) {
    if (c < nchunks) {
      constexpr int VPR = K::KC * 2 / 16;              // 16-byte vectors per row of a chunk
      const uint32_t slot = sb + K::S_W + (uint32_t)(c % CS_SLOTS) * K::W_B;
      for (int i = tid; i < K::CPC * VPR; i += CS_THREADS) {
        const int n = i / VPR, v = i - n * VPR;
        cs_cp_async16(slot + (uint32_t)n * K::PW + v * 16, wrow0 + (size_t)n * Ktot + (size_t)c * K::KC + v * 8);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");     // (an empty group keeps the count uniform)
  };
#pragma unroll
  for (int c = 0; c < CS_SLOTS; ++c) issue_chunk(c);
  // the chunks that do not fit the ring yet: into L2 now (DRAM latency would otherwise sit between the MMAs)
  for (int i = tid; i < K::CPC * ((Ktot - CS_SLOTS * K::KC) * 2 / 128); i += CS_THREADS) {
    const int lpr = (Ktot - CS_SLOTS * K::KC) * 2 / 128, n = i / lpr, v = i - n * lpr;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(wrow0 + (size_t)n * Ktot + (size_t)CS_SLOTS * K::KC + v * 64));
  }
  cluster_sync();                      // barrier inits of all CTAs before any remote store
  pdl_wait();                          // the new rows come from the previous kernel

  // ---- input rows of the tile: concatenation row x = s (r_begin + r) + j, from the carry or from the new rows.
  //      cp.async as well: sixteen plain loads per thread, each followed by its shared-memory store, are sixteen
  //      dependent L2 round trips (the compiler keeps the order: it cannot prove the buffers distinct).
  const int x0 = a.s * r_begin, xrows = a.s * (rows - 1) + a.k;
  {
    constexpr int VPR = C * 2 / 16;
    for (int i = tid; i < xrows * VPR; i += CS_THREADS) {
      const int xr = i / VPR, v = i - xr * VPR, x = x0 + xr;
      const uint32_t off = (uint32_t)(K::S_X + (xr & 1) * K::XREG + (xr >> 1) * K::PX + v * 16);
      if (x < n_in) cs_cp_async16(sb + off, (x < a.n_carry ? a.carry + (size_t)x * C : a.fresh + (size_t)(x - a.n_carry) * C) + v * 8);
      else *reinterpret_cast<uint4*>(sm + off) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // rows the next step sees again (at most k - 1): one CTA copies them to the other carry buffer
    if (blockIdx.x == 0) {
      const int c0 = a.s * a.n_out, nc = n_in - c0;
      for (int i = tid; i < nc * VPR; i += CS_THREADS) {
        const int xr = i / VPR, v = i - xr * VPR, x = c0 + xr;
        const bf16* src = x < a.n_carry ? a.carry + (size_t)x * C : a.fresh + (size_t)(x - a.n_carry) * C;
        *reinterpret_cast<uint4*>(a.carry_out + (size_t)xr * C + v * 8) = *reinterpret_cast<const uint4*>(src + v * 8);
      }
    }
  }

  // ---- the product: warp = (m tile w % 2, n tiles [NTW (w / 2), NTW (w / 2 + 1)) of the CTA's channels) ----
  float acc[K::NTW][4] = {};
  const int mt = warp & 1, cg = warp >> 1;             // m tile, column group
  const bool mma_warp = cg * K::NTW < K::NT && mt * 16 < rows;
  for (int c = 0; c < nchunks; ++c) {
    // groups so far: weight chunks 0..3, the input rows, then one per iteration; the first wait covers the input rows
    if (c == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.wait_group %0;" ::"n"(CS_SLOTS - 1) : "memory");
    __syncthreads();                   // chunk c (and, the first time, the input rows) visible to every warp
    if (mma_warp) {
      const int tap = (c * K::KC) / C, ci0 = (c * K::KC) % C;
      const uint32_t a_lane = sb + K::S_X + (uint32_t)((tap & 1) * K::XREG) + (uint32_t)(mt * 16 + (lane & 15) + (tap >> 1)) * K::PX +
                              (uint32_t)(ci0 * 2) + (uint32_t)(lane >> 4) * 16;
      const uint32_t b_lane = sb + K::S_W + (uint32_t)(c % CS_SLOTS) * K::W_B + (uint32_t)(cg * K::NTW * 8 + (lane & 7)) * K::PW +
                              (uint32_t)((lane >> 3) & 1) * 16;
      // fragments of k step ks + 1 are requested before the MMAs of step ks (the ldmatrix statements stay in order)
      uint32_t af[2][4], bf[2][K::NTW][2];
      cs_ldsm_x4(a_lane, af[0]);
#pragma unroll
      for (int nt = 0; nt < K::NTW; ++nt) cs_ldsm_x2(b_lane + (uint32_t)(nt * 8) * K::PW, bf[0][nt]);
#pragma unroll
      for (int ks = 0; ks < K::KC / 16; ++ks) {
        const int cur = ks & 1;
        if (ks + 1 < K::KC / 16) {
          cs_ldsm_x4(a_lane + (ks + 1) * 32, af[cur ^ 1]);
#pragma unroll
          for (int nt = 0; nt < K::NTW; ++nt) cs_ldsm_x2(b_lane + (uint32_t)(nt * 8) * K::PW + (ks + 1) * 32, bf[cur ^ 1][nt]);
        }
#pragma unroll
        for (int nt = 0; nt < K::NTW; ++nt) cs_mma(acc[nt], af[cur], bf[cur][nt]);
      }
    }
    __syncthreads();                   // every warp is past its reads of the slot
    issue_chunk(c + CS_SLOTS);
  }
  // ---- bias, LayerNorm statistics over the C channels of a row (4 column groups x 8 CTAs), GELU, store ----
  const int ch0 = rank * K::CPC + cg * K::NTW * 8 + 2 * q;      // this lane's first channel pair (mma warps)
  float* part = reinterpret_cast<float*>(sm + K::S_PART);
  constexpr int NCG = K::NT / K::NTW;                  // column groups that exist
  float v[K::NTW][4];
  const int r0 = mt * 16 + g, r1 = r0 + 8;             // this lane's rows of the tile
  {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < K::NTW; ++nt) {
      float2 bia = make_float2(0.f, 0.f);
      if (mma_warp && a.bias != nullptr) bia = *reinterpret_cast<const float2*>(a.bias + ch0 + nt * 8);
      v[nt][0] = acc[nt][0] + bia.x; v[nt][1] = acc[nt][1] + bia.y; v[nt][2] = acc[nt][2] + bia.x; v[nt][3] = acc[nt][3] + bia.y;
      s0 += v[nt][0] + v[nt][1]; s1 += v[nt][2] + v[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    if (q == 0 && cg < NCG) { part[cg * CS_RT + r0] = mma_warp ? s0 : 0.f; part[cg * CS_RT + r1] = mma_warp ? s1 : 0.f; }
  }
  __syncthreads();
  {
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int w = 0; w < NCG; ++w) { m0 += part[w * CS_RT + r0]; m1 += part[w * CS_RT + r1]; }
    m0 *= 1.0f / K::CPC; m1 *= 1.0f / K::CPC;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < K::NTW; ++nt) {
      const float d0 = v[nt][0] - m0, d1 = v[nt][1] - m0, d2 = v[nt][2] - m1, d3 = v[nt][3] - m1;
      s0 += d0 * d0 + d1 * d1; s1 += d2 * d2 + d3 * d3;
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    if (q == 0 && cg < NCG) { part[(CS_WARPS + cg) * CS_RT + r0] = mma_warp ? s0 : 0.f; part[(CS_WARPS + cg) * CS_RT + r1] = mma_warp ? s1 : 0.f; }
  }
  __syncthreads();
  {
    // (row, destination CTA): this CTA's mean and centred sum of squares of the row over its C/8 channels
    const int r = tid >> 3, d = tid & 7;
    float s = 0.f, m2 = 0.f;
#pragma unroll
    for (int w = 0; w < NCG; ++w) { s += part[w * CS_RT + r]; m2 += part[(CS_WARPS + w) * CS_RT + r]; }
    cs_st_async_v2f(mapa(sb + K::S_RX + (uint32_t)((rank * CS_RT + r) * 8), (uint32_t)d), s * (1.0f / K::CPC), m2, mapa(bar, (uint32_t)d));
  }
  bool ok = mbar_wait(bar, 0);
  if (tid < CS_RT) {
    const float2* rx = reinterpret_cast<const float2*>(sm + K::S_RX);
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < CS_CL; ++j) mean += rx[j * CS_RT + tid].x;
    mean *= 1.0f / CS_CL;
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < CS_CL; ++j) { const float2 p = rx[j * CS_RT + tid]; const float dv = p.x - mean; m2 += p.y + (float)K::CPC * dv * dv; }
    reinterpret_cast<float2*>(sm + K::S_FIN)[tid] = make_float2(mean, 1.0f / sqrtf(m2 * (1.0f / C) + 1e-5f));
  }
  __syncthreads();
  if (mma_warp && ok) {
    const float2* fin = reinterpret_cast<const float2*>(sm + K::S_FIN);
    const float2 st0 = fin[r0], st1 = fin[r1];
#pragma unroll
    for (int nt = 0; nt < K::NTW; ++nt) {
      const int ch = ch0 + nt * 8;
      const float2 gm = *reinterpret_cast<const float2*>(a.gamma + ch), bt = *reinterpret_cast<const float2*>(a.beta + ch);
      if (r0 < rows)
        *reinterpret_cast<uint32_t*>(a.out + (size_t)(r_begin + r0) * C + ch) =
            pack_bf16x2(gelu_tanh((v[nt][0] - st0.x) * st0.y * gm.x + bt.x), gelu_tanh((v[nt][1] - st0.x) * st0.y * gm.y + bt.y));
      if (r1 < rows)
        *reinterpret_cast<uint32_t*>(a.out + (size_t)(r_begin + r1) * C + ch) =
            pack_bf16x2(gelu_tanh((v[nt][2] - st1.x) * st1.y * gm.x + bt.x), gelu_tanh((v[nt][3] - st1.x) * st1.y * gm.y + bt.y));
    }
  }
  // (no cluster barrier at the end: a CTA that has passed its mbarrier wait has received every remote store it will
  //  ever get, and the peers it stored to are held in their own wait until those stores have landed)
}

template <int C>
w2vs_status_t launch_c(const CsArgs& a, cudaStream_t st) {
  using K = CSK<C>;
  auto kern = conv_step_kernel<C>;
  const size_t smem = K::S_END + 128;
  static PerDeviceOnce once;
  bool& done = once.here();
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv_step smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    done = true;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(CS_CL * ((a.n_out + CS_RT - 1) / CS_RT))); lc.blockDim = dim3(CS_THREADS);
  lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = g_pdl_on ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, kern, a);
  if (e != cudaSuccess) { set_error("conv_step_kernel launch: %s", cudaGetErrorString(e)); cudaGetLastError(); return W2VS_CUDA_ERROR; }
  W2VS_CHECK_LAUNCH("conv_step_kernel");
  return W2VS_OK;
}

// ---- feature LayerNorm + post_extract_proj + append to the stream's frame buffer, one launch ----------------------
// (wav2vec2.py:566-569 on the new frames of a decision step; the chain: LayerNorm launch, weight-streaming GEMM, row
// copy).  One CTA per 8 output columns, 8 warps split K: every CTA normalises the (at most 64) new rows itself into
// shared memory -- 16 KB of L2 reads -- while its 8 weight rows, requested before griddepcontrol.wait, are in flight;
// mma.sync.m16n8k16 with the k index of a 32-wide block permuted identically for both operands (lane (g, q) takes the
// 8 consecutive elements 32 b + 8 q ..: 16-byte accesses, no ldmatrix), fixed-order reduction of the 8 partial sums.
constexpr int FP_MAX_MT = 4;

template <int MT, int KB>      // m16 tiles of rows; 32-wide k blocks per warp (K = 256 KB)
__global__ void __launch_bounds__(CS_THREADS)
feat_proj_kernel(const bf16* __restrict__ x, int rows, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const bf16* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out, int N) {
  constexpr int K = 256 * KB, PA = K * 2 + 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sa = smem_raw;                                             // normalised rows [MT * 16][PA] bf16
  float* part = reinterpret_cast<float*>(smem_raw + MT * 16 * PA);    // [8 warps][MT * 16][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int n0 = blockIdx.x * 8;
  pdl_launch_dependents();
  const int kbeg = warp * (K / CS_WARPS) + 8 * q;
  uint4 wv[KB];
#pragma unroll
  for (int b = 0; b < KB; ++b) wv[b] = __ldg(reinterpret_cast<const uint4*>(W + (size_t)(n0 + g) * K + kbeg + 32 * b));
  pdl_wait();
  // LayerNorm: one warp per row, the row in registers (K / 32 elements per lane, 16-byte loads)
  constexpr int VPL = K / 256;                  // uint4 per lane and row
  for (int r = warp; r < MT * 16; r += CS_WARPS) {
    uint4 u[VPL];
    float v[VPL][8], s = 0.f;
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        u[j] = *reinterpret_cast<const uint4*>(x + (size_t)r * K + (j * 32 + lane) * 8);
        const float2 a0 = unpack_bf16x2(u[j].x), a1 = unpack_bf16x2(u[j].y), a2 = unpack_bf16x2(u[j].z), a3 = unpack_bf16x2(u[j].w);
        v[j][0] = a0.x; v[j][1] = a0.y; v[j][2] = a1.x; v[j][3] = a1.y; v[j][4] = a2.x; v[j][5] = a2.y; v[j][6] = a3.x; v[j][7] = a3.y;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[j][e];
      }
    }
    const float mean = warp_sum(s) * (1.0f / K);
    float qq = 0.f;
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[j][e] - mean; qq = fmaf(d, d, qq); }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(qq) * (1.0f / K) + 1e-5f);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (r < rows) {
        const int c = (j * 32 + lane) * 8;
        float gg[8], bb[8], y[8];
        load8(gamma + c, gg);
        load8(beta + c, bb);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (v[j][e] - mean) * rstd * gg[e] + bb[e];
        o = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      }
      *reinterpret_cast<uint4*>(sa + r * PA + (j * 32 + lane) * 16) = o;
    }
  }
  __syncthreads();
  float acc[MT][4] = {};
#pragma unroll
  for (int b = 0; b < KB; ++b) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint4 lo = *reinterpret_cast<const uint4*>(sa + (mt * 16 + g) * PA + (kbeg + 32 * b) * 2);
      const uint4 hi = *reinterpret_cast<const uint4*>(sa + (mt * 16 + g + 8) * PA + (kbeg + 32 * b) * 2);
      const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y}, a1[4] = {lo.z, hi.z, lo.w, hi.w};
      const uint32_t b0[2] = {wv[b].x, wv[b].y}, b1[2] = {wv[b].z, wv[b].w};
      cs_mma(acc[mt], a0, b0);
      cs_mma(acc[mt], a1, b1);
    }
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float* p = part + ((size_t)warp * MT * 16 + mt * 16 + g) * 8 + 2 * q;
    p[0] = acc[mt][0]; p[1] = acc[mt][1]; p[64] = acc[mt][2]; p[65] = acc[mt][3];
  }
  __syncthreads();
  for (int o = tid; o < rows * 8; o += CS_THREADS) {
    const int r = o >> 3, c = o & 7;
    float v = bias != nullptr ? bias[n0 + c] : 0.f;
#pragma unroll
    for (int w = 0; w < CS_WARPS; ++w) v += part[((size_t)w * MT * 16 + r) * 8 + c];      // fixed order
    out[(size_t)r * N + n0 + c] = v;
  }
}

template <int KB>
w2vs_status_t launch_fp(const FeatProjArgs& a, cudaStream_t st) {
  const int mt = (a.rows + 15) / 16;
  constexpr int K = 256 * KB;
  const size_t smem = (size_t)mt * 16 * (K * 2 + 16) + (size_t)CS_WARPS * mt * 16 * 8 * 4;
  const dim3 grid((unsigned)(a.N / 8));
#define W2VS_FP_CASE(MT_)                                                                                         \
  {                                                                                                               \
    static PerDeviceOnce once;                                                                                    \
    if (!once.here()) {                                                                                           \
      cudaFuncSetAttribute(feat_proj_kernel<MT_, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FP_MAX_MT * 16 * (K * 2 + 16) + CS_WARPS * FP_MAX_MT * 16 * 8 * 4)); \
      once.here() = true;                                                                                         \
    }                                                                                                             \
    launch_pdl(feat_proj_kernel<MT_, KB>, grid, dim3(CS_THREADS), smem, st, (const bf16*)a.x, a.rows, a.gamma, a.beta, \
               (const bf16*)a.W, a.bias, a.out, a.N);                                                             \
  }
  switch (mt) {
    case 1: W2VS_FP_CASE(1) break;
    case 2: W2VS_FP_CASE(2) break;
    case 3: W2VS_FP_CASE(3) break;
    default: W2VS_FP_CASE(4) break;
  }
#undef W2VS_FP_CASE
  W2VS_CHECK_LAUNCH("feat_proj_kernel");
  return W2VS_OK;
}

}  // namespace

bool feat_proj_applicable(const FeatProjArgs& a) {
  return (a.K == 512 || a.K == 256) && a.rows >= 1 && a.rows <= 16 * FP_MAX_MT && a.N % 8 == 0 && a.gamma != nullptr && a.beta != nullptr;
}

w2vs_status_t launch_feat_proj(const FeatProjArgs& a, cudaStream_t st) {
  W2VS_REQUIRE(feat_proj_applicable(a), "feature projection step: configuration not supported");
  return a.K == 512 ? launch_fp<2>(a, st) : launch_fp<1>(a, st);
}

bool conv_step_applicable(const ConvStepArgs& h) {
  return (h.C == 512 || h.C == 64) && h.k >= 2 && h.k <= 3 && h.s == 2 && h.n_out >= 1 && h.n_carry >= 0 && h.n_fresh >= 0 &&
         h.n_carry + h.n_fresh >= h.s * (h.n_out - 1) + h.k && h.gamma != nullptr && h.beta != nullptr;
}

w2vs_status_t launch_conv_step(const ConvStepArgs& h, cudaStream_t st) {
  W2VS_REQUIRE(conv_step_applicable(h), "conv step: configuration not supported");
  CsArgs a{};
  a.carry = (const bf16*)h.carry; a.n_carry = h.n_carry; a.fresh = (const bf16*)h.fresh; a.n_fresh = h.n_fresh;
  a.W = (const bf16*)h.W; a.bias = h.bias; a.gamma = h.gamma; a.beta = h.beta;
  a.out = (bf16*)h.out; a.n_out = h.n_out; a.carry_out = (bf16*)h.carry_out; a.k = h.k; a.s = h.s;
  return h.C == 512 ? launch_c<512>(a, st) : launch_c<64>(a, st);
}

}  // namespace w2vs
