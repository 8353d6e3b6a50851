// bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA:
//     C[M,N] = A[M,K] . W[N,K]^T + bias  (+GELU)  (+fp32 residual)      fp32 accumulation in TMEM
//
// Covers every matrix product of the encoder in bf16 mode: the strided feature-extractor convs 1..6
// (as implicit GEMMs over the channels-last activation), post_extract_proj, QKV / out_proj and
// fc1(+GELU) / fc2(+residual)  (wav2vec2.py:725,568,950-973; multihead_attention.py:162-194).
//
// Structure (one persistent CTA per SM, 192 threads, static round-robin tile schedule):
//   warp 0   : TMA producer  - cp.async.bulk.tensor 2D loads of a 128xBK A tile and a BNxBK W tile per
//              stage into 128B-swizzled shared memory, mbarrier complete_tx signalling
//   warp 1   : MMA issuer    - one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//              K=16) BK/16 times per stage; tcgen05.commit releases the smem stage / publishes the
//              accumulator.  Also owns TMEM alloc/dealloc.
//   warps 2-5: epilogue      - tcgen05.ld 32x32b (one accumulator row per thread), bias/GELU/residual
//              in registers, 16-byte global stores.  Two accumulator stages (2 x BN TMEM columns) let
//              the epilogue of tile i overlap the MMAs of tile i+1.
//
// Implicit-GEMM view of Conv1d(C_in -> C_out, k, stride s) on activations [rows, C_in] (row-major):
// output row r needs the k*C_in contiguous inputs starting at row r*s.  The A tensor map describes the
// plain matrix [rows*1, a_row_len = s*C_in]; K index kk >= a_row_len wraps to the next row:
// (x, y) = (kk % a_row_len, r + kk / a_row_len).  Ordinary GEMMs have a_row_len >= K (no wrap).
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"

namespace w2vs {

__device__ int g_tc_fault = 0;  // set when a pipeline wait timed out (diagnostics; see w2vs_debug_fault)

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr unsigned long long WAIT_TIMEOUT_NS = 4000000000ull;  // 4 s: fail loudly instead of hanging the GPU

template <int BN> struct TileCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kBStageBytes = BN * BK * 2;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBStageBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // power of two for BN in {64,128,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Returns false on timeout (a bug or a bad descriptor); callers then abandon their loops so that the
// kernel terminates instead of wedging the device.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && global_ns() - t0 > WAIT_TIMEOUT_NS) {
      atomicExch(&g_tc_fault, 1);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major operand tile in 128B-swizzled smem: rows of 128 B, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                                // leading byte offset  bits [16,30) (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN, typename TC>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const float* __restrict__ bias, const float* residual, TC* C, int64_t ldc, int M, int N,
               int K, int a_row_len, int gelu) {
  using Cfg = TileCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + S * A_STAGE_BYTES;
  const uint32_t bars = smem_base + S * Cfg::kStageBytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * S, bar_tfull = bars + 16 * S, bar_tempty = bars + 16 * S + 16;
  const uint32_t tmem_slot = bars + 16 * S + 32;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + BM - 1) / BM;
  const int n_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!(ok = mbar_wait(bar_empty + 8 * stage, phase ^ 1))) break;
          mbar_expect_tx(bar_full + 8 * stage, Cfg::kStageBytes);
          const int kk = kb * BK;
          tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, bar_full + 8 * stage, kk % a_row_len, m0 + kk / a_row_len);
          tma_load_2d(sB + stage * Cfg::kBStageBytes, &tmB, bar_full + 8 * stage, kk, n0);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        if (!(ok = mbar_wait(bar_tempty + 8 * as, aphase ^ 1))) break;
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!(ok = mbar_wait(bar_full + 8 * stage, phase))) break;
          tc_fence_after();
          const uint32_t a_addr = sA + stage * A_STAGE_BYTES, b_addr = sB + stage * Cfg::kBStageBytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16(tmem_c, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(bar_empty + 8 * stage);  // smem stage reusable once these MMAs retire
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        tc_commit(bar_tfull + 8 * as);       // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      ok = mbar_wait(bar_tfull + 8 * as, aphase);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < M && col0 < N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // groups of 8 columns
            const int col = col0 + g * 8;
            if (col >= N) break;        // N % 8 == 0
            float v[8];
            const float4 b0 = bias ? *reinterpret_cast<const float4*>(bias + col) : make_float4(0, 0, 0, 0);
            const float4 b1 = bias ? *reinterpret_cast<const float4*>(bias + col + 4) : make_float4(0, 0, 0, 0);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = __uint_as_float(r[g * 8 + e]) + bb[e];
              if (gelu) v[e] = gelu_erf(v[e]);
            }
            if (residual) {
              float rr[8];
              load8(residual + (size_t)row * ldc + col, rr);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += rr[e];
            }
            store8(C + (size_t)row * ldc + col, v);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;   // resolved once; benign race (same value)
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

w2vs_status_t make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                       uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return W2VS_CUDA_ERROR; }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (inner=%llu rows=%llu stride=%llu)", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_elems);
    return W2VS_CUDA_ERROR;
  }
  return W2VS_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, typename TC>
w2vs_status_t launch_bn(const GemmArgs& g, cudaStream_t st) {
  using Cfg = TileCfg<BN>;
  alignas(64) CUtensorMap tmA, tmB;
  const int64_t a_row_len = g.lda;
  // A: plain matrix [a_rows, lda]; if K <= lda only the first K columns are addressed.
  const uint64_t a_inner = (uint64_t)(g.K <= a_row_len ? g.K : a_row_len);
  W2VS_TRY(make_map(&tmA, g.A, a_inner, (uint64_t)g.a_rows, (uint64_t)a_row_len, BM));
  W2VS_TRY(make_map(&tmB, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, BN));
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc smem attribute: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
    attr_done = true;
  }
  const int tiles = (int)(ceil_div64(g.M, BM) * ceil_div64(g.N, BN));
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc_kernel<BN, TC><<<grid, 192, Cfg::kSmemBytes, st>>>(
      tmA, tmB, g.bias, g.residual, (TC*)g.C, g.ldc, g.M, g.N, g.K, (int)a_row_len,
      (g.flags & W2VS_EPI_GELU) ? 1 : 0);
  W2VS_CHECK_LAUNCH("gemm_tc_kernel");
  return W2VS_OK;
}

template <typename TC>
w2vs_status_t launch_tc_typed(const GemmArgs& g, cudaStream_t st) {
  if (g.N % 256 == 0) return launch_bn<256, TC>(g, st);
  if (g.N % 128 == 0) return launch_bn<128, TC>(g, st);
  if (g.N % 64 == 0 && g.N < 256) return launch_bn<64, TC>(g, st);
  return launch_bn<256, TC>(g, st);
}

}  // namespace

w2vs_status_t launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
  W2VS_REQUIRE(g.dtype_ab == W2VS_BF16, "tcgen05 GEMM takes bf16 operands");
  W2VS_REQUIRE(g.K % 8 == 0 && g.N % 8 == 0, "GEMM needs K % 8 == 0 and N % 8 == 0");
  W2VS_REQUIRE(g.lda % 8 == 0 && g.ldc % 8 == 0, "GEMM leading dims must be multiples of 8");
  W2VS_REQUIRE(g.K <= g.lda || g.lda % BK == 0, "wrapped (conv) A rows need lda % 64 == 0");
  W2VS_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0, "GEMM operands must be 16-byte aligned");
  if (g.M <= 0) return W2VS_OK;
  return g.dtype_c == W2VS_F32 ? launch_tc_typed<float>(g, st) : launch_tc_typed<bf16>(g, st);
}

w2vs_status_t debug_read_tc_fault(int* out) {
  int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_tc_fault, sizeof(int));
  if (e != cudaSuccess) { set_error("read g_tc_fault: %s", cudaGetErrorString(e)); return W2VS_CUDA_ERROR; }
  *out = v;
  return W2VS_OK;
}

}  // namespace w2vs
