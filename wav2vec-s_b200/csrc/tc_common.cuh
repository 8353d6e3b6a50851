// tcgen05 / TMEM / TMA / mbarrier inline-PTX helpers shared by the sm_100a tensor-core kernels
// (k_gemm_tc2.cu, k_attn_tc.cu).  Each including .cu defines its own fault flag via W2VS_TC_FAULT_FLAG.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace w2vs {
namespace tc {

constexpr unsigned long long WAIT_TIMEOUT_NS = 4000000000ull;  // fail loudly instead of wedging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier that may live in the peer CTA (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// CTA-scope wait (the default).  NB: an .acquire.cluster wait makes ptxas emit CCTL.IVALL (L1 invalidate)
// after every probe, so the cluster-scope variant below is reserved for barriers that peer CTAs arrive on.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Blocking probes: with a suspend-time hint the hardware parks the thread until the phase completes (or the hint
// expires) instead of returning after a few cycles.  A spinning probe loop is not free: ncu attributed 24 % of the
// attention kernel's issued instructions to it (the single-thread MMA / TMA roles wait most of the time and share
// an SM sub-partition's issue port with the softmax / epilogue warps they are waiting for).
constexpr uint32_t WAIT_SUSPEND_NS = 100000;                                  // per probe
constexpr uint32_t WAIT_MAX_PROBES = (uint32_t)(WAIT_TIMEOUT_NS / WAIT_SUSPEND_NS) + 16;
__device__ __forceinline__ bool mbar_try_wait_park(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(WAIT_SUSPEND_NS) : "memory");
  return ok != 0;
}
// false on timeout (a bug / bad descriptor): callers abandon their loops so the kernel ends instead of wedging.
// Spinning flavour: lowest wake-up latency, for the roles on a pipeline's critical path (TMA producer, MMA issuer,
// GEMM epilogue).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfff) == 0 && global_ns() - t0 > WAIT_TIMEOUT_NS) {
      atomicExch(W2VS_TC_FAULT_FLAG, 1);
      return false;
    }
  }
  return true;
}
// Parking flavour: the probe carries a suspend-time hint, so a waiting warp stops competing for its sub-partition's
// issue port (measured in the attention kernel: a quarter of all issued instructions were barrier probes).
__device__ __forceinline__ bool mbar_wait_park(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  for (uint32_t probes = 0; probes < WAIT_MAX_PROBES; ++probes)
    if (mbar_try_wait_park(bar, parity)) return true;
  atomicExch(W2VS_TC_FAULT_FLAG, 1);
  return false;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0xfff) == 0 && global_ns() - t0 > WAIT_TIMEOUT_NS) {
      atomicExch(W2VS_TC_FAULT_FLAG, 1);
      return false;
    }
  }
  return true;
}
// 2-CTA TMA load: data lands in this CTA's smem, completion bytes are counted on `bar` (a shared::cluster
// address, here always the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(x), "r"(y), "r"(z) : "memory");
}
// out[tile] += smem tile, the fp32 additions done by the L2 reduction units (in-place residual update)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// One thread of the (converged) warp.  A single-thread role must be entered through this and not through `lane == 0`:
// inside an elect.sync region ptxas knows that one thread is active and issues UTCHMMA / UTMALDG with their operands
// moved to uniform registers once; behind `lane == 0` it wraps EVERY such instruction in a loop of ELECT +
// R2UR.BROADCAST + BRA.U.ANY (~100 cycles per instruction: the attention kernel's MMA thread took 500 cycles to issue
// four MMAs and paced the whole kernel).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// completion of all prior MMAs of this thread -> arrive on the barrier at this smem offset in both CTAs
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                                // leading byte offset  (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset   8 rows x 128 B
  d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld_chunk<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
template <> __device__ __forceinline__ void tmem_ld_chunk<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


// ---- host: tensor-map encoding through the driver entry point (no -lcuda link dependency) ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
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

inline w2vs_status_t make_map(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base,
                              uint64_t inner, uint64_t rows, uint64_t row_stride_elems, uint32_t box_inner,
                              uint32_t box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return W2VS_CUDA_ERROR; }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (inner=%llu rows=%llu stride=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_elems, box_inner,
              box_rows);
    return W2VS_CUDA_ERROR;
  }
  return W2VS_OK;
}

// 3-D map over column blocks of one row-major matrix: dimensions {inner (columns of a block), block, rows} with
// strides {block_stride, row_stride} -- ascending, as a column block is narrower than a row; boxes {box_inner, 1,
// box_rows} land in shared memory exactly like a 2-D {box_inner, box_rows} box, and are clipped at `inner` columns.
inline w2vs_status_t make_map3(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base,
                               uint64_t inner, uint64_t blocks, uint64_t rows, uint64_t block_stride_elems,
                               uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_rows,
                               CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return W2VS_CUDA_ERROR; }
  cuuint64_t dims[3] = {inner, blocks, rows};
  cuuint64_t strides[2] = {block_stride_elems * (uint64_t)elem_bytes, row_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box_inner, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed: %d (inner=%llu blocks=%llu rows=%llu)", (int)r,
              (unsigned long long)inner, (unsigned long long)blocks, (unsigned long long)rows);
    return W2VS_CUDA_ERROR;
  }
  return W2VS_OK;
}

using ::w2vs::num_sms;

}  // namespace tc
}  // namespace w2vs
