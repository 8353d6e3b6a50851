// Development aid: TMEM read/write bandwidth and MUFU / FMNMX throughput micro-benchmarks (one CTA, clock64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tmem_bw tools/tmem_bw.cu && ./tools/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__global__ void k_tmem(int nwarps_active, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(base, r0); tmem_ld32(base + 32, r1); tmem_ld32(base + 64, r2); tmem_ld32(base + 96, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += r0[j] ^ r1[j] ^ r2[j] ^ r3[j];
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
// MUFU.EX2 / FMNMX3 / FFMA2 issue throughput with nw warps per SM (all on distinct or shared SMSPs)
template <int OP>
__global__ void k_alu(int iters, long long* out, float* sink, float seed) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = seed + j * 0.001f + threadIdx.x * 1e-6f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(x[(j + 1) & 15]));
      if (OP == 2) { uint64_t v; asm volatile("mov.b64 %0, {%1, %2}; fma.rn.f32x2 %0, %0, %0, %0; mov.b64 {%1, %2}, %0;" : "=l"(v), "+f"(x[j]), "+f"(x[(j + 8) & 15])); }
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[j]));
      if (OP == 4) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
      if (OP == 5) { uint32_t u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x[j]), "f"(x[(j + 1) & 15])); x[j] = __uint_as_float(u & 0x3f803f80u); }
      if (OP == 6) { uint32_t u; asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
                     asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x[j]), "f"(x[(j + 1) & 15])); x[(j + 3) & 15] = __uint_as_float(u & 0x3f803f80u); }
      if (OP == 7) { uint32_t a = __float_as_uint(x[j]) + 0x8000u, b = __float_as_uint(x[(j + 1) & 15]) + 0x8000u, u;
                     asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u) : "r"(a), "r"(b)); x[j] = __uint_as_float(u & 0x3f803f80u); }
      if (OP == 8) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
                     uint32_t a = __float_as_uint(x[j]) + 0x8000u, b = __float_as_uint(x[(j + 1) & 15]) + 0x8000u, u;
                     asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u) : "r"(a), "r"(b)); x[(j + 3) & 15] = __uint_as_float(u & 0x3f803f80u); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[threadIdx.x >> 5] = t1 - t0;
  float s = 0; for (int j = 0; j < 16; ++j) s += x[j];
  sink[threadIdx.x] = s;
}
// dependent-chain latency (one chain per thread): cycles per op
template <int OP>
__global__ void k_lat(int iters, long long* out, float* sink, float seed) {
  float a = seed + threadIdx.x * 1e-6f, b = seed * 0.5f, c = seed * 0.25f;
  uint64_t v;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(v));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
      if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  float x, y;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
  sink[threadIdx.x] = a + x + y;
}
int main() {
  long long* out; uint32_t* sink; cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 4096 * 4);
  long long h[32];
  const int iters = 200;
  for (int nw : {1, 4, 8}) {
    k_tmem<<<1, 256>>>(nw, iters, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("tmem ld 32x32b.x32 x4 per iter, %d warps: %.1f cycles/iter (16 KB per warp-iter) -> %.1f B/clk/SM  err=%s\n", nw,
           (double)h[0] / iters, 16384.0 * nw / ((double)h[0] / iters), cudaGetErrorString(cudaGetLastError()));
  }
  const char* names[] = {"ex2.approx", "max.f32", "fma.f32x2", "fma.f32", "tanh.approx", "cvt.bf16x2", "ex2+cvt", "iadd2+prmt", "ex2+iadd2prmt"};
  for (int op = 0; op < 9; ++op)
    for (int nthreads : {32, 128, 256}) {
      if (op == 0) k_alu<0><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 1) k_alu<1><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 2) k_alu<2><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 3) k_alu<3><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 4) k_alu<4><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 5) k_alu<5><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 6) k_alu<6><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 7) k_alu<7><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      if (op == 8) k_alu<8><<<1, nthreads>>>(1000, out, (float*)sink, 0.5f);
      cudaDeviceSynchronize();
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-12s %3d threads: %.2f cycles per warp-instruction per warp, %.1f thread-ops/clk/SM\n", names[op], nthreads,
             (double)h[0] / 16000.0, 16000.0 * nthreads / (double)h[0]);
    }
  const char* lnames[] = {"fma.f32", "fma.f32x2", "add.f32x2", "max.f32", "max3.f32", "ex2.approx"};
  for (int op = 0; op < 6; ++op) {
    if (op == 0) k_lat<0><<<1, 32>>>(1000, out, (float*)sink, 0.5f);
    if (op == 1) k_lat<1><<<1, 32>>>(1000, out, (float*)sink, 0.5f);
    if (op == 2) k_lat<2><<<1, 32>>>(1000, out, (float*)sink, 0.5f);
    if (op == 3) k_lat<3><<<1, 32>>>(1000, out, (float*)sink, 0.5f);
    if (op == 4) k_lat<4><<<1, 32>>>(1000, out, (float*)sink, 0.5f);
    if (op == 5) k_lat<5><<<1, 32>>>(1000, out, (float*)sink, 0.5f);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("latency %-12s: %.2f cycles per dependent op\n", lnames[op], (double)h[0] / 16000.0);
  }
  return 0;
}
