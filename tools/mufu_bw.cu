// Development aid: sustained throughput of MUFU.EX2 (and of F2FP bf16x2 packs, alone and mixed with it) per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_bw tools/mufu_bw.cu && ./tools/mufu_bw
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 3) { uint32_t h = __float_as_uint(x[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h)); x[i] = __uint_as_float(h | 0x80008000u); }
      if (MODE == 4) { uint32_t h = __float_as_uint(x[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h)); x[i] = __uint_as_float(h | 0x80008000u); }
      if (MODE == 1 || MODE == 2) { if (i & 1) { uint32_t p; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[i]), "f"(x[i - 1])); acc ^= p; } }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, int ctas_per_sm) {
  int nsm = 148, iters = 4096;
  float* out; long long* cyc;
  cudaMalloc(&out, (size_t)nsm * ctas_per_sm * threads * 4); cudaMalloc(&cyc, (size_t)nsm * ctas_per_sm * 8);
  k<MODE><<<nsm * ctas_per_sm, threads>>>(out, cyc, iters);
  k<MODE><<<nsm * ctas_per_sm, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<nsm * ctas_per_sm, threads>>>(out, cyc, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1]; cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("  [%.3f ms] ", ms);
  double mufu = (MODE == 1 ? 0.0 : (MODE >= 3 ? 16.0 : 8.0)) * iters * threads * ctas_per_sm, pack = (MODE == 0 ? 0.0 : 4.0) * iters * threads * ctas_per_sm;
  printf("%-28s %4d threads x %d CTAs/SM: %lld cycles, ex2 %.2f lanes/clk/SM, pack %.2f lanes/clk/SM\n", name, threads, ctas_per_sm, h[0], mufu / h[0], pack / h[0]);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2 only", 128, 1); run<0>("ex2 only", 512, 1); run<0>("ex2 only", 1024, 1); run<0>("ex2 only", 128, 2); run<0>("ex2 only", 128, 3);
  run<0>("ex2 only", 128, 4); run<0>("ex2 only", 512, 2); run<0>("ex2 only", 512, 4); run<0>("ex2 only", 32, 4); run<0>("ex2 only", 32, 8); run<0>("ex2 only", 64, 8);
  run<1>("bf16x2 pack only", 512, 2);
  run<2>("ex2 + pack", 512, 2);
  run<3>("ex2.f16x2 (elements)", 512, 2);
  run<4>("ex2.bf16x2 (elements)", 512, 2);
  return 0;
}
