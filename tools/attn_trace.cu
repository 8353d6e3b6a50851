// Development aid (not part of the library): runs attn_tc_kernel once on a cfg3-shaped problem with clock64
// tracing of one heavy CTA, prints the per-tile event timeline of softmax warp 0 and of the MMA thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DW2VS_ATTN_TRACE -o attn_trace tools/attn_trace.cu \
//        wav2vec-s_b200/csrc/layout.cu && ./attn_trace
#include <vector>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "../wav2vec-s_b200/csrc/k_attn_tc.cu"

int main() {
  using namespace w2vs;
  const int B = 64, T2 = 1000, main_ctx = 16, rc = 8, heads = 16, D = 1024;
  const int M = T2 + (T2 / main_ctx) * rc;
  bf16 *qkv, *ctx; uint8_t* kp;
  if (cudaMalloc(&qkv, (size_t)B * M * 3 * D * 2 + (1 << 20)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&ctx, (size_t)B * M * D * 2);
  cudaMalloc(&kp, (size_t)B * M);
  cudaMemset(kp, 0, (size_t)B * M);
  std::vector<uint16_t> h((size_t)B * M * 3 * D);
  unsigned s = 1;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (uint16_t)(0x3c00u + ((s >> 20) & 0x1ff)) ^ ((s >> 3) & 0x8000u); }
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  uint8_t* pblk; cudaMalloc(&pblk, 4096); cudaMemset(pblk, 0, 4096);
  AttnArgs a{};
  a.pad_blk = pblk;
  a.qkv = qkv; a.keypad = kp; a.ctx = ctx; a.dtype = W2VS_BF16; a.B = B; a.T2 = T2; a.main_ctx = main_ctx; a.rc = rc;
  a.heads = heads; a.D = D;
  { int nb = 0; cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attn_tc_kernel, N_THREADS, SMEM_BYTES);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, attn_tc_kernel);
    printf("occupancy: %d CTAs/SM (regs %d, static smem %zu, dyn smem %d)\n", nb, fa.numRegs, fa.sharedSizeBytes, SMEM_BYTES); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_attention_tc(a, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) launch_attention_tc(a, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("attn_tc: %.1f us per launch (B=%d)  err=%s\n", ms * 100, B, cudaGetErrorString(cudaGetLastError()));
  { std::vector<uint16_t> o((size_t)B * M * D);
    cudaMemcpy(o.data(), ctx, o.size() * 2, cudaMemcpyDeviceToHost);
    double sum = 0, asum = 0; unsigned long long x = 0;
    for (size_t i = 0; i < o.size(); ++i) { uint32_t u = (uint32_t)o[i] << 16; float f; memcpy(&f, &u, 4); sum += f; asum += fabs(f); x = x * 1315423911ull + o[i]; }
    printf("ctx checksum: sum %.6f abs %.6f hash %016llx\n", sum, asum, x); }
#ifdef W2VS_ATTN_TRACE
  static long long tr[2][64][16];
  cudaMemcpyFromSymbol(tr, g_attn_trace, sizeof(tr));
  long long t0 = tr[0][0][0];
  printf("softmax warp0: tile | bar_in Sloaded sfull max+xch rescaled exp_done pfull pvdone_seen after_any before_sfull_wait ev10 epi_in pvdone stored item_of (cycles since first event)\n");
  for (int it = 0; it < 12; ++it) {
    printf("%2d |", it);
    for (int e = 0; e < 15; ++e) printf(" %7lld", tr[0][it][e] ? tr[0][it][e] - t0 : -1);
    printf("\n");
  }
  printf("MMA thread: tile | loop_top sfree_done S_issued vfull pfull_done PV_issued [S_done PV_done]\n");
  for (int it = 0; it < 12; ++it) {
    printf("%2d |", it);
    for (int e = 0; e < 8; ++e) printf(" %7lld", tr[1][it][e] ? tr[1][it][e] - t0 : -1);
    printf("\n");
  }
#endif
  return 0;
}
