// Shared device/host helpers for the wav2vec-S sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/w2vs.h"

typedef __nv_bfloat16 bf16;

namespace w2vs {

// ---- per-thread diagnostics ------------------------------------------------------------------
extern thread_local char g_last_error[512];
extern thread_local int64_t g_launch_count;
void set_error(const char* fmt, ...);
// Optional per-launch device timing (w2vs_prof_*): when enabled on this host thread, every launch is
// followed by an event record on its stream; consecutive events bracket one kernel.
extern thread_local bool g_prof_on;
void prof_mark(const char* what, cudaStream_t st);

#define W2VS_CHECK_LAUNCH(what)                                                        \
  do {                                                                                 \
    ++::w2vs::g_launch_count;                                                          \
    if (::w2vs::g_prof_on) ::w2vs::prof_mark(what, st);                                \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ::w2vs::set_error("%s: %s (%s:%d)", what, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return W2VS_CUDA_ERROR;                                                          \
    }                                                                                  \
  } while (0)

#define W2VS_TRY(expr)                      \
  do {                                      \
    w2vs_status_t s__ = (expr);             \
    if (s__ != W2VS_OK) return s__;         \
  } while (0)

#define W2VS_REQUIRE(cond, msg)                                         \
  do {                                                                  \
    if (!(cond)) {                                                      \
      ::w2vs::set_error("invalid value: %s (%s)", msg, #cond);          \
      return W2VS_INVALID_VALUE;                                        \
    }                                                                   \
  } while (0)

// ---- programmatic dependent launch (incremental steps are a chain of ~200 tiny kernels) -------------------------
// Kernels that start with pdl_prologue() may be launched with launch_pdl(): the grid is set up while its
// predecessor still runs, and griddepcontrol.wait holds it until the predecessor has completed and flushed its
// writes -- ordering and results are those of a plain stream launch, the launch latency is hidden.  Both
// instructions are no-ops for a normal launch.
extern thread_local bool g_pdl_on;     // set by the incremental driver (stream.cu)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }
template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl_on ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- per-device host-side caches ---------------------------------------------------------------------------------
// cudaFuncSetAttribute and the SM count are properties of one device: everything the launchers remember is keyed
// by the device that is current on the calling thread, so one process can drive several GPUs (one model replica
// per device) through the same library.  Races between host threads are benign (idempotent values).
constexpr int kMaxDevices = 64;
static inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < kMaxDevices ? d : 0;
}
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  bool& here() { return done[current_device()]; }
};
static inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int d = current_device();
  if (!n[d]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d);
    n[d] = v > 0 ? v : 148;
  }
  return n[d];
}
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- dtype helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(int16_t v) { return (float)v * (1.0f / 32768.0f); }   // 16-bit PCM sample
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(p);
}

// 8 consecutive elements <-> 8 floats (16 B of bf16 or 32 B of fp32); pointers must be aligned.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  __half2 h[4] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3]), __floats2half2_rn(v[4], v[5]),
                  __floats2half2_rn(v[6], v[7])};
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(h);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- math --------------------------------------------------------------------------------------
// erf GELU, x * Phi(x) (torch nn.GELU / fairseq modules/gelu.py:24-25).  erf by Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7) with one MUFU.RCP + one MUFU.EX2 issued as the raw approx instructions: 14
// instructions, no branches.  (__frcp_rn / exp2f() expand to ~28 instructions with a slow-path call, which
// made the fc1 epilogue 2.5x slower than the MMAs it has to keep pace with.)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-p, e, 1.0f);          // erf(|x|/sqrt2)
  const float hx = 0.5f * x;
  return fmaf(copysignf(erf_abs, x), hx, hx);       // 0.5 x (1 + erf(x/sqrt2))
}

// ---- packed fp32 pairs (sm_100 f32x2 pipe): one instruction for two lanes of elementwise arithmetic -------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two erf-GELUs at once: same formula as gelu_erf, 20 instructions per pair instead of 28
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t x = pack2(x0, x1);
  const uint64_t z = pack2(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
  float d0, d1;
  unpack2(ffma2(pack2(0.3275911f, 0.3275911f), z, pack2(1.0f, 1.0f)), d0, d1);
  const uint64_t t = pack2(rcp_approx(d0), rcp_approx(d1));
  uint64_t p = ffma2(pack2(1.061405429f, 1.061405429f), t, pack2(-1.453152027f, -1.453152027f));
  p = ffma2(p, t, pack2(1.421413741f, 1.421413741f));
  p = ffma2(p, t, pack2(-0.284496736f, -0.284496736f));
  p = ffma2(p, t, pack2(0.254829592f, 0.254829592f));
  p = fmul2(p, t);
  float e0, e1;
  unpack2(fmul2(fmul2(z, z), pack2(-1.4426950408889634f, -1.4426950408889634f)), e0, e1);
  const uint64_t e = pack2(ex2_approx(e0), ex2_approx(e1));
  float q0, q1;
  unpack2(ffma2(p, e, pack2(-1.0f, -1.0f)), q0, q1);     // q = -erf(|x|/sqrt2)
  // copysign(erf_abs, x): magnitude bits of q, sign bit of x
  const float s0 = __uint_as_float((__float_as_uint(q0) & 0x7fffffffu) | (__float_as_uint(x0) & 0x80000000u));
  const float s1 = __uint_as_float((__float_as_uint(q1) & 0x7fffffffu) | (__float_as_uint(x1) & 0x80000000u));
  const uint64_t hx = fmul2(x, pack2(0.5f, 0.5f));
  unpack2(ffma2(pack2(s0, s1), hx, hx), x0, x1);
}

// erf-GELU for bf16 outputs: x * Phi(x) with Phi(x) = (1 + tanh(g(x))) / 2, g(x) = x (c0 + c1 x^2 + c2 x^4)
// fitted (minimax over [-8, 8]) to the exact erf form: |error| <= 2.5e-5 before the MUFU.TANH approximation
// (relative error 2^-11 on tanh, i.e. <= 2.4e-4 |x| on the result) -- one eighth of a bf16 ulp of the outputs
// it feeds.  One MUFU op and 7 FMA-pipe cycles per element instead of two MUFU ops and 14 cycles: the erf form
// made the fc1 epilogue, conv0 and the conv LayerNorm+GELU pass MUFU/FMA-pipe bound.  c2 < 0, so x^2 is clamped at
// 64 (g stays monotonic; tanh is saturated there anyway).  The fp32 mode (1e-4 bar) keeps gelu_erf.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_tanh2(float& x0, float& x1) {
  const uint64_t x = pack2(x0, x1);
  float q0, q1;
  unpack2(fmul2(x, x), q0, q1);
  const uint64_t x2 = pack2(fminf(q0, 64.f), fminf(q1, 64.f));
  uint64_t p = ffma2(pack2(-0.00035151765342717335f, -0.00035151765342717335f), x2,
                     pack2(0.037005651782227346f, 0.037005651782227346f));
  p = ffma2(p, x2, pack2(0.7975078774034253f, 0.7975078774034253f));
  float g0, g1;
  unpack2(fmul2(p, x), g0, g1);
  const uint64_t t = pack2(tanh_approx(g0), tanh_approx(g1));
  const uint64_t hx = fmul2(x, pack2(0.5f, 0.5f));
  unpack2(ffma2(t, hx, hx), x0, x1);
}
// packed form: two values in, two values out (no unpacking around the polynomial)
__device__ __forceinline__ uint64_t gelu_tanh2p(uint64_t x) {
  float q0, q1;
  unpack2(fmul2(x, x), q0, q1);
  const uint64_t x2 = pack2(fminf(q0, 64.f), fminf(q1, 64.f));
  uint64_t p = ffma2(pack2(-0.00035151765342717335f, -0.00035151765342717335f), x2,
                     pack2(0.037005651782227346f, 0.037005651782227346f));
  p = ffma2(p, x2, pack2(0.7975078774034253f, 0.7975078774034253f));
  float g0, g1;
  unpack2(fmul2(p, x), g0, g1);
  const uint64_t t = pack2(tanh_approx(g0), tanh_approx(g1));
  const uint64_t hx = fmul2(x, pack2(0.5f, 0.5f));
  return ffma2(t, hx, hx);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = fminf(x * x, 64.f);
  float p = fmaf(-0.00035151765342717335f, x2, 0.037005651782227346f);
  p = fmaf(p, x2, 0.7975078774034253f);
  const float hx = 0.5f * x;
  return fmaf(tanh_approx(p * x), hx, hx);
}
// GELU by output type: bf16 results take the tanh form, fp32 results the erf form
template <typename TOut> __device__ __forceinline__ void gelu2(float& a, float& b) {
  if (sizeof(TOut) == 2) gelu_tanh2(a, b); else gelu_erf2(a, b);
}
template <typename TOut> __device__ __forceinline__ float gelu1(float a) {
  return sizeof(TOut) == 2 ? gelu_tanh(a) : gelu_erf(a);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace w2vs
