// Incremental (chunk-by-chunk) forward -- placeholder until the cached-left-context path lands.
#include "common.cuh"
#include "kernels.h"
#include "layout.h"
using namespace w2vs;
extern "C" {
w2vs_status_t w2vs_stream_state_size(const w2vs_config*, int32_t, int32_t, int32_t, int32_t, int32_t, size_t*,
                                     size_t*, size_t*) {
  set_error("unsupported: incremental mode not built yet");
  return W2VS_UNSUPPORTED;
}
w2vs_status_t w2vs_stream_init(const w2vs_config*, int32_t, int32_t, int32_t, int32_t, int32_t, void*, size_t,
                               void*, size_t, void*) {
  set_error("unsupported: incremental mode not built yet");
  return W2VS_UNSUPPORTED;
}
w2vs_status_t w2vs_stream_step(const w2vs_config*, const void*, void*, void*, const void*, int32_t, int32_t,
                               int32_t, void*, int32_t, int32_t*, void*, size_t, void*) {
  set_error("unsupported: incremental mode not built yet");
  return W2VS_UNSUPPORTED;
}
}
