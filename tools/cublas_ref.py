"""Library reference point (not a product path): cuBLAS bf16 GEMM (torch.matmul) on the transformer-layer shapes of
cfg3, timed the same way as bench.py's per-kernel figures, next to this repo's tcgen05 kernel through the C ABI.
    python tools/cublas_ref.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from wav2vec_s_b200 import ops
M = 95744
shapes = [("qkv", 1024, 3072), ("out_proj", 1024, 1024), ("fc1", 1024, 4096), ("fc2", 4096, 1024)]
dev = "cuda"
torch.manual_seed(0)
def timeit(fn, reps=48):
    for _ in range(8): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, K, N in shapes:
    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t_lib = timeit(lambda: torch.matmul(A, W.t(), out=out))
    t_ours = timeit(lambda: ops.gemm(A, W, bias, None, out_dtype=torch.bfloat16))
    fl = 2.0 * M * N * K / 1e12
    print(f"{name:9s} M={M} N={N} K={K}: cuBLAS bf16 (no epilogue) {t_lib*1e3:7.1f} us = {fl/t_lib*1e3:6.0f} TFLOP/s | "
          f"gemm_tc2 (+bias, bf16 out) {t_ours*1e3:7.1f} us = {fl/t_ours*1e3:6.0f} TFLOP/s")
