"""Thin Python wrappers over the single-operator entry points of the C ABI (w2vs_op_*).

These are the unit-test surface of the kernels: the same launchers the forward uses, callable on
torch CUDA tensors.  No fallback: non-CUDA tensors raise.
"""
import ctypes as C

import torch

from . import cabi


def _dt(t):
    if t.dtype == torch.float32:
        return cabi.F32
    if t.dtype == torch.bfloat16:
        return cabi.BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _stream(t):
    if t.device.type != "cuda":
        raise RuntimeError("w2vs ops need CUDA tensors (no CPU fallback)")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def gemm(A, W, bias=None, residual=None, out_dtype=None, gelu=False, impl=cabi.GEMM_AUTO, M=None, K=None, lda=None,
         splitk=False):
    """C[M,N] = A[M,K] . W[N,K]^T + bias (+GELU) (+residual).  ``lda < K`` gives the overlapping-row
    (strided conv) view of a flat A buffer."""
    N = W.size(0)
    K = W.size(1) if K is None else K
    M = A.size(0) if M is None else M
    lda = A.stride(0) if lda is None else lda
    out_dtype = out_dtype or A.dtype
    if residual is not None:
        # the library adds the residual in place (C aliases the residual, as the encoder's residual stream does)
        assert out_dtype == torch.float32 and tuple(residual.shape) == (M, N)
        Cm = residual.to(torch.float32).clone()
        residual = Cm
    else:
        Cm = torch.empty((M, N), dtype=out_dtype, device=A.device)
    flags = (cabi.EPI_GELU if gelu else 0) | (cabi.EPI_SPLITK if splitk else 0)
    cabi.check(cabi.lib().w2vs_op_gemm(impl, _dt(A), _dt(Cm), _ptr(A), lda, _ptr(W), _ptr(bias), _ptr(residual),
                                       _ptr(Cm), N, M, N, K, flags, _stream(A)), "w2vs_op_gemm")
    return Cm


def layernorm(x, gamma, beta, act_dtype=torch.bfloat16, gelu=False, want_f32=True):
    rows, N = x.shape
    o32 = torch.empty((rows, N), dtype=torch.float32, device=x.device) if want_f32 else None
    oa = torch.empty((rows, N), dtype=act_dtype, device=x.device)
    cabi.check(cabi.lib().w2vs_op_layernorm(_dt(x), _ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), _ptr(o32),
                                            _dt(oa), _ptr(oa), N, rows, N, int(gelu), _stream(x)),
               "w2vs_op_layernorm")
    return o32, oa


def attention(qkv, keypad, T_pad, main_ctx, right_ctx, heads, impl=0):
    """qkv [B, M, 3D], keypad bool/uint8 [B, M] -> ctx [B, M, D]."""
    B, M, D3 = qkv.shape
    D = D3 // 3
    ctx = torch.empty((B, M, D), dtype=qkv.dtype, device=qkv.device)
    kp = keypad.to(torch.uint8).contiguous()
    cabi.check(cabi.lib().w2vs_op_attention(impl, _dt(qkv), _ptr(qkv), _ptr(kp), _ptr(ctx), B, T_pad, main_ctx,
                                            right_ctx, heads, D, _stream(qkv)), "w2vs_op_attention")
    return ctx
