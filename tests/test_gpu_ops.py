"""GPU unit tests of the kernels through the C ABI (w2vs_op_*), each against a plain PyTorch fp32
reference of the same op on the same seeded inputs."""
import numpy as np
import pytest
import torch

from wav2vec_s_b200 import cabi, ops
from oracle import w2vs_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(y, ref):
    return float((y.float() - ref.float()).abs().max() / ref.float().abs().max())


def _gemm_ref(A, W, bias, res, gelu):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if gelu:
        y = torch.nn.functional.gelu(y)
    if res is not None:
        y = y + res
    return y


@pytest.mark.parametrize("impl", [cabi.GEMM_TCGEN05_2CTA, cabi.GEMM_SIMT])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 128, 192), (1000, 512, 1536), (257, 3072, 1024),
                                   (4096, 1024, 4096), (24, 384, 128), (130, 64, 128), (5000, 768, 512),
                                   # a few hundred rows (a batch of incremental streams): the in-place fp32 products
                                   # run split over K with L2 reductions (uneven last range: 17 K blocks in two ranges)
                                   (384, 1024, 4096), (384, 1024, 1024), (96, 1024, 4096), (200, 512, 2048),
                                   (384, 1024, 1088), (100, 256, 4672)])
def test_gemm_bf16(impl, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(DEV, torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    for gelu, use_res, odt in [(False, False, torch.bfloat16), (True, False, torch.bfloat16), (False, True, torch.float32)]:
        y = ops.gemm(A, W, bias, res if use_res else None, out_dtype=odt, gelu=gelu, impl=impl)
        ref = _gemm_ref(A, W, bias, res if use_res else None, gelu)
        tol = 1e-2 if odt == torch.bfloat16 else 2e-5 * max(1, K // 512) + 1e-5
        assert rel(y, ref) < tol, (impl, M, N, K, gelu, use_res)
        if use_res:
            # fixed reduction order unless the caller allows split-K (the incremental steps of a batch of streams do)
            assert torch.equal(y, ops.gemm(A, W, bias, res, out_dtype=odt, impl=impl))
            ys = ops.gemm(A, W, bias, res, out_dtype=odt, impl=impl, splitk=True)
            assert rel(ys, ref) < tol, (impl, M, N, K, "split-K")


@pytest.mark.parametrize("M", [1, 7, 16, 24, 33, 48, 64])
@pytest.mark.parametrize("N,K", [(1024, 1024), (3072, 1024), (4096, 1024), (1024, 4096), (1024, 512), (512, 256)])
def test_gemm_skinny_bf16(M, N, K):
    """Weight-streaming kernel of the incremental steps (M <= 64) against the fp32 restatement and against
    what AUTO picks for these shapes (it must be this kernel)."""
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(DEV, torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    for gelu, use_res, odt in [(False, False, torch.bfloat16), (True, False, torch.bfloat16), (False, True, torch.float32),
                               (False, False, torch.float32)]:
        y = ops.gemm(A, W, bias, res if use_res else None, out_dtype=odt, gelu=gelu, impl=cabi.GEMM_SKINNY)
        ref = _gemm_ref(A, W, bias, res if use_res else None, gelu)
        tol = 1e-2 if odt == torch.bfloat16 else 2e-5 * max(1, K // 512) + 1e-5
        assert rel(y, ref) < tol, (M, N, K, gelu, use_res)
        y2 = ops.gemm(A, W, bias, res if use_res else None, out_dtype=odt, gelu=gelu, impl=cabi.GEMM_AUTO)
        assert torch.equal(y, y2)


@pytest.mark.parametrize("impl", [cabi.GEMM_TCGEN05_2CTA, cabi.GEMM_SIMT])
@pytest.mark.parametrize("rows,C,k,s", [(1000, 512, 3, 2), (777, 512, 2, 2), (300, 64, 3, 2), (129, 64, 2, 2)])
def test_gemm_strided_conv_view(impl, rows, C, k, s):
    """Conv1d(C->C, k, stride s) on channels-last rows == GEMM with lda = s*C < K = k*C."""
    g = torch.Generator(device="cpu").manual_seed(rows + C)
    x = (torch.randn(rows * s + 64, C, generator=g) * 0.5).to(DEV, torch.bfloat16)   # + slack rows
    w = (torch.randn(C, C, k, generator=g) * 0.05)
    wk = w.permute(0, 2, 1).reshape(C, k * C).contiguous().to(DEV, torch.bfloat16)        # [C_out][j*C_in+ci]
    y = ops.gemm(x, wk, None, None, out_dtype=torch.float32, impl=impl, M=rows, K=k * C, lda=s * C)
    xin = x[: rows * s + (k - s)].float().t().unsqueeze(0)                                 # [1, C, t]
    ref = torch.nn.functional.conv1d(xin, w.to(DEV, torch.bfloat16).float(), stride=s)[0].t()[:rows]
    assert rel(y, ref) < 5e-5


def test_gemm_fp32_simt():
    g = torch.Generator(device="cpu").manual_seed(3)
    A = torch.randn(333, 768, generator=g).to(DEV)
    W = (torch.randn(1024, 768, generator=g) * 0.05).to(DEV)
    b = torch.randn(1024, generator=g).to(DEV)
    y = ops.gemm(A, W, b, None, out_dtype=torch.float32, gelu=True)
    ref = torch.nn.functional.gelu((A.double() @ W.double().t() + b.double())).float()
    assert rel(y, ref) < 2e-6


@pytest.mark.parametrize("N", [64, 128, 512, 768, 1024])
def test_layernorm(N):
    g = torch.Generator(device="cpu").manual_seed(N)
    x = (torch.randn(1001, N, generator=g) * 3 + 1).to(DEV)
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(N, generator=g)).to(DEV)
    for gelu in (False, True):
        o32, oa = ops.layernorm(x, gamma, beta, torch.bfloat16, gelu)
        ref = torch.nn.functional.layer_norm(x, (N,), gamma, beta, 1e-5)
        if gelu:
            ref = torch.nn.functional.gelu(ref)
        assert rel(o32, ref) < 2e-6
        assert rel(oa, ref) < 5e-3
    o32, _ = ops.layernorm(x.bfloat16(), gamma, beta, torch.float32, False)
    assert rel(o32, torch.nn.functional.layer_norm(x.bfloat16().float(), (N,), gamma, beta, 1e-5)) < 2e-6


def _attention_ref(qkv, keypad, T2, main, rc, heads):
    """Dense-mask restatement: additive -1e4 block mask + -inf key padding (reference semantics)."""
    B, M, D3 = qkv.shape
    D = D3 // 3
    hd = D // heads
    q, k, v = qkv.float().split(D, dim=-1)
    _, _, masked = O.block_mask_structure(T2, main, rc)
    am = torch.zeros(M, M, device=qkv.device).masked_fill(masked.to(qkv.device), -1e4)
    q = q.view(B, M, heads, hd).transpose(1, 2) * hd ** -0.5
    k = k.view(B, M, heads, hd).transpose(1, 2)
    v = v.view(B, M, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2) + am
    s = s.masked_fill(keypad.bool()[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, M, D)


@pytest.mark.parametrize("T2,main,rc", [(500, 16, 8), (18, 16, 8), (12, 16, 8), (38, 16, 8), (250, 8, 4),
                                        (200, 32, 16), (100, 16, 0), (2, 16, 8), (1000, 16, 8),
                                        # context_type="sampling" block sizes (wav2vec_S.py:392-395): not powers of two
                                        (300, 20, 10), (310, 12, 6), (500, 30, 14), (260, 10, 4), (400, 24, 12),
                                        # M = 6748 tokens: padded keys beyond the 32nd 128-token block (90 s utterances)
                                        (4500, 16, 8)])
@pytest.mark.parametrize("dtype,impl", [(torch.float32, 1), (torch.bfloat16, 1), (torch.bfloat16, 2), (torch.bfloat16, 3)])
def test_attention(T2, main, rc, dtype, impl):
    B, heads, D = 2, 3, 192
    nb = T2 // main
    M = T2 + nb * rc
    g = torch.Generator(device="cpu").manual_seed(T2 + main)
    qkv = (torch.randn(B, M, 3 * D, generator=g) * 1.5).to(DEV, dtype)
    # key padding as the encoder produces it: ragged tail of utterance 1, copies inherit + out-of-range
    valid = [T2, max(1, int(T2 * 0.6))]
    pm = torch.zeros(B, T2, dtype=torch.bool)
    pm[1, valid[1]:] = True
    rc_idx, oor, _ = O.block_mask_structure(T2, main, rc)
    kp = torch.cat([pm, pm[:, rc_idx] | oor[None]], 1) if rc > 0 else pm
    ctx = ops.attention(qkv, kp.to(DEV), T2, main, rc, heads, impl=impl)
    ref = _attention_ref(qkv, kp.to(DEV), T2, main, rc, heads)
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    # rows of padded queries are don't-care in the encoder but are still well defined here
    assert rel(ctx, ref) < tol
