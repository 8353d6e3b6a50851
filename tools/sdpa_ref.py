"""Library reference point for the attention kernel (not a product path): what the reference's
F.multi_head_attention_forward reaches on the GPU for the cfg3 shape -- torch SDPA with the dense additive block mask
(the reference materialises it, wav2vec_S.py:444-489) -- next to attn_tc_kernel through the C ABI.
    python tools/sdpa_ref.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from wav2vec_s_b200 import ops
B, H, T2, main, rc, hd = 64, 16, 1000, 16, 8, 64
nb = T2 // main; M = T2 + nb * rc; D = H * hd
dev = "cuda"; torch.manual_seed(0)
qkv = (torch.randn(B, M, 3 * D, device=dev) * 0.5).to(torch.bfloat16)
# dense mask exactly as gen_block_attn_mask builds it
qblk = torch.cat([torch.arange(T2) // main, torch.arange(nb).repeat_interleave(rc)])
kblk_main = torch.arange(T2) // main
masked = torch.ones(M, M, dtype=torch.bool)
masked[:, :T2] = qblk[:, None] < kblk_main[None, :]
masked[:, T2:] = qblk[:, None] != torch.arange(nb).repeat_interleave(rc)[None, :]
mask = torch.zeros(M, M).masked_fill(masked, -1e4).to(dev, torch.bfloat16)
q, k, v = [t.reshape(B, M, H, hd).transpose(1, 2) for t in qkv.split(D, dim=-1)]
def timeit(fn, reps=24):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_sdpa = timeit(lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=mask))
kp = torch.zeros(B, M, dtype=torch.uint8, device=dev)
t_ours = timeit(lambda: ops.attention(qkv, kp, T2, main, rc, H))
vis = int((~masked).sum())
print(f"attention B={B} H={H} M={M}: torch SDPA with the dense mask {t_sdpa*1e3:8.1f} us | attn_tc_kernel {t_ours*1e3:8.1f} us "
      f"({4.0*D*vis*B/t_ours/1e9:.0f} algorithmic TFLOP/s on the {100*vis/(M*M):.1f} % visible pairs)")
try:
    from flash_attn import flash_attn_func
    qf, kf, vf = [t.reshape(B, M, H, hd) for t in qkv.split(D, dim=-1)]
    t_fa = timeit(lambda: flash_attn_func(qf, kf, vf, causal=True))
    print(f"flash_attn 2.8 causal (a different, denser mask: 50 % of the pairs; library kernel) {t_fa*1e3:8.1f} us")
except Exception as e:
    print("flash_attn not usable here:", type(e).__name__, e)
