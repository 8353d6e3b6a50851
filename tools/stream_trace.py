"""Phase timeline of the persistent incremental-step kernel (k_stream_fused.cu) on the GPU box.
    python tools/stream_trace.py [seconds_of_context]
Feeds a large-model stream up to the given left context, runs one more decision step and prints, for the first
and the last CTA of the grid, the mean duration of every phase and of every grid barrier over the layers."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wav2vec_s_b200 as W  # noqa: E402
from wav2vec_s_b200 import cabi  # noqa: E402

LARGE = dict(extractor_mode="layer_norm", encoder_layers=24, encoder_embed_dim=1024, encoder_ffn_embed_dim=4096,
             encoder_attention_heads=16, layer_norm_first=True, conv_bias=True, pos_type="sin", main_context=16,
             right_context=8)
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
torch.manual_seed(0)
m = W.BlockWiseWav2Vec2Model(LARGE).to("cuda", torch.bfloat16).eval()
L = int(seconds * 16000)
wav = torch.randn(1, L + 5120 * 4).cuda()
st = m.open_stream(B=1, max_seconds=seconds + 5, max_new_samples=7760 + 400, step_impl=2)
pos = 0
while pos < L:
    n = 7760 if pos == 0 else 5120
    st.step(wav[:, pos:pos + n])
    pos += n
lat = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    st.step(wav[:, pos:pos + 5120])
    e1.record()
    torch.cuda.synchronize()
    lat.append(e0.elapsed_time(e1))
    pos += 5120
buf = (C.c_uint64 * (2 * 64 * 12))()
cabi.check(cabi.lib().w2vs_debug_fused_trace(buf, len(buf)), "trace")
raw = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
tc = raw[:2 * 64 * 12].reshape(2, 64, 12)[:, :24, 11]
t = raw[:2 * 64 * 12].reshape(2, 64, 12)[:, :24, :11]
print("SM clock during the kernel (MHz):", [round(float((tc[c, 23] - tc[c, 0]) / ((t[c, 23, 0] - t[c, 0, 0]) / 1e3)), 1) for c in range(2)])
names = ["LN+QKV", "barrier", "attention", "barrier", "out_proj", "barrier", "LN+fc1", "barrier", "fc2", "barrier"]
print(f"step latency (ms): {lat}; left context {pos / 16000:.1f} s")
for c, who in enumerate(("first CTA", "last CTA")):
    d = np.diff(t[c], axis=1) / 1e3
    print(who, " ".join(f"{n}={v:.2f}" for n, v in zip(names, d.mean(0))), f"| layer {d.sum(1).mean():.2f} us, "
          f"all layers {(t[c, -1, 10] - t[c, 0, 0]) / 1e3:.1f} us")
for c, who in enumerate(("first CTA", "last CTA")):
    t0 = t[c, 12, 0]
    print(who, "layer 12 phase marks (us from layer start):", " ".join(f"{i}:{(int(t[c, 12, i]) - int(t0)) / 1e3:.2f}" for i in range(11)))
flags = C.c_int32(0)
cabi.check(cabi.lib().w2vs_debug_fault_flags(C.byref(flags)), "faults")
print("fault flags:", flags.value)
