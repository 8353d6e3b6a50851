"""Phase timeline of the cluster incremental-step kernel (k_stream_cluster.cu) on the GPU box.
    python tools/cluster_trace.py [seconds_of_context]
Feeds a large-model stream up to the given left context, runs a few more decision steps and prints, for CTA 0, the
mean duration of every stage of the two phases of a layer (globaltimer stamps, w2vs_debug_cluster_trace)."""
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
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
m = W.BlockWiseWav2Vec2Model(LARGE).to("cuda", torch.bfloat16).eval()
L = int(seconds * 16000)
wav = torch.randn(1, L + 5120 * 16).cuda()
st = m.open_stream(B=1, max_seconds=seconds + 5, max_new_samples=7760 + 400, step_impl=impl)
pos = 0
while pos < L:
    n = 7760 if pos == 0 else 5120
    st.step(wav[:, pos:pos + n])
    pos += n
lat0 = []
for _ in range(6):                      # untraced steps first: the stamps cost the kernel a few percent
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    st.step(wav[:, pos:pos + 5120])
    e1.record()
    torch.cuda.synchronize()
    lat0.append(round(e0.elapsed_time(e1), 4))
    pos += 5120
print(f"step latency without stamps (ms): {lat0}")
cabi.check(cabi.lib().w2vs_debug_cluster_trace(None, 1), "trace on")
lat = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    st.step(wav[:, pos:pos + 5120])
    e1.record()
    torch.cuda.synchronize()
    lat.append(round(e0.elapsed_time(e1), 4))
    pos += 5120
print(f"step latency (ms): {lat}; left context {pos / 16000:.1f} s")
if impl == 3:
    buf = (C.c_uint64 * (64 * 32))()
    cabi.check(cabi.lib().w2vs_debug_cluster_trace(buf, len(buf)), "trace")
    t = np.frombuffer(buf, dtype=np.uint64).astype(np.int64).reshape(64, 32)[:24]
    ev = [0, 18, 19, 1, 15, 2, 3, 4, 5, 6, 7, 13, 8, 20, 21, 9, 16, 10, 11, 17, 12, 14]
    names = ["x slice + local stats", "stats exchange", "normalise", "wait weights", "QKV+scatter", "reduce+gather", "attention", "merge-send", "merge+ctx gather",
             "out_proj", "grid barrier", "(next phase start)", "x slice + local stats", "stats exchange", "normalise", "wait weights", "fc1+scatter", "gelu+gather",
             "wait weights", "fc2", "grid barrier"]
    seq = t[:, ev]
    d = np.diff(seq, axis=1) / 1e3
    print("CTA 0, mean over layers (us):")
    for n, v in zip(names, d.mean(0)):
        print(f"  {n:20s} {v:6.2f}")
    print(f"  phase C grid barrier: bar.sync {(t[:, 22] - t[:, 12]).mean() / 1e3:.2f}, release + arrive {(t[:, 23] - t[:, 22]).mean() / 1e3:.2f}, "
          f"poll + bar.sync {(t[:, 14] - t[:, 23]).mean() / 1e3:.2f} us")
    print(f"  QKV stage: product {(t[:, 24] - t[:, 15]).mean() / 1e3:.2f}, sends {(t[:, 25] - t[:, 24]).mean() / 1e3:.2f}, "
          f"early loads + sync + wait for the peers {(t[:, 2] - t[:, 25]).mean() / 1e3:.2f} us")
    print(f"  layer total {(seq[:, -1] - seq[:, 0]).mean() / 1e3:.2f} us; all layers {(t[23, 14] - t[0, 0]) / 1e3:.1f} us")
flags = C.c_int32(0)
cabi.check(cabi.lib().w2vs_debug_fault_flags(C.byref(flags)), "faults")
print("fault flags:", flags.value)
