"""Per-kernel device time of one steady-state incremental step (development aid).
    python tools/stream_kernel_probe.py [B]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, wav2vec_s_b200 as W
from wav2vec_s_b200 import cabi
from wav2vec_s_b200.model import EncoderStream
cfg = bench.model_cfg("large"); B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0); torch.manual_seed(0)
model = W.BlockWiseWav2Vec2Model(cfg).to(dev, torch.bfloat16).eval()
L = 30 * 16000; wav = torch.randn(B, L, device=dev)
st = model.open_stream(B=B, max_seconds=31, max_new_samples=7760 + 400)
pos = 7760
st.step(wav[:, :pos], EncoderStream.NONE)
for _ in range(40):
    st.step(wav[:, pos:pos + 5120], EncoderStream.NONE); pos += 5120
state = {"pos": pos}
def one():
    st.step(wav[:, state["pos"]:state["pos"] + 5120], EncoderStream.NONE); state["pos"] += 5120
prof = cabi.profile_step(one, reps=4, detail=True)
tot = sum(v["ms"] for k, v in prof.items() if "[" in k or "_kernel" in k)
print(f"B={B}: {tot:.3f} ms of kernels per step")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    if "[" in k or "_kernel" in k:
        print(f"  {v['ms']:.4f} ms  x{v['count']:<4d} {k}")
