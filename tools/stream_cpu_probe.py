import sys, time, statistics, torch
sys.path.insert(0, '/root/repo')
import bench, wav2vec_s_b200 as W
from wav2vec_s_b200.model import EncoderStream
cfg = bench.model_cfg("large"); B=int(sys.argv[1]) if len(sys.argv)>1 else 1
dev = torch.device("cuda",0); torch.manual_seed(0)
model = W.BlockWiseWav2Vec2Model(cfg).to(dev, torch.bfloat16).eval()
L = 30*16000; wav = torch.randn(B, L, device=dev)
bounds=[7760]
while bounds[-1]+5120 < L: bounds.append(bounds[-1]+5120)
bounds.append(L)
for rep in range(2):
    st = model.open_stream(B=B, max_seconds=31, max_new_samples=7760+400)
    cpu=[]; pos=0
    torch.cuda.synchronize()
    for n in bounds:
        t0=time.perf_counter()
        y = st.step(wav[:, pos:n], EncoderStream.FINAL if n>=L else EncoderStream.NONE)
        t1=time.perf_counter()
        torch.cuda.synchronize()
        t2=time.perf_counter()
        cpu.append(((t1-t0)*1e3,(t2-t0)*1e3)); pos=n
print("B",B,"cpu launch ms p50", statistics.median(c[0] for c in cpu[1:-1]), "total ms p50", statistics.median(c[1] for c in cpu[1:-1]))
