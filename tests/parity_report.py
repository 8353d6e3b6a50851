"""Per-stage max-abs-rel error of the CUDA path vs the reference golden vectors / the oracle (GPU box).
    python tests/parity_report.py > gpurun_out/parity_report.txt"""
import os, sys, warnings
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import wav2vec_s_b200 as W
from oracle import cases
from oracle import w2vs_oracle as O
from helpers import load_golden, case_inputs, valid_rel_err

print(f"{'case':32s} {'dtype':5s} {'conv_out':>9s} {'post_proj':>9s} {'layer0':>9s} {'mid':>9s} {'output':>9s}")
for name, c in cases.CASES.items():
    if c.get("api", "fairseq") != "fairseq":
        continue
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    otaps = {}
    yo, fmo = O.extract_features(sd, cfg, wav, pm, taps=otaps)
    for dtype in (torch.float32, torch.bfloat16):
        m = W.Wav2VecSModel(cfg); m.load_state_dict(sd, strict=False); m = m.to("cuda", dtype).eval()
        taps = {}
        y, fm = m._encode(wav.cuda().to(dtype if os.environ.get("W2VS_SRC_BF16") else torch.float32), padding_mask=None if pm is None else pm.cuda(), taps=taps)
        fmask = g["fmask"]
        T2 = otaps["enc_in"].size(0)
        fm_t2 = None
        if fmo is not None:
            fm_t2 = torch.nn.functional.pad(fmo, (0, T2 - fmo.size(1)), value=True).numpy()
        elif T2 != yo.size(1):
            fm_t2 = np.zeros((yo.size(0), T2), dtype=bool); fm_t2[:, yo.size(1):] = True
        L = cfg["encoder_layers"]
        e_conv = valid_rel_err(taps["conv_out"].cpu(), torch.from_numpy(g["conv_out"]).transpose(1, 2))
        e_proj = valid_rel_err(taps["post_proj"].cpu(), g["post_proj"], fmask) if "post_proj" in g else float("nan")
        e_l0 = valid_rel_err(taps["layers"][0, :, :T2].cpu().transpose(0, 1), otaps["layer0"], fm_t2, time_first=True)
        e_mid = valid_rel_err(taps["layers"][L // 2, :, :T2].cpu().transpose(0, 1), otaps[f"layer{L//2}"], fm_t2, time_first=True)
        e_out = valid_rel_err(y.cpu(), g["y"], fmask)
        print(f"{name:32s} {'fp32' if dtype == torch.float32 else 'bf16':5s} {e_conv:9.2e} {e_proj:9.2e} {e_l0:9.2e} {e_mid:9.2e} {e_out:9.2e}")
