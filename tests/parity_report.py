"""Max-abs-rel error of the CUDA path vs the reference golden vectors / the oracle on every parity case (GPU box).
    python tests/parity_report.py > gpurun_out/parity_report.txt
Columns: per-stage errors against the fp32 reference (golden / oracle taps); `output` = encoder output against the
reference's fp32 output; `same-w` (bf16 rows) = against the fp32-arithmetic oracle on the identical, bf16-valued
weights -- the bf16-mode contract (tests/helpers.py)."""
import os, sys, warnings
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import wav2vec_s_b200 as W
from oracle import cases, synth
from oracle import w2vs_oracle as O
from helpers import load_golden, case_inputs, valid_rel_err, bf16_valued

print(f"{'case':32s} {'dtype':5s} {'conv_out':>9s} {'post_proj':>9s} {'layer0':>9s} {'mid':>9s} {'output':>9s} {'same-w':>9s}")
for name, c in cases.CASES.items():
    if c.get("api", "fairseq") != "fairseq":
        continue
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    otaps = {}
    yo, fmo = O.extract_features(sd, cfg, wav, pm, taps=otaps)
    yq, _ = O.extract_features(bf16_valued(sd), cfg, wav, pm)
    k = c.get("compact")
    for dtype in (torch.float32, torch.bfloat16):
        m = W.Wav2VecSModel(cfg); m.load_state_dict(sd, strict=False); m = m.to("cuda", dtype).eval()
        taps = {}
        y, fm = m._encode(wav.cuda(), padding_mask=None if pm is None else pm.cuda(), taps=taps)
        fmask = g["fmask"]
        T2 = otaps["enc_in"].size(0)
        fm_t2 = None
        if fmo is not None:
            fm_t2 = torch.nn.functional.pad(fmo, (0, T2 - fmo.size(1)), value=True).numpy()
        elif T2 != yo.size(1):
            fm_t2 = np.zeros((yo.size(0), T2), dtype=bool); fm_t2[:, yo.size(1):] = True
        L = cfg["encoder_layers"]
        n_conv = len(O.conv_layers_of(cfg))
        e_conv = valid_rel_err(taps["conv_out"].cpu(), otaps[f"conv{n_conv - 1}"].transpose(1, 2))
        e_proj = valid_rel_err(taps["post_proj"].cpu(), otaps["post_proj"], None if fmo is None else fmo.numpy())
        e_l0 = valid_rel_err(taps["layers"][0, :, :T2].cpu().transpose(0, 1), otaps["layer0"], fm_t2, time_first=True)
        e_mid = valid_rel_err(taps["layers"][L // 2, :, :T2].cpu().transpose(0, 1), otaps[f"layer{L//2}"], fm_t2, time_first=True)
        e_out = valid_rel_err(y[:, ::k].cpu(), g["y"], fmask) if k else valid_rel_err(y.cpu(), g["y"], fmask)
        e_same = valid_rel_err(y.cpu(), yq, None if fmo is None else fmo.numpy()) if dtype == torch.bfloat16 else float("nan")
        print(f"{name:32s} {'fp32' if dtype == torch.float32 else 'bf16':5s} {e_conv:9.2e} {e_proj:9.2e} {e_l0:9.2e} {e_mid:9.2e} {e_out:9.2e} {e_same:9.2e}")
        del m, taps, y
        torch.cuda.empty_cache()

# shapes beyond the golden set: large 30 s, base 2 x 15 s ragged, 90 s ragged (M = 6748 tokens)
def extra(label, cfg, sd, wav, pm):
    yo, fmo = O.extract_features(sd, cfg, wav, pm)
    yq, _ = O.extract_features(bf16_valued(sd), cfg, wav, pm)
    fmn = None if fmo is None else fmo.numpy()
    for dtype in (torch.float32, torch.bfloat16):
        m = W.Wav2VecSModel(cfg); m.load_state_dict(sd, strict=False); m = m.to("cuda", dtype).eval()
        y, _ = m.extract_features(wav.cuda(), None if pm is None else pm.cuda())
        e_same = valid_rel_err(y.cpu(), yq, fmn) if dtype == torch.bfloat16 else float("nan")
        print(f"{label:32s} {'fp32' if dtype == torch.float32 else 'bf16':5s} {'':>9s} {'':>9s} {'':>9s} {'':>9s} {valid_rel_err(y.cpu(), yo, fmn):9.2e} {e_same:9.2e}")
        del m, y
        torch.cuda.empty_cache()

SR = 16000
cfg = O.large_cfg(); sd = synth.make_state_dict(cfg, cases.WSEED)
extra("large_30s", cfg, sd, synth.make_waveform(1, 30 * SR, cases.XSEED + 30), None)
cfg = O.base_cfg(); sd = synth.make_state_dict(cfg, cases.WSEED)
wav = synth.make_waveform(2, 15 * SR, cases.XSEED + 15); pm = O.lengths_to_padding_mask(torch.tensor([15 * SR, 9 * SR + 1234]))
extra("base_2x15s_ragged", cfg, sd, wav.masked_fill(pm, 0.0), pm)
cfg = cases.tiny(encoder_layers=2, encoder_embed_dim=256, encoder_ffn_embed_dim=512, encoder_attention_heads=4, layer_norm_first=True, conv_bias=True)
sd = synth.make_state_dict(cfg, cases.WSEED + 90)
wav = synth.make_waveform(3, 90 * SR, cases.XSEED + 90); pm = O.lengths_to_padding_mask(torch.tensor([90 * SR, 60 * SR, 56 * SR + 3210]))
extra("ragged_90s_60s_56s (M=6748)", cfg, sd, wav.masked_fill(pm, 0.0), pm)
