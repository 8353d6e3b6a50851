"""GPU parity at the shapes BASELINE.json quotes and beyond them (round-1 verdict, item 1): the CUDA path through
the reference-shaped classes and the C ABI against the oracle on the same seeded inputs -- and against golden
vectors of the unmodified reference where tests/golden holds them -- at

  * configs[0] exactly: base, 1 x 10 s, fp32 (T = 499: the look-ahead window overruns T', 4 masked copies);
  * one utterance of configs[2] / configs[4]: large, 1 x 20 s and 1 x 30 s;
  * configs[1]'s utterance length, ragged: base, 2 x 15 s;
  * configs[3]: large, chunk-by-chunk incremental over 30 s against the oracle's offline rows;
  * utterances of 60-90 s in one ragged batch (M > 4096 tokens: more than 32 of the attention kernel's 128-token
    padding-flag blocks), which the reference allows (--max-audio-positions 3200000 = 200 s).

Tolerances are BASELINE.json's: max-abs-rel 1e-4 in fp32 mode against the fp32 oracle; 2e-2 in bf16 mode against the
fp32-arithmetic oracle on the identical (bf16-valued) weights, plus the comparison with the oracle on the original
fp32 weights at 2e-2 for the large model and at the measured ceiling of helpers.BF16_TOL_FP32_WEIGHTS for the base
model, whose weight rounding alone is 1.6e-2 (see helpers.py).  Masks bit-exact.  No fallbacks: above the bound
fails."""
import numpy as np
import pytest
import torch

import wav2vec_s_b200 as W
from wav2vec_s_b200.model import EncoderStream
from oracle import cases, synth
from oracle import w2vs_oracle as O
from helpers import load_golden, case_inputs, valid_rel_err, bf16_valued, FP32_TOL, BF16_TOL, BF16_TOL_FP32_WEIGHTS

pytestmark = pytest.mark.gpu
TOL = {torch.float32: FP32_TOL, torch.bfloat16: BF16_TOL}
IDS = ["fp32", "bf16"]
DTYPES = [torch.float32, torch.bfloat16]
SR = 16000
_cache = {}


def oracle_once(key, fn):
    """The CPU oracle takes seconds at these sizes: one run per case, shared by the dtype variants."""
    if key not in _cache:
        _cache[key] = fn()
    return _cache[key]


def build(cls, cfg, sd, dtype):
    m = cls(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {"mask_emb"}
    return m.to("cuda", dtype).eval()


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
@pytest.mark.parametrize("name", ["cfg1_base_10s", "large_20s"])
def test_baseline_config_vs_reference_golden_and_oracle(name, dtype):
    """configs[0] (base 1 x 10 s) and one utterance of configs[2] (large 20 s): against the unmodified reference's
    output (golden, every k-th frame) and against the oracle on every frame."""
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    yo = oracle_once(name, lambda: O.extract_features(sd, cfg, wav, pm)[0])
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    y, fm = m.extract_features(wav.cuda(), None)
    assert fm is None and list(y.shape) == g["y_shape"].tolist() and y.dtype == dtype
    k = cases.CASES[name]["compact"]
    e_gold = valid_rel_err(y[:, ::k].cpu(), g["y"])
    e_orac = valid_rel_err(y.cpu(), yo)
    if dtype == torch.float32:
        print(f"\n[parity] {name} fp32: vs reference golden {e_gold:.3e}, vs oracle {e_orac:.3e}")
        assert e_gold < FP32_TOL and e_orac < FP32_TOL
        return
    yq = oracle_once((name, "bf16w"), lambda: O.extract_features(bf16_valued(sd), cfg, wav, pm)[0])
    e_same_w = valid_rel_err(y.cpu(), yq)
    print(f"\n[parity] {name} bf16: vs oracle on identical (bf16-valued) weights {e_same_w:.3e}; on fp32 weights: "
          f"reference golden {e_gold:.3e}, oracle {e_orac:.3e}")
    assert e_same_w < BF16_TOL
    wtol = BF16_TOL if cfg["encoder_layers"] == 24 else BF16_TOL_FP32_WEIGHTS
    assert e_gold < wtol and e_orac < wtol


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
def test_large_30s_vs_oracle(dtype):
    """One utterance of configs[4] (large, 30 s: T = 1499, M = 2244)."""
    cfg = O.large_cfg()
    sd = synth.make_state_dict(cfg, cases.WSEED)
    wav = synth.make_waveform(1, 30 * SR, cases.XSEED + 30)
    yo = oracle_once("large_30s", lambda: O.extract_features(sd, cfg, wav, None)[0])
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    y, _ = m.extract_features(wav.cuda(), None)
    err = valid_rel_err(y.cpu(), yo)
    print(f"\n[parity] large_30s {dtype}: vs oracle {err:.3e}")
    assert tuple(y.shape) == (1, 1499, 1024) and err < TOL[dtype]
    if dtype == torch.bfloat16:
        yq = oracle_once("large_30s_bf16w", lambda: O.extract_features(bf16_valued(sd), cfg, wav, None)[0])
        e2 = valid_rel_err(y.cpu(), yq)
        print(f"[parity] large_30s bf16: vs oracle on identical (bf16-valued) weights {e2:.3e}")
        assert e2 < BF16_TOL


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
def test_base_2x15s_ragged_vs_oracle(dtype):
    """configs[1]'s utterance length in a ragged pair (base, 15 s: T = 749, M = 1118), bf16 samples in bf16 mode
    as the reference trainer feeds them (trainer.py:1120-1129)."""
    cfg = O.base_cfg()
    sd = synth.make_state_dict(cfg, cases.WSEED)
    wav = synth.make_waveform(2, 15 * SR, cases.XSEED + 15)
    lens = torch.tensor([15 * SR, 9 * SR + 1234])
    pm = O.lengths_to_padding_mask(lens)
    wav = wav.masked_fill(pm, 0.0)
    src = wav.to(dtype)
    yo, fmo = oracle_once(("base_2x15s", dtype), lambda: O.extract_features(sd, cfg, src.float(), pm))
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    y, fm = m.extract_features(src.cuda(), pm.cuda())
    assert torch.equal(fm.cpu(), fmo)
    err = valid_rel_err(y.cpu(), yo, fmo.numpy())
    print(f"\n[parity] base_2x15s_ragged {dtype}: vs oracle (fp32 weights) {err:.3e}")
    if dtype == torch.float32:
        assert err < FP32_TOL
        return
    yq, _ = oracle_once(("base_2x15s", "bf16w"), lambda: O.extract_features(bf16_valued(sd), cfg, src.float(), pm))
    e2 = valid_rel_err(y.cpu(), yq, fmo.numpy())
    print(f"[parity] base_2x15s_ragged bf16: vs oracle on identical (bf16-valued) weights {e2:.3e}")
    assert e2 < BF16_TOL and err < BF16_TOL_FP32_WEIGHTS


@pytest.mark.parametrize("dtype,step_impl", [(torch.float32, 0), (torch.bfloat16, 1), (torch.bfloat16, 2), (torch.bfloat16, 0)],
                         ids=["fp32", "bf16-chain", "bf16-persistent-kernel", "bf16-default-cluster-kernel"])
def test_incremental_large_30s_vs_oracle_offline_rows(dtype, step_impl):
    """configs[3]: large, chunk by chunk (first chunk 24 frames = 7760 samples, then 16 frames = 5120 samples) with
    cached left context over 30 s; every emitted frame against the oracle's OFFLINE row of the complete utterance
    (what the reference's prefix re-encoding converges to, rain/simul/transducer_agent.py:138-167)."""
    cfg = O.large_cfg()
    sd = synth.make_state_dict(cfg, cases.WSEED)
    wav = synth.make_waveform(1, 30 * SR, cases.XSEED + 30)
    yo = oracle_once("large_30s", lambda: O.extract_features(sd, cfg, wav, None)[0])      # [1, T, D]
    m = build(W.BlockWiseWav2Vec2Model, cfg, sd, dtype)
    st = m.open_stream(B=1, max_seconds=31, max_new_samples=7760 + 400, step_impl=step_impl)
    dev = wav.cuda()
    outs, pos, L = [], 0, wav.size(1)
    while pos < L:
        n = min(7760 if pos == 0 else 5120, L - pos)
        outs.append(st.step(dev[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0).transpose(0, 1)                                                 # [1, T, D]
    assert tuple(y.shape) == tuple(yo.shape)
    err = valid_rel_err(y.cpu(), yo)
    print(f"\n[parity] incremental large_30s {dtype} step_impl={step_impl}: vs oracle offline rows {err:.3e}")
    assert err < TOL[dtype]


def _long_cfg(pre_ln):
    # narrow model with the real head size (64), so that the tcgen05 attention kernel runs; small convs keep the
    # CPU oracle at a few seconds for 90 s of audio
    return cases.tiny(encoder_layers=2, encoder_embed_dim=256, encoder_ffn_embed_dim=512, encoder_attention_heads=4,
                      layer_norm_first=pre_ln, conv_bias=pre_ln)


@pytest.mark.parametrize("dtype", DTYPES, ids=IDS)
@pytest.mark.parametrize("pre_ln", [True, False], ids=["preln", "postln"])
def test_ragged_60s_90s_vs_oracle(pre_ln, dtype):
    """A ragged batch of 90 s, 60 s and 56.2 s utterances: T = 4499, M = 6748 tokens, i.e. 53 padding-flag blocks of
    128 tokens per utterance -- padded main keys AND padded look-ahead copies beyond token 4096 (the round-1 kernel
    shifted its 32-bit flag mask by >= 32 there and never read those padding bytes)."""
    cfg = _long_cfg(pre_ln)
    sd = synth.make_state_dict(cfg, cases.WSEED + 90)
    L = 90 * SR
    wav = synth.make_waveform(3, L, cases.XSEED + 90)
    lens = torch.tensor([L, 60 * SR, 56 * SR + 3210])
    pm = O.lengths_to_padding_mask(lens)
    wav = wav.masked_fill(pm, 0.0)
    yo, fmo = oracle_once(("long", pre_ln), lambda: O.extract_features(sd, cfg, wav, pm))
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    assert m.geometry(L).tokens > 4096
    y, fm = m.extract_features(wav.cuda(), pm.cuda())
    assert torch.equal(fm.cpu(), fmo)
    # per utterance, so that a failure names the one with padded keys beyond token 4096
    for b in range(3):
        keep = ~fmo[b]
        err = float((y[b].float().cpu()[keep] - yo[b][keep]).abs().max() / yo[b][keep].abs().max())
        print(f"\n[parity] ragged_90s pre_ln={pre_ln} {dtype} utterance {b} ({int(lens[b]) / SR:.1f} s): {err:.3e}")
        assert err < TOL[dtype], (b, err)


def test_incremental_large_16_streams_vs_oracle_offline_rows():
    """configs[3] with a batch of streams: 16 lock-step streams of the large model (384 tokens per decision step: the
    tcgen05 GEMM path of the incremental mode), 6 s each, bf16, every emitted frame against the oracle's offline rows
    of the same utterances (identical, bf16-valued weights)."""
    cfg = O.large_cfg()
    sd = synth.make_state_dict(cfg, cases.WSEED)
    B, L = 16, 6 * SR
    wav = synth.make_waveform(B, L, cases.XSEED + 16)
    yq = O.extract_features(bf16_valued(sd), cfg, wav, None)[0]                       # [B, T, D]
    m = build(W.BlockWiseWav2Vec2Model, cfg, sd, torch.bfloat16)
    st = m.open_stream(B=B, max_seconds=7, max_new_samples=7760 + 400)
    dev = wav.cuda()
    outs, pos = [], 0
    while pos < L:
        n = min(7760 if pos == 0 else 5120, L - pos)
        outs.append(st.step(dev[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0).transpose(0, 1)
    assert tuple(y.shape) == tuple(yq.shape)
    err = valid_rel_err(y.cpu(), yq)
    print(f"\n[parity] incremental large 16 streams bf16: vs oracle offline rows (identical weights) {err:.3e}")
    assert err < BF16_TOL
