"""The oracle against the LIVE reference (its unmodified source files executed through oracle/ref_shim.py) on inputs
the golden vectors do not cover: fresh seeds, other lengths, the GroupNorm / conv-position variants.  Needs
/root/reference, which exists in the build container only -- skipped elsewhere (the committed golden vectors of
tests/test_oracle_golden.py are the pin that travels)."""
import warnings

import pytest
import torch

from oracle import cases, ref_shim, synth
from oracle import w2vs_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")
warnings.filterwarnings("ignore")

VARIANTS = {
    "preln_ragged": dict(cfg=cases.tiny(layer_norm_first=True, conv_bias=True), B=3, L=9137, ragged=True),
    "postln_odd": dict(cfg=cases.tiny(encoder_layers=12), B=2, L=7013, ragged=False),
    "groupnorm_posconv": dict(cfg=cases.tiny(extractor_mode="default", pos_type="conv", encoder_layers=2), B=2, L=6400, ragged=True),
    "ctx_8_4": dict(cfg=cases.tiny(main_context=8, right_context=4), B=2, L=8000, ragged=False),
    "rc0": dict(cfg=cases.tiny(right_context=0, layer_norm_first=True), B=1, L=5000, ragged=False),
    # block sizes the reference draws with context_type="sampling" (wav2vec_S.py:392-395): even main in 8..32,
    # even right context <= main / 2 -- not powers of two
    "ctx_20_10": dict(cfg=cases.tiny(main_context=20, right_context=10, layer_norm_first=True), B=2, L=16000, ragged=True),
    "ctx_12_6": dict(cfg=cases.tiny(main_context=12, right_context=6), B=2, L=9000, ragged=False),
    "ctx_30_14": dict(cfg=cases.tiny(main_context=30, right_context=14, layer_norm_first=True), B=1, L=24000, ragged=False),
    "ctx_10_4": dict(cfg=cases.tiny(main_context=10, right_context=4), B=3, L=7000, ragged=True),
}


def _inputs(v, seed):
    sd = synth.make_state_dict(v["cfg"], seed)
    wav = synth.make_waveform(v["B"], v["L"], seed + 1)
    pm = None
    if v["ragged"]:
        pm = O.lengths_to_padding_mask(synth.make_lengths(v["B"], v["L"], seed + 2))
        wav = wav.masked_fill(pm, 0.0)
    return sd, wav, pm


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_fairseq_extract_features(name):
    v = VARIANTS[name]
    sd, wav, pm = _inputs(v, 100)
    with torch.no_grad():
        m = ref_shim.build_fairseq_model(v["cfg"])
        m.load_state_dict(sd, strict=False)
        y_ref, fm_ref = m.extract_features(wav.clone(), pm)
    y, fm = O.extract_features(sd, v["cfg"], wav, pm)
    assert (fm is None) == (fm_ref is None)
    if fm is not None:
        assert torch.equal(fm, fm_ref)
    assert O.max_abs_rel(y, y_ref) < 2e-5


@pytest.mark.parametrize("main,rc", [(16, 8), (8, 4), (32, 16), (20, 10), (12, 6), (16, 0)])
@pytest.mark.parametrize("finished", [False, True])
def test_rain_forward_infer(finished, main, rc):
    cfg = cases.tiny(layer_norm_first=True, main_context=main, right_context=rc)
    sd, wav, _ = _inputs(dict(cfg=cfg, B=1, L=12880 + 320 * 5, ragged=False), 200)
    with torch.no_grad():
        m = ref_shim.build_rain_model(cfg)
        m.load_state_dict(sd, strict=False)
        o = m(wav.clone(), None, None, finished, True)
    y, fm = O.rain_forward(sd, cfg, wav, None, finished=finished, is_infer=True)
    assert tuple(y.shape) == tuple(o["encoder_out"][0].shape)
    assert O.max_abs_rel(y, o["encoder_out"][0]) < 2e-5
