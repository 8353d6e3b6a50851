"""The oracle restatement (oracle/w2vs_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import w2vs_oracle as O
from helpers import load_golden, case_inputs, valid_rel_err

FP32_TOL = 2e-5  # oracle and reference run the same torch ops; only op ordering may differ

FAIRSEQ = [n for n, c in cases.CASES.items() if c.get("api", "fairseq") == "fairseq" and not c.get("compact")]
COMPACT = [n for n, c in cases.CASES.items() if c.get("compact")]
RAIN = [n for n, c in cases.CASES.items() if c.get("api") == "rain"]
STREAM = [n for n, c in cases.CASES.items() if c.get("api") == "stream"]


@pytest.mark.parametrize("name", FAIRSEQ)
def test_fairseq_api(name):
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    assert cfg == g["cfg"]
    taps = {}
    y, fm = O.extract_features(sd, cfg, wav, pm, taps=taps)
    assert tuple(y.shape) == g["y"].shape
    if g["fmask"].size:
        assert np.array_equal(fm.numpy(), g["fmask"])          # bool structure: bit exact
    else:
        assert fm is None
    n_conv = len(O.conv_layers_of(cfg))
    assert valid_rel_err(taps[f"conv{n_conv-1}"], g["conv_out"]) < FP32_TOL
    if "post_proj" in g:
        # the reference zeroes padded frames of this tensor in place (index_put, wav2vec_S.py:358)
        assert valid_rel_err(taps["post_proj"], g["post_proj"], g["fmask"]) < FP32_TOL
    assert valid_rel_err(y, g["y"], g["fmask"]) < FP32_TOL


@pytest.mark.parametrize("name", RAIN)
def test_rain_api(name):
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    y, fm = O.rain_forward(sd, cfg, wav, pm, **cases.CASES[name].get("kwargs", {}))
    assert tuple(y.shape) == g["y"].shape
    assert np.array_equal(fm.numpy(), g["fmask"])
    assert valid_rel_err(y, g["y"], g["fmask"], time_first=True) < FP32_TOL


@pytest.mark.parametrize("name", STREAM)
def test_stream_prefix_recompute(name):
    g = load_golden(name)
    cfg, sd, wav, _, _ = case_inputs(name)
    chunks = O.streaming_prefix_recompute(sd, cfg, wav)
    assert [c[1].size(0) for c in chunks] == g["chunk_sizes"].tolist()
    assert [c[0] for c in chunks] == g["prefix_samples"].tolist()
    y = torch.cat([c[1] for c in chunks], dim=0)
    assert valid_rel_err(y, g["y"]) < FP32_TOL
    # the property incremental mode relies on (SURVEY.md section 5 "long-context"):
    # committed frames of each 16k+8 prefix equal the offline rows
    n_full = sum(g["chunk_sizes"].tolist()[:-1])
    assert valid_rel_err(g["y"][:n_full], g["y_offline"][:n_full]) < 1e-5


def test_mask_structure_counts():
    # cfg3 M=1496 -> 769 856 visible pairs of 2 238 016
    rc_idx, oor, masked = O.block_mask_structure(1000, 16, 8)
    assert masked.shape == (1496, 1496)
    assert int((~masked).sum()) == 769856
    # cfg1: T'=500 -> 4 out-of-range rc copies
    rc_idx, oor, masked = O.block_mask_structure(500, 16, 8)
    assert int(oor.sum()) == 4 and masked.shape == (748, 748)


def test_frame_padding_formula():
    # valid frames = min(T, ceil(len / floor(L/T))), not the conv formula
    L, T = 32000, 99
    lens = torch.tensor([28280, 32000, 1000])
    fm = O.frame_padding_mask(O.lengths_to_padding_mask(lens), T)
    w = L // T
    assert (~fm).sum(1).tolist() == [min(T, -(-int(l) // w)) for l in lens]


def test_waveform_frontend_restatement():
    """oracle.waveform_frontend against the two reference statements it restates, evaluated directly."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    pcm = (torch.randn(2, 5000, generator=g) * 2000).to(torch.int16)
    x = O.waveform_frontend(pcm, False)
    assert x.dtype == torch.float32
    assert torch.equal(x, torch.from_numpy((pcm.numpy() / 32768.0).astype("float32")))      # transducer_searcher.py:76-78
    lens = torch.tensor([5000, 3210])
    y = O.waveform_frontend(pcm, True, lens)
    for b in range(2):
        n = int(lens[b])
        ref = F.layer_norm(x[b, :n], x[b, :n].shape)                                       # raw_audio_dataset.py:69-72
        assert torch.equal(y[b, :n], ref)
        assert torch.equal(y[b, n:], x[b, n:])


@pytest.mark.parametrize("name", COMPACT)
def test_baseline_shapes_compact(name):
    """BASELINE.json shapes (cfg1 base 1 x 10 s exactly; one large 20 s utterance of configs[2]): the oracle against
    the unmodified reference's output, stored every k-th frame (oracle/cases.py)."""
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    assert cfg == g["cfg"]
    y, fm = O.extract_features(sd, cfg, wav, pm)
    assert list(y.shape) == g["y_shape"].tolist() and fm is None and g["fmask"].size == 0
    k = cases.CASES[name]["compact"]
    assert valid_rel_err(y[:, ::k], g["y"]) < FP32_TOL
