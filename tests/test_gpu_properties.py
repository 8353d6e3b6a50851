"""Size-independent properties of the CUDA path at the model sizes BASELINE.json quotes (wav2vec-S large, 24 layers,
20 s utterances), where the CPU oracle would take minutes per case: determinism, independence of the utterances of
a batch, causality of the block mask (a frame never depends on audio beyond its block's look-ahead), and agreement
of the full-utterance and the incremental path.  Bit-exact where the arithmetic is identical by construction."""
import pytest
import torch

import wav2vec_s_b200 as W
from wav2vec_s_b200.model import EncoderStream

pytestmark = pytest.mark.gpu
LARGE = dict(extractor_mode="layer_norm", encoder_layers=24, encoder_embed_dim=1024, encoder_ffn_embed_dim=4096,
             encoder_attention_heads=16, layer_norm_first=True, conv_bias=True, pos_type="sin",
             main_context=16, right_context=8)
SR = 16000


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return W.BlockWiseWav2Vec2Model(LARGE).to("cuda", torch.bfloat16).eval()


def _wav(B, seconds, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, seconds * SR, generator=g).cuda()


def test_deterministic_and_batch_order_invariant(model):
    wav = _wav(6, 20, 1)
    y1 = model.extract_features(wav, None)[0]
    y2 = model.extract_features(wav, None)[0]
    assert torch.equal(y1, y2)                                   # no atomics, fixed reduction orders
    perm = torch.tensor([3, 0, 5, 1, 4, 2], device="cuda")
    y3 = model.extract_features(wav[perm], None)[0]
    assert torch.equal(y3, y1[perm])                             # utterances do not see each other
    assert torch.isfinite(y1.float()).all()


def test_block_mask_causality_at_full_length(model):
    """Changing the audio from second 15 on leaves every frame whose block and look-ahead end before that point
    bit-identical; later frames do change (the mask is not trivially empty)."""
    wav = _wav(2, 20, 2)
    y = model.extract_features(wav, None)[0]
    wav2 = wav.clone()
    cut = 15 * SR
    wav2[:, cut:] = _wav(2, 5, 3)
    y2 = model.extract_features(wav2, None)[0]
    first_changed_frame = (cut - 400) // 320 + 1                 # first conv frame whose receptive field reaches `cut`
    safe = (first_changed_frame - 8) // 16 * 16                  # blocks whose 8 look-ahead frames are untouched
    assert safe > 600
    assert torch.equal(y[:, :safe], y2[:, :safe])
    assert not torch.equal(y[:, safe + 32:], y2[:, safe + 32:])


def test_ragged_batch_matches_single_utterances(model):
    """An utterance encoded inside a longer, padded batch equals the same utterance encoded alone (on its valid
    frames; the look-ahead tokens sit at different offsets, so the sums are reordered: bf16 tolerance)."""
    wav = _wav(3, 20, 4)
    lens = torch.tensor([20 * SR, 13 * SR + 137, 7 * SR + 3999], device="cuda")
    pm = torch.arange(wav.size(1), device="cuda")[None, :] >= lens[:, None]
    wav = wav.masked_fill(pm, 0.0)
    y, fm = model.extract_features(wav, pm)
    for b in range(3):
        n = int(lens[b])
        yb = model.extract_features(wav[b:b + 1, :n], None)[0][0]
        valid = int((~fm[b]).sum())
        t = min(valid, yb.size(0))
        assert abs(valid - yb.size(0)) <= 1                      # mask arithmetic (ceil(len/w)) vs conv arithmetic
        err = (y[b, :t].float() - yb[:t].float()).abs().max() / yb.float().abs().max()
        assert err < 2e-2, (b, float(err))


def test_incremental_equals_full_utterance_at_30s(model):
    """Chunk-by-chunk with cached left context == one full-utterance call (rain is_infer, finished), 30 s."""
    wav = _wav(1, 30, 5)
    ref = model(wav, None, finished=True, is_infer=True)["encoder_out"][0]           # [T, 1, D]
    st = model.open_stream(B=1, max_seconds=31, max_new_samples=7760 + 400)
    outs, pos, L = [], 0, wav.size(1)
    while pos < L:
        n = min(7760 if pos == 0 else 5120, L - pos)
        outs.append(st.step(wav[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0)
    assert tuple(y.shape) == tuple(ref.shape)
    err = (y.float() - ref.float()).abs().max() / ref.float().abs().max()
    assert err < 2e-2, float(err)
    # A second stream over the same audio.  The default path of one large-model stream (the cluster step kernel) adds
    # the partial products of its clusters into the residual stream with fp32 reductions at L2, in whatever order
    # they arrive: the sums differ in the last bit, a bf16 rounding flips somewhere in the 24 layers and two runs differ
    # by a bf16 step or two at the output (measured 7e-3 of the output range) -- both inside the parity bar.  The
    # operator chain (step_impl = 1) gives the same bits every time: its split-key attention merges its partial
    # states in a fixed order, the completion counters only elect the CTA that does it.
    def again(step_impl):
        st2 = model.open_stream(B=1, max_seconds=31, max_new_samples=7760 + 400, step_impl=step_impl)
        outs2, pos = [], 0
        while pos < L:
            n = min(7760 if pos == 0 else 5120, L - pos)
            outs2.append(st2.step(wav[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
            pos += n
        return torch.cat(outs2, 0)
    scale = ref.float().abs().max()
    y2 = again(0)
    assert float((y2.float() - ref.float()).abs().max() / scale) < 2e-2
    assert float((y2.float() - y.float()).abs().max() / scale) < 1.5e-2
    c1, c2 = again(1), again(1)
    assert torch.equal(c1, c2)
    assert float((c1.float() - ref.float()).abs().max() / scale) < 2e-2
