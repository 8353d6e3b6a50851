"""Incremental (chunk-by-chunk, cached left context) mode against the reference's streaming driver
semantics: the reference re-encodes the whole prefix at every step (rain/simul/transducer_agent.py:138-167);
golden chunks come from exactly that loop run on the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import wav2vec_s_b200 as W
from wav2vec_s_b200.model import EncoderStream
from oracle import cases
from oracle import w2vs_oracle as O
from helpers import load_golden, case_inputs, valid_rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
STREAM = [n for n, c in cases.CASES.items() if c.get("api") == "stream"]


def build(cfg, sd, dtype):
    m = W.BlockWiseWav2Vec2Model(cfg)
    m.load_state_dict(sd, strict=False)
    return m.to("cuda", dtype).eval()


@pytest.mark.parametrize("name", STREAM)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_stream_equals_reference_prefix_recompute(name, dtype):
    g = load_golden(name)
    cfg, sd, wav, _, _ = case_inputs(name)
    m = build(cfg, sd, dtype)
    src = wav.cuda().to(dtype)
    st = m.open_stream(B=1, max_seconds=4.0, max_new_samples=20000)
    prefixes = g["prefix_samples"].tolist()
    outs, pos = [], 0
    for i, n in enumerate(prefixes):
        last = i == len(prefixes) - 1
        y = st.step(src[:, pos:n], EncoderStream.FINAL if last else EncoderStream.NONE)
        outs.append(y)
        pos = n
    assert [int(o.size(0)) for o in outs] == g["chunk_sizes"].tolist()     # same frames at the same steps
    y = torch.cat(outs, 0).cpu()
    assert valid_rel_err(y, g["y"]) < TOL[dtype]
    assert valid_rel_err(y, g["y_offline"]) < TOL[dtype]                    # incremental == offline


@pytest.mark.parametrize("name", STREAM)
def test_stream_irregular_chunks_equal_offline(name):
    cfg, sd, wav, _, _ = case_inputs(name)
    m = build(cfg, sd, torch.float32)
    src = wav.cuda()
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    rs = np.random.RandomState(11)
    st = m.open_stream(B=1, max_seconds=4.0, max_new_samples=9000)
    outs, pos, L = [], 0, src.size(1)
    while pos < L:
        n = min(int(rs.choice([1, 37, 399, 400, 401, 1600, 5120, 8999])), L - pos)
        outs.append(st.step(src[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0).cpu()
    assert tuple(y.shape) == tuple(ref.shape)
    assert valid_rel_err(y, ref) < 1e-4


@pytest.mark.parametrize("main,rc", [(8, 4), (32, 16), (20, 10), (12, 6), (16, 0)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_stream_other_block_sizes_equal_offline(main, rc, dtype):
    """Incremental mode with block sizes other than 16 / 8 (the reference's sampling range, wav2vec_S.py:392-395,
    and no look-ahead): irregular chunks, FINAL flush == the offline rain forward on the whole utterance."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True, main_context=main, right_context=rc)
    sd = synth.make_state_dict(cfg, 21 + main)
    wav = synth.make_waveform(1, 30000, 77 + main)
    m = build(cfg, sd, dtype)
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    rs = np.random.RandomState(main)
    st = m.open_stream(B=1, max_seconds=3.0, max_new_samples=9000)
    src = wav.cuda()
    outs, pos, L = [], 0, src.size(1)
    while pos < L:
        n = min(int(rs.choice([37, 400, 1600, 5120, 8999])), L - pos)
        outs.append(st.step(src[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0).cpu()
    assert tuple(y.shape) == tuple(ref.shape)
    assert valid_rel_err(y, ref) < TOL[dtype]


def test_stream_batch_lockstep():
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    from oracle import synth
    sd = synth.make_state_dict(cfg, 5)
    wav = synth.make_waveform(3, 20000, 99)
    m = build(cfg, sd, torch.float32)
    st = m.open_stream(B=3, max_seconds=3.0, max_new_samples=8000)
    outs, pos = [], 0
    for n in (7760, 5120, 5120, 2000):
        outs.append(st.step(wav[:, pos:pos + n].cuda(), EncoderStream.FINAL if pos + n >= 20000 else 0))
        pos += n
    y = torch.cat(outs, 0).cpu()
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    assert tuple(y.shape) == tuple(ref.shape)
    assert valid_rel_err(y, ref) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_incremental_forward_is_drop_in_for_prefix_recompute(dtype):
    """forward(prefix, incremental_state=dict, is_infer=True) returns what the reference returns for
    the same prefix (any prefix length, block aligned or not), computed from cached state."""
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    from oracle import synth
    sd = synth.make_state_dict(cfg, 5)
    wav = synth.make_waveform(1, 30000, 17)
    m = build(cfg, sd, dtype)
    state = {}
    prefixes = [7760, 8160, 12880, 13000, 17000, 23120, 23121, 28000, 30000]
    for i, n in enumerate(prefixes):
        fin = i == len(prefixes) - 1
        out = m(wav[:, :n].cuda().to(dtype), None, state, fin, True)
        y = out["encoder_out"][0]
        ref, pm = O.rain_forward(sd, cfg, wav[:, :n], None, finished=fin, is_infer=True)
        assert tuple(y.shape) == tuple(ref.shape), n
        assert tuple(out["encoder_padding_mask"][0].shape) == tuple(pm.shape)
        if ref.numel():
            assert valid_rel_err(y.cpu(), ref) < TOL[dtype], n


def test_stream_unsupported_modes_fail_loudly():
    cfg = cases.tiny(extractor_mode="default", pos_type="conv", encoder_layers=2)
    from oracle import synth
    m = build(cfg, synth.make_state_dict(cfg, 1), torch.float32)
    with pytest.raises(W.cabi.W2vsError) as e:
        m.open_stream(B=1, max_seconds=1.0)
    assert e.value.status == W.cabi.UNSUPPORTED


def test_stream_takes_pcm16_chunks():
    """16-bit PCM chunks (what an audio source delivers; the reference agent divides by 32768 on the host,
    rain/simul/transducer_searcher.py:74-80) give the same frames as the float waveform."""
    name = STREAM[0]
    cfg, sd, wav, _, _ = case_inputs(name)
    m = build(cfg, sd, torch.float32)
    pcm = (wav * 3000).clamp(-32768, 32767).to(torch.int16)
    x = O.waveform_frontend(pcm, False)
    ref, _ = O.rain_forward(sd, cfg, x, None, finished=True, is_infer=True)
    st = m.open_stream(B=1, max_seconds=4.0, max_new_samples=9000)
    outs, pos, L = [], 0, pcm.size(1)
    src = pcm.cuda()
    while pos < L:
        n = min(5120 if pos else 7760, L - pos)
        outs.append(st.step(src[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0).cpu()
    assert tuple(y.shape) == tuple(ref.shape)
    assert valid_rel_err(y, ref) < TOL[torch.float32]


@pytest.mark.parametrize("pre_ln", [True, False], ids=["preln", "postln"])
@pytest.mark.parametrize("B,main,rc", [(1, 16, 8), (2, 8, 4), (1, 20, 10), (3, 6, 2)])
def test_fused_step_kernel_equals_operator_chain(pre_ln, B, main, rc):
    """bf16 decision steps can run as one persistent cooperative kernel (k_stream_fused.cu, stream_step_impl = 2) when
    a step has at most 32 tokens; the kernel-per-operator chain (the default) computes the same rows with the same
    roundings, up to summation order.  Irregular chunks, so that steps of every size (full blocks, the short blocks of the final
    flush) and several blocks per call occur; both against each other and against the oracle's offline rows."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=pre_ln, conv_bias=pre_ln, main_context=main, right_context=rc,
                     encoder_layers=4)
    sd = synth.make_state_dict(cfg, 40 + main)
    L = 26000
    wav = synth.make_waveform(B, L, 50 + main)
    m = build(cfg, sd, torch.bfloat16)
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    src = wav.cuda()
    ys = []
    for impl in (2, 1):
        rs = np.random.RandomState(3)
        st = m.open_stream(B=B, max_seconds=3.0, max_new_samples=9000, step_impl=impl)
        W.cabi.launch_count(reset=True)
        outs, pos = [], 0
        while pos < L:
            n = min(int(rs.choice([37, 400, 1600, 5120, 8999])), L - pos)
            outs.append(st.step(src[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
            pos += n
        torch.cuda.synchronize()
        ys.append((torch.cat(outs, 0).float().cpu(), W.cabi.launch_count(reset=True)))
    (y_fused, n_fused), (y_chain, n_chain) = ys
    assert n_fused < n_chain                      # the fused path really ran (far fewer launches)
    assert tuple(y_fused.shape) == tuple(ref.shape)
    assert valid_rel_err(y_fused, y_chain) < 1e-2
    assert valid_rel_err(y_fused, ref) < TOL[torch.bfloat16] and valid_rel_err(y_chain, ref) < TOL[torch.bfloat16]


@pytest.mark.parametrize("main,rc", [(16, 8), (8, 4), (20, 10), (6, 2), (16, 0)])
def test_cluster_step_kernel_equals_operator_chain(main, rc):
    """One stream of a pre-LN bf16 model can run its decision steps as one kernel of thread-block clusters
    (k_stream_cluster.cu, stream_step_impl = 3: one cluster of 8 CTAs per attention head, two grid barriers per layer,
    fp32 reductions at L2 into the residual stream).  Same rows as the kernel-per-operator chain with the same
    roundings, up to summation order; irregular chunks, so that steps of every size occur (full blocks, the short
    blocks of the final flush, several blocks per call); both also against the oracle's offline rows."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True, main_context=main, right_context=rc, encoder_layers=4)
    sd = synth.make_state_dict(cfg, 70 + main)
    L = 26000
    wav = synth.make_waveform(1, L, 80 + main)
    m = build(cfg, sd, torch.bfloat16)
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    src = wav.cuda()
    ys = []
    for impl in (3, 1):
        rs = np.random.RandomState(3)
        st = m.open_stream(B=1, max_seconds=3.0, max_new_samples=9000, step_impl=impl)
        W.cabi.launch_count(reset=True)
        outs, pos = [], 0
        while pos < L:
            n = min(int(rs.choice([37, 400, 1600, 5120, 8999])), L - pos)
            outs.append(st.step(src[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
            pos += n
        torch.cuda.synchronize()
        ys.append((torch.cat(outs, 0).float().cpu(), W.cabi.launch_count(reset=True)))
    (y_cl, n_cl), (y_chain, n_chain) = ys
    assert n_cl < n_chain                      # the cluster kernel really ran (far fewer launches)
    assert tuple(y_cl.shape) == tuple(ref.shape)
    e_chain, e_ref = valid_rel_err(y_cl, y_chain), valid_rel_err(y_cl, ref)
    print(f"\n[parity] cluster step kernel main={main} rc={rc}: vs chain {e_chain:.3e}, vs oracle {e_ref:.3e}")
    assert e_chain < 1e-2
    assert e_ref < TOL[torch.bfloat16] and valid_rel_err(y_chain, ref) < TOL[torch.bfloat16]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_stream_grows_past_its_initial_capacity(dtype):
    """A stream opened for 0.5 s keeps going: its state moves into larger buffers (w2vs_stream_grow, K/V cache and
    frame buffer copied on the device) instead of returning INVALID_VALUE -- the reference driver re-encodes prefixes
    of any length up to --max-audio-positions.  Same frames as the offline forward."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 61)
    L = 40000                                   # 2.5 s = 124 frames; the stream starts with room for 26
    wav = synth.make_waveform(2, L, 62)
    m = build(cfg, sd, dtype)
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    st = m.open_stream(B=2, max_seconds=0.5, max_new_samples=6000)
    cap0 = st.max_frames
    src = wav.cuda()
    outs, pos = [], 0
    while pos < L:
        n = min(5120, L - pos)
        outs.append(st.step(src[:, pos:pos + n], EncoderStream.FINAL if pos + n >= L else EncoderStream.NONE))
        pos += n
    y = torch.cat(outs, 0).cpu()
    assert st.max_frames > cap0 and tuple(y.shape) == tuple(ref.shape)
    assert valid_rel_err(y, ref) < TOL[dtype]


def test_incremental_forward_from_host_prefix_uploads_new_samples_only():
    """The streaming driver keeps the growing prefix on the host (rain/simul/transducer_searcher.py:728-731 uploads all
    of it at every step); with incremental_state the model takes the host tensor as is and moves only the samples it
    has not seen: same result as feeding a device prefix."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 5)
    wav = synth.make_waveform(1, 30000, 17)                      # host tensor
    m = build(cfg, sd, torch.float32)
    s_host, s_dev = {}, {}
    for i, n in enumerate([7760, 12880, 13000, 23120, 30000]):
        fin = n == 30000
        y_h = m(wav[:, :n], None, s_host, fin, True)["encoder_out"][0]
        y_d = m(wav[:, :n].cuda(), None, s_dev, fin, True)["encoder_out"][0]
        assert y_h.is_cuda and torch.equal(y_h, y_d)
    ref, _ = O.rain_forward(sd, cfg, wav, None, finished=True, is_infer=True)
    assert valid_rel_err(y_h.cpu(), ref) < 1e-4
