"""GPU parity: the CUDA path (through the reference-shaped host classes and the C ABI) against
(a) golden vectors produced by the unmodified reference and (b) the oracle's per-stage taps, on
the seeded cases of oracle/cases.py.  Tolerances are the ones BASELINE.json states: max-abs-rel
1e-4 in fp32 mode, 2e-2 in bf16 mode; bool/int structure bit-exact."""
import numpy as np
import pytest
import torch

import wav2vec_s_b200 as W
from oracle import cases
from oracle import w2vs_oracle as O
from helpers import load_golden, case_inputs, valid_rel_err, bf16_valued, FP32_TOL, BF16_TOL, BF16_TOL_FP32_WEIGHTS

pytestmark = pytest.mark.gpu
# Intermediate taps (conv stack, projection, one layer) are diagnostics that name the stage when an output check
# fails; the contract (BASELINE.json) is on the encoder OUTPUT.  In bf16 mode the taps of the un-normalised
# stages sit at 1-2e-2 from operand rounding alone (the reference's own bf16 run is 2.1-2.7e-2 off its fp32
# output on these cases), so they get 1.5x the output bound.
BF16_STAGE_TOL = 3e-2

FAIRSEQ = [n for n, c in cases.CASES.items() if c.get("api", "fairseq") == "fairseq" and not c.get("compact")]
COMPACT = [n for n, c in cases.CASES.items() if c.get("compact")]
RAIN = [n for n, c in cases.CASES.items() if c.get("api") == "rain"]


def build(cls, cfg, sd, dtype):
    m = cls(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {"mask_emb"}
    return m.to("cuda", dtype).eval()


@pytest.mark.parametrize("name", FAIRSEQ)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_extract_features_vs_reference_golden(name, dtype):
    g = load_golden(name)
    cfg, sd, wav, pm, _ = case_inputs(name)
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    stol = FP32_TOL if dtype == torch.float32 else BF16_STAGE_TOL
    taps = {}
    # identical waveform to the one the golden vectors were made from (fp32 samples; conv0 reads fp32 or bf16
    # samples in either mode).  bf16-quantised samples are covered by test_bf16_waveform_vs_oracle below.
    src = wav.cuda()
    y, fm = m._encode(src, padding_mask=None if pm is None else pm.cuda(), taps=taps)
    y2, fm2 = m.extract_features(src, None if pm is None else pm.cuda())
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    assert tuple(y.shape) == g["y"].shape and y.dtype == dtype
    if g["fmask"].size:
        assert np.array_equal(fm.cpu().numpy(), g["fmask"])       # bit exact
        assert np.array_equal(fm2.cpu().numpy(), g["fmask"])
    else:
        assert fm is None and fm2 is None
    fmask = g["fmask"]
    # stage by stage, so a failure names the stage
    conv_ref = torch.from_numpy(g["conv_out"]).transpose(1, 2)    # [B,T,C]
    assert valid_rel_err(taps["conv_out"].cpu(), conv_ref) < stol, "conv stack"
    if "post_proj" in g:
        assert valid_rel_err(taps["post_proj"].cpu(), g["post_proj"], fmask) < stol, "post_extract_proj"
    T2 = m.geometry(wav.size(1)).frames_pad
    l0 = taps["layers"][0, :, :T2].cpu().transpose(0, 1)           # [T2,B,D] like the reference hook
    l0_ref = torch.from_numpy(g["layer0"])[:T2]
    fm_t2 = None
    if fmask.size:
        fm_t2 = np.pad(fmask, ((0, 0), (0, T2 - fmask.shape[1])), constant_values=True)
    elif T2 != g["y"].shape[1]:
        fm_t2 = np.zeros((y.size(0), T2), dtype=bool)
        fm_t2[:, g["y"].shape[1]:] = True
    assert valid_rel_err(l0, l0_ref, fm_t2, time_first=True) < stol, "encoder layer 0"
    assert valid_rel_err(y.cpu(), g["y"], fmask) < tol, "encoder output"


@pytest.mark.parametrize("name", FAIRSEQ)
def test_bf16_waveform_vs_oracle(name):
    """bf16 mode fed with bf16 samples (what the reference trainer does to the batch, trainer.py:1120-1129), against
    the reference algorithm in fp32 arithmetic on the identical inputs: the bf16-valued waveform and the bf16-valued
    weights the model holds (2e-2, BASELINE.json), and on the original fp32 weights (helpers.BF16_TOL_FP32_WEIGHTS)."""
    cfg, sd, wav, pm, _ = case_inputs(name)
    m = build(W.Wav2VecSModel, cfg, sd, torch.bfloat16)
    src = wav.to(torch.bfloat16)
    y, fm = m.extract_features(src.cuda(), None if pm is None else pm.cuda())
    yo, fmo = O.extract_features(sd, cfg, src.float(), pm)
    yq, _ = O.extract_features(bf16_valued(sd), cfg, src.float(), pm)
    assert y.dtype == torch.bfloat16
    if fmo is None:
        assert fm is None
    else:
        assert torch.equal(fm.cpu(), fmo)
    fmask = None if fmo is None else fmo.numpy()
    assert valid_rel_err(y.cpu(), yq, fmask) < BF16_TOL, "identical (bf16-valued) weights"
    assert valid_rel_err(y.cpu(), yo, fmask) < BF16_TOL_FP32_WEIGHTS, "original fp32 weights"


@pytest.mark.parametrize("name", FAIRSEQ)
def test_all_layers_vs_oracle_fp32(name):
    """Every layer's residual stream (main tokens) against the oracle, fp32."""
    cfg, sd, wav, pm, _ = case_inputs(name)
    m = build(W.Wav2VecSModel, cfg, sd, torch.float32)
    taps, otaps = {}, {}
    y, fm = m._encode(wav.cuda(), padding_mask=None if pm is None else pm.cuda(), taps=taps)
    yo, fmo = O.extract_features(sd, cfg, wav, pm, taps=otaps)
    T2 = otaps["enc_in"].size(0)
    fm_t2 = None
    if fmo is not None:
        fm_t2 = torch.nn.functional.pad(fmo, (0, T2 - fmo.size(1)), value=True).numpy()
    elif T2 != yo.size(1):
        fm_t2 = np.zeros((yo.size(0), T2), dtype=bool)
        fm_t2[:, yo.size(1):] = True
    assert valid_rel_err(taps["enc_in"][:, :T2].cpu().transpose(0, 1), otaps["enc_in"], fm_t2, time_first=True) < FP32_TOL
    for n in range(cfg["encoder_layers"]):
        ours = taps["layers"][n, :, :T2].cpu().transpose(0, 1)
        assert valid_rel_err(ours, otaps[f"layer{n}"], fm_t2, time_first=True) < FP32_TOL, f"layer {n}"
    assert valid_rel_err(y.cpu(), yo, None if fmo is None else fmo.numpy()) < FP32_TOL


@pytest.mark.parametrize("main,rc,L", [(20, 10, 16000), (12, 6, 9000), (30, 14, 24000), (10, 4, 7000), (26, 12, 40000)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_sampling_range_contexts_vs_oracle(main, rc, L, dtype):
    """Block sizes the reference draws with context_type="sampling" (wav2vec_S.py:392-395: even main context in
    8..32, even right context <= main / 2) -- not powers of two, so block boundaries fall inside the kernels' 64- and
    128-token tiles.  Ragged batch, against the oracle (itself checked against the live reference on these sizes in
    tests/test_oracle_vs_reference.py)."""
    from oracle import synth
    cfg = cases.tiny(main_context=main, right_context=rc, layer_norm_first=(main % 4 == 0))
    B = 3
    sd = synth.make_state_dict(cfg, cases.WSEED + main)
    wav = synth.make_waveform(B, L, cases.XSEED + main)
    pm = O.lengths_to_padding_mask(synth.make_lengths(B, L, cases.LSEED + main))
    wav = wav.masked_fill(pm, 0.0)
    yo, fmo = O.extract_features(sd, cfg, wav, pm)
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    y, fm = m.extract_features(wav.cuda(), pm.cuda())
    assert torch.equal(fm.cpu(), fmo)
    assert valid_rel_err(y.cpu(), yo, fmo.numpy()) < (FP32_TOL if dtype == torch.float32 else BF16_TOL)


def test_context_type_sampling_draws_like_reference():
    """context_type="sampling": the host draws (main, right) per call with Python's `random`, exactly as
    wav2vec_S.py:392-395 does, and the encoder runs with the drawn block sizes."""
    import random
    from oracle import synth
    cfg = cases.tiny(context_type="sampling", layer_norm_first=True)
    sd = synth.make_state_dict(cfg, cases.WSEED)
    wav = synth.make_waveform(2, 20000, cases.XSEED)
    m = build(W.Wav2VecSModel, cfg, sd, torch.float32)
    for seed in (3, 11, 12345):
        random.seed(seed)
        y, _ = m.extract_features(wav.cuda(), None)
        random.seed(seed)
        main = random.randint(4, 16) * 2
        rc = min(random.randint(2, 8) * 2, main // 2)
        yo, _ = O.extract_features(sd, dict(cfg, context_type="constant", main_context=main, right_context=rc), wav, None)
        assert valid_rel_err(y.cpu(), yo) < FP32_TOL, (seed, main, rc)


@pytest.mark.parametrize("embed_dim,heads", [(768, 12), (1024, 16)], ids=["Dg48", "Dg64"])
def test_posconv_tensor_core_path_vs_oracle(embed_dim, heads):
    """pos_type="conv" at the real widths (Conv1d(D, D, 128, groups=16), group width 48 / 64): in bf16 mode every
    group runs as an implicit GEMM on the tcgen05 kernel (group-major bf16 copy of the frames, channels padded to
    64).  Checked against the oracle (wav2vec2.py:791-804 restated) at the encoder input -- features + GELU(conv) --
    and at the output, ragged batch so that zeroed padded frames and the SamePad borders matter; the fp32 mode
    (direct kernel) runs beside it at its own tolerance."""
    from oracle import synth
    cfg = O.default_cfg(extractor_mode="layer_norm", pos_type="conv", encoder_layers=2, encoder_embed_dim=embed_dim,
                        encoder_ffn_embed_dim=2 * embed_dim, encoder_attention_heads=heads, layer_norm_first=True)
    B, L = 3, 40000
    sd = synth.make_state_dict(cfg, cases.WSEED)
    wav = synth.make_waveform(B, L, cases.XSEED)
    lens = synth.make_lengths(B, L, cases.LSEED)
    pm = O.lengths_to_padding_mask(lens)
    wav = wav.masked_fill(pm, 0.0)
    otaps = {}
    yo, fmo = O.extract_features(sd, cfg, wav, pm, taps=otaps)
    T2 = otaps["enc_in"].size(0)
    fm_t2 = torch.nn.functional.pad(fmo, (0, T2 - fmo.size(1)), value=True).numpy()
    T = yo.size(1)
    keep = (~fmo)[:, :, None]                                                     # [B,T,1] real frames
    pos_ref = (otaps["enc_in"][:T].transpose(0, 1) - otaps["post_proj"]) * keep   # the GELU(conv) term alone
    for dtype, tol, stol in ((torch.float32, FP32_TOL, FP32_TOL), (torch.bfloat16, BF16_TOL, BF16_STAGE_TOL)):
        m = build(W.Wav2VecSModel, cfg, sd, dtype)
        taps = {}
        y, fm = m._encode(wav.cuda(), padding_mask=pm.cuda(), taps=taps)
        assert torch.equal(fm.cpu(), fmo)
        e_in = taps["enc_in"][:, :T2].cpu().transpose(0, 1)
        assert valid_rel_err(e_in, otaps["enc_in"], fm_t2, time_first=True) < stol, f"encoder input ({dtype})"
        # the positional term in isolation (the features dominate enc_in): our enc_in minus our own features, so
        # that the conv stack's rounding cancels and what is left is this stage's error
        pos = (taps["enc_in"][:, :T].cpu() - taps["post_proj"].cpu()) * keep
        pos_err = float((pos - pos_ref).abs().max() / pos_ref.abs().max())
        assert pos_err < (1e-3 if dtype == torch.float32 else stol), f"positional conv term ({dtype}): {pos_err:.3e}"
        err = valid_rel_err(y.cpu(), yo, fmo.numpy())
        if dtype == torch.bfloat16:
            yq, _ = O.extract_features(bf16_valued(sd), cfg, wav, pm)
            # the contract: identical (bf16-valued) weights.  (Against the fp32-valued weights this 2-layer random-init
            # model sits at 2.2-2.5e-2 from the weight rounding alone -- so does the reference's own bf16 run; that
            # comparison is asserted on the released architectures, tests/test_gpu_baseline_shapes.py.)
            assert valid_rel_err(y.cpu(), yq, fmo.numpy()) < BF16_TOL, "encoder output, identical (bf16-valued) weights"
        else:
            assert err < tol, f"encoder output ({dtype})"


@pytest.mark.parametrize("name", RAIN)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_rain_forward_vs_reference_golden(name, dtype):
    g = load_golden(name)
    cfg, sd, wav, pm, lens = case_inputs(name)
    m = build(W.BlockWiseWav2Vec2Model, cfg, sd, dtype)
    kw = cases.CASES[name].get("kwargs", {})
    out = m(wav.cuda().to(dtype), None if pm is None else pm.cuda(), **kw)
    y, fm = out["encoder_out"][0], out["encoder_padding_mask"][0]
    assert tuple(y.shape) == g["y"].shape
    assert np.array_equal(fm.cpu().numpy(), g["fmask"])
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert valid_rel_err(y.cpu(), g["y"], g["fmask"], time_first=True) < tol
    assert set(out) == {"encoder_out", "encoder_padding_mask", "encoder_embedding", "encoder_states",
                        "src_tokens", "src_lengths", "dec1_state", "dec1_padding_mask"}
    if lens is not None:
        # length-based entry (OnlineW2V2TransformerEncoder.forward) gives the same bits as the mask entry
        out2 = m(wav.cuda().to(dtype), None, src_lengths=lens.cuda(), mask_len=int(lens.max()), **kw)
        assert torch.equal(out2["encoder_out"][0], y)
        assert torch.equal(out2["encoder_padding_mask"][0], fm)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("ragged", [False, True])
def test_waveform_front_end_pcm16_and_normalisation(dtype, ragged):
    """16-bit PCM in, per-utterance normalisation inside the first conv layer's load:
    against the oracle's restatement of the two host-side steps followed by the encoder."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 21)
    g = torch.Generator().manual_seed(8)
    B, L = 3, 11000
    pcm = (torch.randn(B, L, generator=g) * 3000 + 500).clamp(-32768, 32767).to(torch.int16)   # DC offset, small gain
    lens, pm = None, None
    if ragged:
        lens = synth.make_lengths(B, L, 3)
        pm = O.lengths_to_padding_mask(lens)
        pcm = pcm.masked_fill(pm, 0)
    m = build(W.Wav2VecSModel, cfg, sd, dtype)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    for normalize in (False, True):
        m.normalize_waveform = normalize
        y, fm = m.extract_features(pcm.cuda(), None if pm is None else pm.cuda())
        x = O.waveform_frontend(pcm, normalize, lens)
        yo, fmo = O.extract_features(sd, cfg, x, pm)
        if fmo is not None:
            assert torch.equal(fm.cpu(), fmo)
        assert valid_rel_err(y.cpu(), yo, None if fmo is None else fmo.numpy()) < tol, normalize
    # incremental mode takes PCM chunks as well (no normalisation there)
    m.normalize_waveform = False


def test_sample_mask_and_lengths_agree_with_arbitrary_mask():
    """A sample mask that is not a pure length mask (holes) still follows view(B,T,-1).all(-1)."""
    cfg = cases.tiny()
    sd = __import__("oracle.synth", fromlist=["x"]).make_state_dict(cfg, 3)
    m = build(W.Wav2VecSModel, cfg, sd, torch.float32)
    g = torch.Generator().manual_seed(5)
    wav = torch.randn(2, 9001, generator=g)
    pm = torch.zeros(2, 9001, dtype=torch.bool)
    pm[0, 7000:] = True
    pm[1, 5000:5400] = True       # hole in the middle
    y, fm = m.extract_features(wav.cuda(), pm.cuda())
    yo, fmo = O.extract_features(sd, cfg, wav, pm)
    assert torch.equal(fm.cpu(), fmo)
    assert valid_rel_err(y.cpu(), yo, fmo.numpy()) < FP32_TOL


@pytest.mark.parametrize("use_proj", [False, True])
def test_online_encoder_wrapper_vs_oracle(use_proj):
    """OnlineW2V2TransformerEncoder (unidirect_w2v2_encoder.py:534-607): checkpoint -> model, length-based masks,
    optional encoder_proj, the rain output dict -- what the CAAT / SimulEval pipelines call."""
    import argparse
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 31)
    args = argparse.Namespace(main_context=8, right_context=4, encoder_embed_dim=96 if use_proj else 128,
                              use_linear_layer=use_proj)
    enc = W.OnlineW2V2TransformerEncoder(args, wav2vec_ckpt={"args": None, "cfg": {"model": dict(cfg)}, "model": sd})
    enc = enc.cuda().eval()
    B, L = 3, 9000
    wav = synth.make_waveform(B, L, 32)
    lens = synth.make_lengths(B, L, 33)
    pm = O.lengths_to_padding_mask(lens)
    wav = wav.masked_fill(pm, 0.0)
    out = enc(wav.cuda(), lens.cuda())
    y, fm = out["encoder_out"][0], out["encoder_padding_mask"][0]
    ocfg = dict(cfg, main_context=8, right_context=4)
    yo, fmo = O.rain_forward(sd, ocfg, wav, pm)
    if use_proj:
        yo = torch.nn.functional.linear(yo, enc.encoder_proj.weight.detach().cpu(), enc.encoder_proj.bias.detach().cpu())
    assert torch.equal(fm.cpu(), fmo)
    assert tuple(y.shape) == tuple(yo.shape)
    assert valid_rel_err(y.detach().cpu(), yo, fmo.numpy(), time_first=True) < FP32_TOL


@pytest.mark.parametrize("name", ["tiny_preln_bias_ragged", "tiny_postln_ln1", "large_1s"])
def test_fp16_model_vs_oracle(name):
    """`.half()` models (the reference trainer halves the model under --fp16, fairseq/fairseq/trainer.py:86-90, as every
    reference training script does): accepted, fp16 in / fp16 out, arithmetic of the bf16 path (bf16 tensor-core operands,
    fp32 accumulation and residual stream -- include/w2vs.h, w2vs_config.io_dtype).  Against the reference algorithm in
    fp32 arithmetic on the identical (fp16-valued) weights and waveform, at the bf16-mode bound."""
    cfg, sd, wav, pm, _ = case_inputs(name)
    m = build(W.Wav2VecSModel, cfg, sd, torch.float16)
    src = wav.to(torch.float16)
    y, fm = m.extract_features(src.cuda(), None if pm is None else pm.cuda())
    assert y.dtype == torch.float16
    sd16 = {k: (v.to(torch.float16).float() if v.is_floating_point() else v) for k, v in sd.items()}
    yo, fmo = O.extract_features(sd16, cfg, src.float(), pm)
    if fmo is not None:
        assert torch.equal(fm.cpu(), fmo)
    assert valid_rel_err(y.cpu(), yo, None if fmo is None else fmo.numpy()) < BF16_TOL
    # the rain API and the incremental path return fp16 as well
    r = build(W.BlockWiseWav2Vec2Model, cfg, sd, torch.float16)
    out = r(src.cuda(), None if pm is None else pm.cuda())
    assert out["encoder_out"][0].dtype == torch.float16
    assert valid_rel_err(out["encoder_out"][0].transpose(0, 1).cpu(), yo, None if fmo is None else fmo.numpy()) < BF16_TOL
    if cfg["extractor_mode"] == "layer_norm" and pm is None:
        st = r.open_stream(B=src.size(0), max_seconds=2.0, max_new_samples=src.size(1))
        ys = st.step(src.cuda(), 1)
        assert ys.dtype == torch.float16 and valid_rel_err(ys.transpose(0, 1).cpu(), yo) < BF16_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("era", ["argparse", "hydra"])
def test_from_checkpoint_forward_vs_oracle(era, dtype):
    """Checkpoint ingestion end to end (rain/layers/unidirect_w2v2_encoder.py:541-555): a
    composite CAAT-style checkpoint -- encoder under `encoder.w2v2_model.`, foreign decoder / joiner keys, pre-training
    heads -- in both config eras (argparse-era: `args` Namespace without extractor_mode / pos_type, which the reference
    then forces to layer_norm / sin; hydra-era: cfg.model dict), context overridden at load time, strict=False; the
    forward of the model it builds against the oracle."""
    import argparse
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 71)
    full = {"encoder.w2v2_model." + k: v for k, v in sd.items()}
    full["encoder.w2v2_model.mask_emb"] = torch.zeros(cfg["encoder_embed_dim"])
    full["encoder.w2v2_model.quantizer.vars"] = torch.zeros(1, 8, 4)
    full["decoder.embed_tokens.weight"] = torch.zeros(10, 16)
    full["joiner.proj.weight"] = torch.zeros(4, 4)
    if era == "argparse":
        fields = {k: v for k, v in cfg.items() if k not in ("extractor_mode", "pos_type")}
        ckpt = {"args": argparse.Namespace(**fields), "cfg": None, "model": full}
    else:
        ckpt = {"args": None, "cfg": {"model": dict(cfg)}, "model": full}
    m, missing, unexpected = W.BlockWiseWav2Vec2Model.from_checkpoint(ckpt, main_context=8, right_context=4)
    assert set(missing) <= {"mask_emb"} and not any(k.startswith(("decoder.", "joiner.")) for k in unexpected)
    assert (m.encoder.main_context, m.encoder.right_context) == (8, 4)
    assert m.args.extractor_mode == "layer_norm" and m.args.pos_type == "sin"
    m = m.to("cuda", dtype).eval()
    B, L = 2, 9000
    wav = synth.make_waveform(B, L, 72)
    pm = O.lengths_to_padding_mask(synth.make_lengths(B, L, 73))
    wav = wav.masked_fill(pm, 0.0)
    y = m(wav.cuda(), pm.cuda())["encoder_out"][0]
    ocfg = dict(cfg, main_context=8, right_context=4)
    yo, fmo = O.rain_forward(sd if dtype == torch.float32 else bf16_valued(sd), ocfg, wav, pm)
    assert valid_rel_err(y.cpu(), yo, fmo.numpy(), time_first=True) < (FP32_TOL if dtype == torch.float32 else BF16_TOL)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_online_encoder_proj_on_library_gemm(dtype):
    """OnlineW2V2TransformerEncoder.encoder_proj (unidirect_w2v2_encoder.py:559-562,590-594) in both modes: fp32 on the
    CUDA-core product, bf16 on the tcgen05 kernel (M = T x B rows is a few hundred: the 2-CTA kernel with narrow tiles)."""
    import argparse
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 31)
    args = argparse.Namespace(main_context=16, right_context=8, encoder_embed_dim=96, use_linear_layer=True)
    enc = W.OnlineW2V2TransformerEncoder(args, wav2vec_ckpt={"args": None, "cfg": {"model": dict(cfg)}, "model": sd})
    enc = enc.to("cuda", dtype).eval()
    B, L = 4, 40000
    wav = synth.make_waveform(B, L, 32)
    lens = synth.make_lengths(B, L, 33)
    pm = O.lengths_to_padding_mask(lens)
    wav = wav.masked_fill(pm, 0.0)
    out = enc(wav.cuda().to(dtype), lens.cuda())
    y, fm = out["encoder_out"][0], out["encoder_padding_mask"][0]
    assert y.dtype == dtype and y.size(-1) == 96
    wsd = sd if dtype == torch.float32 else bf16_valued(sd)
    yo, fmo = O.rain_forward(wsd, cfg, wav.to(dtype).float(), pm)
    pw, pb = enc.encoder_proj.weight.detach().float().cpu(), enc.encoder_proj.bias.detach().float().cpu()
    yo = torch.nn.functional.linear(yo, pw, pb)
    assert torch.equal(fm.cpu(), fmo)
    assert valid_rel_err(y.detach().cpu(), yo, fmo.numpy(), time_first=True) < (FP32_TOL if dtype == torch.float32 else BF16_TOL)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices in one process")
def test_two_devices_in_one_process():
    """One process, a model replica on each of two devices (the library keeps its per-device caches -- SM count,
    shared-memory opt-ins -- keyed by the current device): same bits on both."""
    from oracle import synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 81)
    wav = synth.make_waveform(2, 12000, 82)
    ys = []
    for dev in ("cuda:0", "cuda:1"):
        m = W.Wav2VecSModel(cfg)
        m.load_state_dict(sd, strict=False)
        m = m.to(dev, torch.bfloat16).eval()
        with torch.cuda.device(dev):
            y, _ = m.extract_features(wav.to(dev), None)
            torch.cuda.synchronize()
        ys.append(y.cpu())
    assert torch.equal(ys[0], ys[1])
