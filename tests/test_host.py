"""CPU-side checks: the C-ABI library loads and exports every symbol include/w2vs.h declares, the
integer geometry is bit-exact against the oracle, and the host classes keep the reference's
state_dict layout and error behaviour.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import wav2vec_s_b200 as W
from wav2vec_s_b200 import cabi
from oracle import cases, synth
from oracle import w2vs_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "w2vs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(w2vs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 15
    handle = C.CDLL(cabi.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/w2vs.h but not exported"
    assert set(syms) == set(cabi.PROTOTYPES), "ctypes binding and header disagree"
    assert cabi.lib().w2vs_status_string(0) == b"ok"


def test_binding_constants_match_the_header():
    """Every #define and enumerator of include/w2vs.h the ctypes binding mirrors has the header's value."""
    src = open(os.path.join(ROOT, "include", "w2vs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    defines = {k: int(v) for k, v in re.findall(r"#define\s+(W2VS_[A-Z0-9_]+)\s+(\d+)\b", src)}
    enums = {}
    for body in re.findall(r"typedef\s+enum\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            name, _, val = item.partition("=")
            nxt = int(val) if val.strip() else nxt
            enums[name.strip()] = nxt
            nxt += 1
    want = {
        "W2VS_MAX_CONV": cabi.W2VS_MAX_CONV, "W2VS_ABI_VERSION": cabi.W2VS_ABI_VERSION,
        "W2VS_EPI_GELU": cabi.EPI_GELU, "W2VS_EPI_SPLITK": cabi.EPI_SPLITK,
    }
    for k, v in want.items():
        assert defines[k] == v, k
    want_enum = {
        "W2VS_OK": cabi.OK, "W2VS_INVALID_VALUE": cabi.INVALID_VALUE, "W2VS_UNSUPPORTED": cabi.UNSUPPORTED,
        "W2VS_WORKSPACE_TOO_SMALL": cabi.WORKSPACE_TOO_SMALL, "W2VS_CUDA_ERROR": cabi.CUDA_ERROR,
        "W2VS_GEMM_AUTO": cabi.GEMM_AUTO, "W2VS_GEMM_SIMT": cabi.GEMM_SIMT,
        "W2VS_GEMM_TCGEN05_2CTA": cabi.GEMM_TCGEN05_2CTA, "W2VS_GEMM_SKINNY": cabi.GEMM_SKINNY,
    }
    for k, v in want_enum.items():
        assert enums[k] == v, k


@pytest.mark.parametrize("cfg", [O.base_cfg(), O.large_cfg(), cases.tiny(), cases.tiny(main_context=8, right_context=4)])
def test_geometry_bit_exact(cfg):
    m = W.Wav2VecSModel(cfg)
    main, rc = cfg["main_context"], cfg["right_context"]
    rs = np.random.RandomState(0)
    Ls = [400, 401, 719, 720, 5760, 8000, 16000, 160000, 240000, 320000, 480000] + list(rs.randint(400, 200000, 40))
    for L in Ls:
        g = m.geometry(int(L), main, rc)
        lens = O.conv_out_lengths(cfg, int(L))
        assert list(g.conv_len[:len(lens)]) == lens
        T = lens[-1]
        T2 = T + (-T) % cfg["required_seq_len_multiple"]
        rc_idx, _, masked = O.block_mask_structure(T2, main, rc)
        assert (g.frames, g.frames_pad, g.n_blocks, g.tokens) == (T, T2, T2 // main, masked.shape[0])
        # allocated rows: each layer's rows = stride * next layer's rows, and cover the valid lengths
        spec = O.conv_layers_of(cfg)
        for i in range(len(spec)):
            assert g.conv_rows[i] >= g.conv_len[i]
            if i > 0:
                assert g.conv_rows[i - 1] == spec[i][2] * g.conv_rows[i]


def test_geometry_rejects_short_input():
    m = W.Wav2VecSModel(O.base_cfg())
    with pytest.raises(cabi.W2vsError) as e:
        m.geometry(399, 16, 8)
    assert e.value.status == cabi.INVALID_VALUE


@pytest.mark.parametrize("name", ["tiny_postln_ln1", "tiny_groupnorm_posconv", "large_1s", "base_1s"])
def test_state_dict_layout_matches_reference(name):
    cfg = cases.CASES[name]["cfg"]
    sd = synth.make_state_dict(cfg, 0)           # reference key layout
    m = W.Wav2VecSModel(cfg)
    ours = m.state_dict()
    assert set(ours) - {"mask_emb"} == set(sd)
    for k, v in sd.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    # a released checkpoint also carries pre-training heads: accepted and ignored, even when strict
    extra = dict(sd)
    extra["mask_emb"] = torch.zeros(cfg["encoder_embed_dim"])
    extra["quantizer.vars"] = torch.zeros(1, 640, 384)
    extra["project_q.weight"] = torch.zeros(4, 4)
    extra["final_proj.bias"] = torch.zeros(4)
    m.load_state_dict(extra, strict=True)
    assert torch.equal(m.state_dict()["layer_norm.weight"], sd["layer_norm.weight"])


def test_rain_model_and_online_encoder_construct():
    import argparse
    cfg = cases.tiny(layer_norm_first=True)
    sd = synth.make_state_dict(cfg, 0)
    ckpt = {"args": None, "cfg": {"model": dict(cfg)}, "model": sd}
    enc = W.OnlineW2V2TransformerEncoder(argparse.Namespace(main_context=8, right_context=4,
                                                            encoder_embed_dim=128, use_linear_layer=False),
                                         wav2vec_ckpt=ckpt)
    assert enc.init_frames == 12 and enc.step_frames == 8
    assert enc.w2v2_model.encoder.main_context == 8 and enc.w2v2_model.encoder.right_context == 4
    assert len(enc.w2v2_model.encoder.layers) == 3


def test_no_cpu_fallback():
    m = W.Wav2VecSModel(cases.tiny())
    with pytest.raises(RuntimeError, match="CUDA"):
        m.extract_features(torch.zeros(1, 4000), None)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 4000), None, mask=True, features_only=True)


def test_bad_context_type_raises_like_reference():
    m = W.Wav2VecSModel(cases.tiny(context_type="bogus"))
    with pytest.raises(ValueError, match="context_type"):
        m.encoder.pick_context()


def test_sinusoidal_table_matches_oracle():
    for D in (128, 768, 1024):
        assert torch.equal(W.sinusoidal_table(300, D), O.sinusoidal_table(300, D))


def test_invalid_config_is_rejected_before_any_launch():
    c = cabi.Config()
    size = C.c_size_t()
    assert cabi.lib().w2vs_packed_weights_size(C.byref(c), C.byref(size)) == cabi.INVALID_VALUE
    assert b"abi_version" in cabi.lib().w2vs_last_error()
    m = W.Wav2VecSModel(cases.tiny(encoder_attention_heads=4))   # head_dim 32
    with pytest.raises(cabi.W2vsError) as e:
        m.geometry(8000, 16, 8)
    assert e.value.status == cabi.UNSUPPORTED


def test_posconv_layout_sizes():
    """pos_type="conv": the bf16 model carries the extra tensor-core operand of the positional conv
    ([groups][D/groups][k * 64] bf16) and its workspace the group-major frame copy; integer layout arithmetic only."""
    cfg = O.default_cfg(extractor_mode="layer_norm", pos_type="conv")            # D = 768, 16 groups of 48, k = 128
    m = W.Wav2VecSModel(cfg)
    D, groups, k, Dgp = 768, 16, 128, 64
    def sizes(dtype):
        c = m._c_config(dtype, 0)
        pw, ws = C.c_size_t(), C.c_size_t()
        cabi.check(cabi.lib().w2vs_packed_weights_size(C.byref(c), C.byref(pw)), "packed")
        cabi.check(cabi.lib().w2vs_get_workspace_size(C.byref(c), 2, 16000, 16, 8, C.byref(ws)), "workspace")
        return pw.value, ws.value
    m_sin = W.Wav2VecSModel(dict(cfg, pos_type="sin"))
    def sizes_sin(dtype):
        c = m_sin._c_config(dtype, 64)
        pw = C.c_size_t()
        cabi.check(cabi.lib().w2vs_packed_weights_size(C.byref(c), C.byref(pw)), "packed")
        return pw.value
    folded = D * (D // groups) * k * 4 + D * 4                                    # fp32 folded weights + bias
    sin_tab = 64 * D * 4
    assert sizes(torch.float32)[0] - (sizes_sin(torch.float32) - sin_tab) == folded
    # (the sinusoidal bf16 model also carries the slab-ordered fc2 copy of the fused incremental step, layout.h)
    w2s = 12 * D * 3072 * 2
    assert sizes(torch.bfloat16)[0] - (sizes_sin(torch.bfloat16) - sin_tab - w2s) == folded + D * k * Dgp * 2
    T = m.geometry(16000, 16, 8).frames
    assert sizes(torch.bfloat16)[1] > 2 * (T + k) * D * 4 + groups * (2 * (T + k) + k) * Dgp * 2


def test_from_checkpoint_formats():
    """Checkpoint ingestion: hydra-era {"cfg": {"model": ..}}, argparse-era {"args": ..},
    and a composite fine-tuned checkpoint with the encoder under a prefix and foreign keys next to it."""
    import argparse
    import torch
    import wav2vec_s_b200 as W
    from oracle import cases, synth
    cfg = cases.tiny(layer_norm_first=True, conv_bias=True)
    sd = synth.make_state_dict(cfg, 3)
    m1, missing, unexpected = W.Wav2VecSModel.from_checkpoint({"cfg": {"model": dict(cfg)}, "model": sd})
    assert not unexpected and set(missing) <= {"mask_emb"}
    m2, missing, unexpected = W.BlockWiseWav2Vec2Model.from_checkpoint(
        {"args": argparse.Namespace(**cfg), "cfg": None, "model": sd}, main_context=8, right_context=4)
    assert not unexpected and (m2.encoder.main_context, m2.encoder.right_context) == (8, 4)
    composite = {"encoder.w2v2_model." + k: v for k, v in sd.items()}
    composite["decoder.embed_tokens.weight"] = torch.zeros(4, 4)
    composite["encoder.encoder_proj.weight"] = torch.zeros(4, 4)
    m3, missing, unexpected = W.Wav2VecSModel.from_checkpoint({"cfg": {"model": {"w2v_args": {"model": dict(cfg)}}},
                                                               "model": composite})
    assert not unexpected and set(missing) <= {"mask_emb"}
    for k, v in m1.state_dict().items():
        if k in sd:
            assert torch.equal(v, sd[k]) and torch.equal(m3.state_dict()[k], sd[k]), k
