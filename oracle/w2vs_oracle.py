"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the wav2vec-S encoder forward.

This is the oracle the CUDA path is checked against.  It is a plain PyTorch (CPU, fp32 by
default) functional restatement of the reference algorithm; every function cites the
reference file:line it follows (paths relative to /root/reference).  It is NOT part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.

Pinning: the reference's tests hold no golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build container
by ``tests/golden/make_golden.py`` through ``oracle/ref_shim.py`` (which executes the
unmodified reference sources) and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks the oracle against them on every run, and
``tests/test_oracle_vs_reference.py`` re-checks against the live reference when present.

The numerical kernels (conv1d, layer_norm, group_norm, GELU, linear, softmax attention) live in
PyTorch, an unpinned third-party dependency of the reference (fairseq/setup.py:210 "torch");
the oracle calls the same torch functions the reference call sites dispatch to.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

DEFAULT_CONV_LAYERS = "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] + [(512,2,2)]"


def default_cfg(**over) -> dict:
    """Reference defaults that shape the path (wav2vec_S.py:43-311; rain base_architecture
    unidirect_w2v2_encoder.py:679-750)."""
    cfg = dict(
        extractor_mode="default", encoder_layers=12, encoder_embed_dim=768,
        encoder_ffn_embed_dim=3072, encoder_attention_heads=12, activation_fn="gelu",
        layer_norm_first=False, conv_feature_layers=DEFAULT_CONV_LAYERS, conv_bias=False,
        conv_pos=128, conv_pos_groups=16, pos_type="sin", context_type="constant",
        main_context=16, right_context=8, required_seq_len_multiple=2,
    )
    cfg.update(over)
    return cfg


def base_cfg(**over) -> dict:
    """Released wav2vec-S base (wav2vec-S_base_librispeech.yaml:50-76)."""
    return default_cfg(extractor_mode="layer_norm", **over)


def large_cfg(**over) -> dict:
    """Released wav2vec-S large (wav2vec-S_large_librivox.yaml:52-90)."""
    return default_cfg(extractor_mode="layer_norm", encoder_layers=24, encoder_embed_dim=1024,
                       encoder_ffn_embed_dim=4096, encoder_attention_heads=16,
                       layer_norm_first=True, conv_bias=True, **over)


def conv_layers_of(cfg) -> List[Tuple[int, int, int]]:
    cl = cfg["conv_feature_layers"]
    return list(eval(cl)) if isinstance(cl, str) else list(cl)


def layer_norm_num(cfg) -> int:
    # wav2vec2.py:317 / wav2vec_S.py:325
    return 1 if cfg["encoder_layers"] == 12 else 7


# --------------------------------------------------------------------------------------
# integer / bool structure (must be bit-exact)
# --------------------------------------------------------------------------------------
def conv_out_lengths(cfg, L: int) -> List[int]:
    """Per-layer output lengths of the un-padded Conv1d stack (wav2vec2.py:725)."""
    out = []
    t = L
    for (_, k, s) in conv_layers_of(cfg):
        t = (t - k) // s + 1
        out.append(t)
    return out


def frame_padding_mask(padding_mask: Optional[Tensor], T: int) -> Optional[Tensor]:
    """Down-sample a sample-level bool mask [B,L] to frames [B,T]
    (wav2vec2.py:560-565; rain unidirect_w2v2_encoder.py:500-505)."""
    if padding_mask is None:
        return None
    extra = padding_mask.size(1) % T
    if extra > 0:
        padding_mask = padding_mask[:, :-extra]
    return padding_mask.view(padding_mask.size(0), T, -1).all(-1)


def lengths_to_padding_mask(lens: Tensor) -> Tensor:
    """fairseq/data/data_utils.py:528-532."""
    max_len = int(lens.max())
    return torch.arange(max_len).view(1, -1) >= lens.view(-1, 1)


def make_positions(is_pad: Tensor, padding_idx: int = 1) -> Tensor:
    """Positions fed to the sinusoidal table: the bool padding mask is used as "tokens" with
    padding_idx=1 (utils.py:250-260, called from sinusoidal_positional_embedding.py:93)."""
    nonpad = (~is_pad).int()
    return (torch.cumsum(nonpad, dim=1) * nonpad).long() + padding_idx


def sinusoidal_table(num: int, dim: int, padding_idx: int = 1) -> Tensor:
    """sinusoidal_positional_embedding.py:36-59 (fp32; sin half then cos half)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(num, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num, -1)
    if dim % 2 == 1:
        e = torch.cat([e, torch.zeros(num, 1)], dim=1)
    e[padding_idx, :] = 0
    return e


def block_mask_structure(T2: int, main: int, rc: int):
    """Index structure of gen_block_attn_mask (wav2vec_S.py:444-489 == rain :68-115).

    Returns (rc_idx[R] clamped gather indices, rc_oor[R] out-of-range flags,
             masked[M,M] bool, True = additive -1e4)."""
    block_idx = torch.arange(T2) // main
    nb = T2 // main
    if rc == 0:
        masked = block_idx.unsqueeze(1) < block_idx.unsqueeze(0)
        return torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.bool), masked
    owner = torch.arange(nb).repeat_interleave(rc)
    rc_idx = ((torch.arange(nb).unsqueeze(1) + 1) * main + torch.arange(rc).unsqueeze(0)).view(-1)
    rc_oor = rc_idx > (T2 - 1)
    rc_idx = rc_idx.clamp(0, T2 - 1)
    full = torch.cat([block_idx, owner])
    m1 = full.unsqueeze(1) < block_idx.unsqueeze(0)
    m2 = full.unsqueeze(1).ne(owner.unsqueeze(0))
    return rc_idx, rc_oor, torch.cat([m1, m2], dim=1)


# --------------------------------------------------------------------------------------
# float stages
# --------------------------------------------------------------------------------------
def conv_feature_extractor(sd: Dict[str, Tensor], cfg, source: Tensor, taps=None) -> Tensor:
    """ConvFeatureExtractionModel.forward (wav2vec2.py:773-781, blocks :715-752).
    [B,L] -> [B,C,T] channels-first."""
    x = source.unsqueeze(1)
    mode = cfg["extractor_mode"]
    n_ln = layer_norm_num(cfg)
    for i, (dim, k, s) in enumerate(conv_layers_of(cfg)):
        p = f"feature_extractor.conv_layers.{i}."
        x = F.conv1d(x, sd[p + "0.weight"], sd.get(p + "0.bias"), stride=s)
        if mode == "layer_norm" and i < n_ln:
            # TransposeLast -> Fp32LayerNorm -> TransposeLast (modules/layer_norm.py:39-50)
            y = F.layer_norm(x.transpose(-2, -1).float(), (dim,),
                             sd[p + "2.1.weight"].float(), sd[p + "2.1.bias"].float(), 1e-5)
            x = y.type_as(x).transpose(-2, -1)
        elif mode == "default" and i == 0:
            # Fp32GroupNorm(dim, dim) (modules/fp32_group_norm.py:13-25)
            y = F.group_norm(x.float(), dim, sd[p + "2.weight"].float(),
                             sd[p + "2.bias"].float(), 1e-5)
            x = y.type_as(x)
        x = F.gelu(x)  # nn.GELU (erf)
        if taps is not None:
            taps[f"conv{i}"] = x
    return x


def pos_conv_embed(sd, cfg, x: Tensor) -> Tensor:
    """Convolutional positional embedding (wav2vec2.py:791-804; SamePad same_pad.py:10-21).
    x: [B,T,D] -> [B,T,D]."""
    k, g = cfg["conv_pos"], cfg["conv_pos_groups"]
    wg, wv = sd["encoder.pos_conv.0.weight_g"], sd["encoder.pos_conv.0.weight_v"]
    # torch.nn.utils.weight_norm(dim=2): norm over all dims except dim 2
    w = wv * (wg / wv.norm(2, dim=(0, 1), keepdim=True))
    y = F.conv1d(x.transpose(1, 2), w, sd["encoder.pos_conv.0.bias"], padding=k // 2, groups=g)
    if k % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


def mha(sd, prefix: str, x: Tensor, heads: int, key_padding_mask: Tensor,
        attn_mask: Optional[Tensor]) -> Tensor:
    """MultiheadAttention fast path (modules/multihead_attention.py:162-194): separate q/k/v
    projection weights, concatenated bias, torch's multi_head_attention_forward."""
    D = x.size(-1)
    out, _ = F.multi_head_attention_forward(
        x, x, x, D, heads, torch.empty([0]),
        torch.cat((sd[prefix + "q_proj.bias"], sd[prefix + "k_proj.bias"],
                   sd[prefix + "v_proj.bias"])),
        None, None, False, 0.0,
        sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"],
        False, key_padding_mask, False, attn_mask,
        use_separate_proj_weight=True,
        q_proj_weight=sd[prefix + "q_proj.weight"],
        k_proj_weight=sd[prefix + "k_proj.weight"],
        v_proj_weight=sd[prefix + "v_proj.weight"])
    return out


def _gelu(x):
    # fairseq/modules/gelu.py:24-25
    return F.gelu(x.float()).type_as(x)


def encoder_layer(sd, cfg, n: int, x: Tensor, pad: Tensor, attn_mask: Tensor) -> Tensor:
    """TransformerSentenceEncoderLayer.forward (wav2vec2.py:921-978)."""
    p = f"encoder.layers.{n}."
    D = x.size(-1)
    H = cfg["encoder_attention_heads"]

    def ln(v, name):
        return F.layer_norm(v, (D,), sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-5)

    def ffn(v):
        v = _gelu(F.linear(v, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
        return F.linear(v, sd[p + "fc2.weight"], sd[p + "fc2.bias"])

    if cfg["layer_norm_first"]:
        x = x + mha(sd, p + "self_attn.", ln(x, "self_attn_layer_norm"), H, pad, attn_mask)
        x = x + ffn(ln(x, "final_layer_norm"))
    else:
        x = ln(x + mha(sd, p + "self_attn.", x, H, pad, attn_mask), "self_attn_layer_norm")
        x = ln(x + ffn(x), "final_layer_norm")
    return x


def _features(sd, cfg, source, padding_mask, taps):
    """Wav2Vec2Model.forward up to the encoder input (wav2vec2.py:544-571)."""
    feats = conv_feature_extractor(sd, cfg, source, taps).transpose(1, 2)  # [B,T,C]
    C = feats.size(-1)
    feats = F.layer_norm(feats, (C,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
    fmask = frame_padding_mask(padding_mask, feats.size(1))
    if "post_extract_proj.weight" in sd:
        feats = F.linear(feats, sd["post_extract_proj.weight"], sd["post_extract_proj.bias"])
    if taps is not None:
        taps["post_proj"] = feats
    return feats, fmask


def _blockwise_encoder(sd, cfg, x, fmask, main, rc, taps):
    """BlockwiseTransformerEncoder.extract_features (wav2vec_S.py:355-440; rain :262-330).
    x [B,T,D] -> (x [T2-trimmed... see callers], layers run on M = T2 + R tokens).
    Returns (x[T,B,D] main tokens incl. seq-padding removed, padding mask [B,T] after
    pad_to_multiple/trim, which may be non-None even if fmask was None)."""
    B, T, D = x.shape
    if fmask is not None:
        x = x.clone()
        x[fmask] = 0
        is_pad = fmask
    else:
        is_pad = torch.zeros(B, T, dtype=torch.bool)
    if cfg["pos_type"] == "conv":
        x = x + pos_conv_embed(sd, cfg, x)
    else:
        pos = make_positions(is_pad)
        table = sinusoidal_table(max(int(pos.max()) + 1, 3), D).to(x.dtype)
        x = x + table.index_select(0, pos.view(-1)).view(B, T, D)
    if not cfg["layer_norm_first"]:
        x = F.layer_norm(x, (D,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5)
    mult = cfg["required_seq_len_multiple"]
    pad_len = (-T) % mult
    pm = fmask
    if pad_len > 0:
        x = F.pad(x, (0, 0, 0, pad_len), value=0)
        if pm is None:
            pm = torch.zeros(B, T + pad_len, dtype=torch.bool)
            pm[:, -pad_len:] = True
        else:
            pm = F.pad(pm, (0, pad_len), value=True)
    T2 = T + pad_len
    x = x.transpose(0, 1)  # [T2,B,D]
    if taps is not None:
        taps["enc_in"] = x
    if pm is None:
        pm = torch.zeros(B, T2, dtype=torch.bool)
    rc_idx, rc_oor, masked = block_mask_structure(T2, main, rc)
    if rc > 0:
        pm_ext = torch.cat([pm, pm.index_select(1, rc_idx) | rc_oor.unsqueeze(0)], dim=1)
        x = torch.cat([x, x.index_select(0, rc_idx)], dim=0)
    else:
        pm_ext = pm
    attn_mask = torch.zeros(masked.shape, dtype=x.dtype).masked_fill(masked, -1e4)
    for n in range(cfg["encoder_layers"]):
        x = encoder_layer(sd, cfg, n, x, pm_ext, attn_mask)
        if taps is not None:
            taps[f"layer{n}"] = x[:T2]
    x = x[:T2]
    pm_out = pm_ext[:, :T2]
    if pad_len > 0:
        x = x[:-pad_len]
        pm_out = pm_out[:, :-pad_len]
    return x, pm_out


def _ctx(cfg, main_context, right_context):
    main = cfg["main_context"] if main_context is None else main_context
    rc = cfg["right_context"] if right_context is None else right_context
    return main, rc


def _final_ln(sd, cfg, x):
    if cfg["layer_norm_first"]:
        D = x.size(-1)
        x = F.layer_norm(x, (D,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5)
    return x


@torch.no_grad()
def extract_features(sd, cfg, source: Tensor, padding_mask: Optional[Tensor] = None,
                     main_context=None, right_context=None, taps=None):
    """fairseq API: Wav2VecSModel.extract_features(source, padding_mask, mask=False)
    (wav2vec2.py:667-669 -> :544-603 -> TransformerEncoder.forward :828-834).
    Returns (x [B,T,D], frame padding mask [B,T] or None)."""
    main, rc = _ctx(cfg, main_context, right_context)
    feats, fmask = _features(sd, cfg, source, padding_mask, taps)
    x, _ = _blockwise_encoder(sd, cfg, feats, fmask, main, rc, taps)
    x = _final_ln(sd, cfg, x.transpose(0, 1))
    return x, fmask


@torch.no_grad()
def rain_forward(sd, cfg, source: Tensor, padding_mask: Optional[Tensor] = None,
                 finished: bool = False, is_infer: bool = False,
                 main_context=None, right_context=None, taps=None):
    """rain API: BlockWiseWav2Vec2Model.forward (unidirect_w2v2_encoder.py:485-531) with
    BlockwiseW2V2TransformerEncoder.forward/extract_features (:254-330).
    Returns (x [T,B,D], encoder_padding_mask [B,T])."""
    main, rc = _ctx(cfg, main_context, right_context)
    feats, fmask = _features(sd, cfg, source, padding_mask, taps)
    x, pm = _blockwise_encoder(sd, cfg, feats, fmask, main, rc, taps)
    if is_infer and not finished and rc > 0:
        x = x[:-rc]
        pm = pm[:, :-rc]
    return _final_ln(sd, cfg, x), pm


@torch.no_grad()
def streaming_prefix_recompute(sd, cfg, wav: Tensor, step_blocks: int = 1):
    """What the reference's SimulEval driver computes (rain/simul/transducer_searcher.py:702-760
    + rain/simul/transducer_agent.py:138-167): at every decision step the encoder is re-run on
    the WHOLE prefix with is_infer=True, and only frames not emitted before are consumed.
    wav: [1,L].  Returns list of (n_samples_in_prefix, new_frames [t_new,1,D])."""
    main, rc = cfg["main_context"], cfg["right_context"]
    L = wav.size(1)
    hop = 1
    for (_, _, s) in conv_layers_of(cfg):
        hop *= s
    out, emitted, k = [], 0, 0
    while True:
        need_frames = main + rc + k * main * step_blocks
        # smallest prefix with exactly need_frames frames
        n = _samples_for_frames(cfg, need_frames)
        fin = n >= L
        n = min(n, L)
        x, _ = rain_forward(sd, cfg, wav[:, :n], None, finished=fin, is_infer=True)
        out.append((n, x[emitted:]))
        emitted = x.size(0)
        if fin:
            return out
        k += 1


def _samples_for_frames(cfg, frames: int) -> int:
    n = frames
    for (_, k, s) in reversed(conv_layers_of(cfg)):
        n = (n - 1) * s + k
    return n


def waveform_frontend(samples: Tensor, normalize: bool, lengths: Optional[Tensor] = None) -> Tensor:
    """The two host-side steps in front of the encoder: 16-bit PCM -> float as the
    SimulEval agent does (rain/simul/transducer_searcher.py:74-80: ``output / 32768.0`` then float32), and the data
    pipeline's per-utterance normalisation (fairseq/data/audio/raw_audio_dataset.py:60-72:
    ``F.layer_norm(feats, feats.shape)`` on each utterance BEFORE the collater pads it with zeros).
    samples [B, L] int16 or float; lengths [B] valid samples or None.  Returns float32 [B, L]."""
    x = (samples.double() / 32768.0).float() if samples.dtype == torch.int16 else samples.float()
    if not normalize:
        return x
    out = x.clone()
    for b in range(x.size(0)):
        n = x.size(1) if lengths is None else int(lengths[b])
        out[b, :n] = F.layer_norm(x[b, :n], (n,))
    return out


def max_abs_rel(y: Tensor, ref: Tensor) -> float:
    """Parity metric of SURVEY.md section 8(d): max|y - ref| / max|ref|."""
    return float((y.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30))
