"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights and waveforms.

Weights are drawn from ``numpy.random.RandomState`` (bit-stable across numpy versions and
machines) with the reference's initialisation *scales* (kaiming-normal convs wav2vec2.py:726,
N(0,0.02) linears transformer_sentence_encoder.py:21-53), plus non-trivial biases / norm affine
parameters so that every term of the forward is exercised.  Because the draw is a pure function
of (cfg, seed), golden fixtures only need to store outputs.
"""
import math
from typing import Dict

import numpy as np
import torch

from .w2vs_oracle import conv_layers_of, layer_norm_num


def make_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict with exactly the reference key layout."""
    rs = np.random.RandomState(seed)
    sd = {}

    def normal(shape, std):
        return torch.from_numpy((rs.standard_normal(size=shape) * std).astype(np.float32))

    def affine(prefix, dim):
        sd[prefix + ".weight"] = 1.0 + normal((dim,), 0.1)
        sd[prefix + ".bias"] = normal((dim,), 0.1)

    c_in = 1
    mode = cfg["extractor_mode"]
    for i, (dim, k, s) in enumerate(conv_layers_of(cfg)):
        p = f"feature_extractor.conv_layers.{i}."
        sd[p + "0.weight"] = normal((dim, c_in, k), math.sqrt(2.0 / (c_in * k)))
        if cfg["conv_bias"]:
            sd[p + "0.bias"] = normal((dim,), 0.05)
        if mode == "layer_norm" and i < layer_norm_num(cfg):
            affine(p + "2.1", dim)
        elif mode == "default" and i == 0:
            affine(p + "2", dim)
        c_in = dim
    D, Fd = cfg["encoder_embed_dim"], cfg["encoder_ffn_embed_dim"]
    affine("layer_norm", c_in)
    if c_in != D:
        sd["post_extract_proj.weight"] = normal((D, c_in), 0.05)
        sd["post_extract_proj.bias"] = normal((D,), 0.05)
    if cfg["pos_type"] == "conv":
        k, g = cfg["conv_pos"], cfg["conv_pos_groups"]
        sd["encoder.pos_conv.0.bias"] = normal((D,), 0.05)
        v = normal((D, D // g, k), math.sqrt(4.0 / (k * D)))
        sd["encoder.pos_conv.0.weight_v"] = v
        sd["encoder.pos_conv.0.weight_g"] = (v.norm(2, dim=(0, 1), keepdim=True)
                                             * (1.0 + normal((1, 1, k), 0.1)))
    else:
        sd["encoder.pos_conv._float_tensor"] = torch.zeros(1)
    for n in range(cfg["encoder_layers"]):
        p = f"encoder.layers.{n}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            # larger than the 0.02 BERT init so that attention is not near-uniform
            sd[p + f"self_attn.{nm}.weight"] = normal((D, D), 0.06)
            sd[p + f"self_attn.{nm}.bias"] = normal((D,), 0.05)
        affine(p + "self_attn_layer_norm", D)
        sd[p + "fc1.weight"] = normal((Fd, D), 0.04)
        sd[p + "fc1.bias"] = normal((Fd,), 0.05)
        sd[p + "fc2.weight"] = normal((D, Fd), 0.04)
        sd[p + "fc2.bias"] = normal((D,), 0.05)
        affine(p + "final_layer_norm", D)
    affine("encoder.layer_norm", D)
    return sd


def make_waveform(B: int, L: int, seed: int = 1234, normalize: bool = True) -> torch.Tensor:
    """Synthetic waveform of SURVEY.md section 8(d): randn, per-utterance zero-mean/unit-var
    (matches ``normalize: true``, fairseq/data/audio/raw_audio_dataset.py:69-72)."""
    rs = np.random.RandomState(seed)
    w = torch.from_numpy(rs.standard_normal(size=(B, L)).astype(np.float32))
    if normalize:
        w = (w - w.mean(dim=1, keepdim=True)) / torch.sqrt(w.var(dim=1, unbiased=False, keepdim=True) + 1e-5)
    return w


def make_lengths(B: int, L: int, seed: int = 4321, lo: float = 0.5) -> torch.Tensor:
    """Ragged lengths len_b ~ U[lo*L, L]; the longest is forced to L."""
    rs = np.random.RandomState(seed)
    lens = rs.randint(int(lo * L), L + 1, size=(B,)).astype(np.int64)
    lens[rs.randint(0, B)] = L
    return torch.from_numpy(lens)
