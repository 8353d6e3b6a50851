"""Host-side mirror of the reference model API for the wav2vec-S encoder forward path.

Same class names, constructor arguments, ``state_dict`` key layout, method names, argument meaning
and return values as the reference (paths relative to the reference repository root):

  * ``Wav2VecSModel``            fairseq/fairseq/models/wav2vec/wav2vec_S.py:314-332
        ``extract_features(source, padding_mask, mask=False) -> (x [B,T,D], padding_mask [B,T] | None)``
        (wav2vec2.py:667-669, :544-603)
  * ``BlockWiseWav2Vec2Model``   rain/layers/unidirect_w2v2_encoder.py:443-531
        ``forward(source, padding_mask, incremental_state, finished, is_infer) -> dict``
  * ``OnlineW2V2TransformerEncoder``  rain/layers/unidirect_w2v2_encoder.py:534-607

The modules below are *parameter holders* only: they give the checkpoint its key layout
(``feature_extractor.conv_layers.{i}.0.weight`` ... ``encoder.layers.{n}.fc2.bias``).  All arithmetic
happens in libw2vs.so (hand-written sm_100a kernels) through the C ABI of include/w2vs.h; PyTorch
provides device memory and the CUDA stream.  There is no CPU / eager fallback: calling the model
with parameters that are not on a CUDA device raises.
"""
import argparse
import ctypes as C
import math
import random
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import cabi

DEFAULT_CONV_LAYERS = "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] + [(512,2,2)]"

# pre-training-only entries of a released checkpoint (wav2vec2.py:354-400); accepted and ignored
_PRETRAIN_PREFIXES = ("quantizer.", "project_q.", "final_proj.", "target_glu.", "input_quantizer.",
                      "project_inp.")

# Reference defaults of the fields that shape the path (Wav2VecSConfig, wav2vec_S.py:43-311; rain
# base_architecture, unidirect_w2v2_encoder.py:679-750).
_DEFAULTS = dict(
    extractor_mode="default", encoder_layers=12, encoder_embed_dim=768, encoder_ffn_embed_dim=3072,
    encoder_attention_heads=12, activation_fn="gelu", layer_norm_first=False,
    conv_feature_layers=DEFAULT_CONV_LAYERS, conv_bias=False, conv_pos=128, conv_pos_groups=16,
    pos_type="sin", context_type="constant", main_context=16, right_context=16,
    required_seq_len_multiple=2, load_pretrained_model_from="", quantize_input=False,
)


def base_architecture(args):
    """Fill every path-shaping field that is absent (rain unidirect_w2v2_encoder.py:679-750)."""
    for k, v in _DEFAULTS.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    return args


def _as_namespace(cfg) -> argparse.Namespace:
    if isinstance(cfg, dict):
        cfg = argparse.Namespace(**cfg)
    elif not isinstance(cfg, argparse.Namespace):
        cfg = argparse.Namespace(**{k: getattr(cfg, k) for k in dir(cfg) if not k.startswith("_")
                                    and not callable(getattr(cfg, k))})
    else:
        cfg = argparse.Namespace(**vars(cfg))
    return base_architecture(cfg)


class _Params(nn.Module):
    """Named parameter holder (no forward: arithmetic lives in the CUDA library)."""

    def __init__(self, **shapes):
        super().__init__()
        for name, shape in shapes.items():
            self.register_parameter(name, nn.Parameter(torch.zeros(shape)))

    def forward(self, *a, **k):
        raise RuntimeError("parameter holder: the forward path runs in libw2vs.so")


class _Slot(nn.Module):
    """Keeps the reference's nn.Sequential numbering for parameter-free modules."""

    def forward(self, *a, **k):
        raise RuntimeError("parameter holder: the forward path runs in libw2vs.so")


class ConvFeatureExtractionModel(nn.Module):
    """Key layout of wav2vec2.py:702-771 (``layer_norm_num`` variant of wav2vec-S)."""

    def __init__(self, conv_layers, dropout=0.0, mode="default", conv_bias=False, layer_norm_num=7):
        super().__init__()
        assert mode in {"default", "layer_norm"}
        self.conv_spec = list(conv_layers)
        self.mode, self.conv_bias, self.layer_norm_num = mode, conv_bias, layer_norm_num
        blocks = []
        c_in = 1
        for i, (dim, k, stride) in enumerate(self.conv_spec):
            conv = _Params(weight=(dim, c_in, k), **({"bias": (dim,)} if conv_bias else {}))
            nn.init.kaiming_normal_(conv.weight)                       # wav2vec2.py:726
            if mode == "layer_norm" and i < layer_norm_num:
                norm = nn.Sequential(_Slot(), _Params(weight=(dim,), bias=(dim,)), _Slot())
                nn.init.ones_(norm[1].weight)
                blocks.append(nn.Sequential(conv, _Slot(), norm, _Slot()))
            elif mode == "default" and i == 0:
                norm = _Params(weight=(dim,), bias=(dim,))
                nn.init.ones_(norm.weight)
                blocks.append(nn.Sequential(conv, _Slot(), norm, _Slot()))
            else:
                blocks.append(nn.Sequential(conv, _Slot(), _Slot()))
            c_in = dim
        self.conv_layers = nn.ModuleList(blocks)

    def has_norm(self, i):
        return (self.mode == "layer_norm" and i < self.layer_norm_num) or (self.mode == "default" and i == 0)


class _AttnParams(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.k_proj = _Params(weight=(D, D), bias=(D,))
        self.v_proj = _Params(weight=(D, D), bias=(D,))
        self.q_proj = _Params(weight=(D, D), bias=(D,))
        self.out_proj = _Params(weight=(D, D), bias=(D,))


class TransformerSentenceEncoderLayer(nn.Module):
    """Key layout of wav2vec2.py:874-919."""

    def __init__(self, D, F):
        super().__init__()
        self.self_attn = _AttnParams(D)
        self.self_attn_layer_norm = _Params(weight=(D,), bias=(D,))
        self.fc1 = _Params(weight=(F, D), bias=(F,))
        self.fc2 = _Params(weight=(D, F), bias=(D,))
        self.final_layer_norm = _Params(weight=(D,), bias=(D,))
        for lin in (self.self_attn.k_proj, self.self_attn.v_proj, self.self_attn.q_proj,
                    self.self_attn.out_proj, self.fc1, self.fc2):
            nn.init.normal_(lin.weight, mean=0.0, std=0.02)   # init_bert_params
        nn.init.ones_(self.self_attn_layer_norm.weight)
        nn.init.ones_(self.final_layer_norm.weight)


class _SinPos(nn.Module):
    """Stands in for SinusoidalPositionalEmbedding: its only state_dict entry is ``_float_tensor``."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_float_tensor", torch.zeros(1))


def sinusoidal_table(num_embeddings: int, embedding_dim: int, padding_idx: int = 1) -> torch.Tensor:
    """fp32 table, sin half then cos half, row `padding_idx` zero
    (modules/sinusoidal_positional_embedding.py:36-59)."""
    half_dim = embedding_dim // 2
    emb = math.log(10000) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, dtype=torch.float) * -emb)
    emb = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * emb.unsqueeze(0)
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1).view(num_embeddings, -1)
    if embedding_dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros(num_embeddings, 1)], dim=1)
    emb[padding_idx, :] = 0
    return emb


class BlockwiseTransformerEncoder(nn.Module):
    """Key layout and attributes of wav2vec_S.py:335-353 / wav2vec2.py:784-826."""

    def __init__(self, args):
        super().__init__()
        D = args.encoder_embed_dim
        self.embedding_dim = D
        self.pos_type = args.pos_type
        if self.pos_type == "conv":
            pc = _Params(bias=(D,), weight_g=(1, 1, args.conv_pos),
                         weight_v=(D, D // args.conv_pos_groups, args.conv_pos))
            std = math.sqrt(4.0 / (args.conv_pos * D))
            nn.init.normal_(pc.weight_v, mean=0, std=std)
            with torch.no_grad():
                pc.weight_g.copy_(pc.weight_v.norm(2, dim=(0, 1), keepdim=True))
            self.pos_conv = nn.Sequential(pc, _Slot(), _Slot())
        else:
            self.pos_conv = _SinPos()
        self.layers = nn.ModuleList([TransformerSentenceEncoderLayer(D, args.encoder_ffn_embed_dim)
                                     for _ in range(args.encoder_layers)])
        self.layer_norm_first = args.layer_norm_first
        self.layer_norm = _Params(weight=(D,), bias=(D,))
        nn.init.ones_(self.layer_norm.weight)
        self.required_seq_len_multiple = args.required_seq_len_multiple
        self.context_type = getattr(args, "context_type", "constant")
        self.main_context = getattr(args, "main_context", 16)
        self.right_context = getattr(args, "right_context", 8)

    def pick_context(self):
        """wav2vec_S.py:392-404 (note: "sampling" draws per call even in eval mode)."""
        if self.context_type == "sampling":
            main_context = random.randint(4, 16) * 2
            right_context = random.randint(2, 8) * 2
            return main_context, min(right_context, main_context // 2)
        if self.context_type == "constant":
            return self.main_context, self.right_context
        raise ValueError(
            "The mode of context_type: ({}) cannot be used. Please check.".format(self.context_type))


class Wav2VecSModel(nn.Module):
    """B200-native drop-in for the reference ``Wav2VecSModel`` encoder forward (features_only path)."""

    def __init__(self, cfg):
        super().__init__()
        args = _as_namespace(cfg)
        self.args = args
        if args.activation_fn != "gelu":
            raise NotImplementedError("activation_fn=%r: the released wav2vec-S models use gelu" % args.activation_fn)
        if getattr(args, "quantize_input", False):
            raise NotImplementedError("quantize_input is a pre-training option outside the encoder forward path")
        feature_enc_layers = eval(args.conv_feature_layers) if isinstance(args.conv_feature_layers, str) \
            else list(args.conv_feature_layers)
        self.embed = feature_enc_layers[-1][0]
        self.feature_extractor = ConvFeatureExtractionModel(
            conv_layers=feature_enc_layers, dropout=0.0, mode=args.extractor_mode,
            conv_bias=args.conv_bias, layer_norm_num=1 if args.encoder_layers == 12 else 7)
        D = args.encoder_embed_dim
        self.post_extract_proj = None
        if self.embed != D:
            self.post_extract_proj = _Params(weight=(D, self.embed), bias=(D,))
            nn.init.normal_(self.post_extract_proj.weight, mean=0.0, std=0.02)
        self.mask_emb = nn.Parameter(torch.zeros(D).uniform_())
        self.encoder = BlockwiseTransformerEncoder(args)
        self.layer_norm = _Params(weight=(self.embed,), bias=(self.embed,))
        nn.init.ones_(self.layer_norm.weight)
        self.max_positions = 8000          # wav2vec_S.py:341 (rain uses 2048; the table grows on demand)
        # Waveform front end.  ``source`` may be 16-bit PCM (torch.int16), read as
        # x / 32768 like the SimulEval agent does (rain/simul/transducer_searcher.py:74-80); with
        # ``normalize_waveform`` every utterance is standardised over its valid samples inside the first conv
        # layer's load, which is the data pipeline's `normalize: true` (raw_audio_dataset.py:60-72) -- feed RAW
        # samples then, not pre-normalised ones.  Full-utterance calls only (a stream has no utterance statistics).
        self.normalize_waveform = bool(getattr(args, "normalize_waveform", False))
        self._packed = None                # (key, uint8 device tensor)
        self._ws = None
        self._ccfg = None
        if getattr(args, "load_pretrained_model_from", ""):
            state = torch.load(args.load_pretrained_model_from, map_location="cpu")
            self.load_state_dict(state["model"], strict=False)

    # ---- reference API surface ----------------------------------------------------------------
    @classmethod
    def build_model(cls, cfg, task=None):
        return cls(cfg)

    def remove_pretraining_modules(self):
        """wav2vec2.py:695-699: the quantizer / projection heads never exist here."""
        return None

    # prefixes under which fine-tuned / CAAT checkpoints keep the wav2vec-S encoder
    # (wav2vec2_asr.py:290-372 `w2v_encoder.w2v_model.`, rain `encoder.w2v2_model.`, unidirect_w2v2_encoder.py:551-552)
    _CKPT_PREFIXES = ("encoder.w2v2_model.", "w2v_encoder.w2v_model.", "w2v2_model.", "w2v_model.")

    @classmethod
    def from_checkpoint(cls, ckpt, main_context=None, right_context=None, strict=False):
        """Build the encoder from a fairseq checkpoint: a path or the loaded
        dict ``{"cfg": {"model": ...} | "args": Namespace, "model": state_dict}`` that
        ``checkpoint_utils.load_checkpoint_to_cpu`` returns and ``OnlineW2V2TransformerEncoder.__init__``
        consumes (unidirect_w2v2_encoder.py:541-555): pre-trained wav2vec-S checkpoints as released, or a
        fine-tuned ASR / CAAT checkpoint in which the encoder sits under one of ``_CKPT_PREFIXES``.  Returns
        (model, missing_keys, unexpected_keys); keys of other sub-modules of a composite checkpoint (decoder,
        joiner, CTC projection) are reported as unexpected and ignored unless ``strict``."""
        if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
            ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
        if ckpt.get("args") is not None:
            margs = vars(ckpt["args"]).copy()
            margs["extractor_mode"] = "layer_norm"                 # rain :545-547 forces both for argparse-era
            margs["pos_type"] = "sin"                              # checkpoints (they predate the two options)
        else:
            model_cfg = ckpt["cfg"]["model"]
            margs = dict(model_cfg) if isinstance(model_cfg, dict) else dict(vars(model_cfg))
            inner = margs.get("w2v_args")                          # fine-tuned: the encoder's own args are nested
            if inner is not None:
                inner = inner["model"] if isinstance(inner, dict) and "model" in inner else inner
                margs = dict(inner) if isinstance(inner, dict) else dict(vars(inner))
        if main_context is not None:
            margs["main_context"] = main_context
        if right_context is not None:
            margs["right_context"] = right_context
        margs["load_pretrained_model_from"] = ""
        sd = ckpt["model"]
        for pre in cls._CKPT_PREFIXES:
            if any(k.startswith(pre) for k in sd):
                sd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
                break
        model = cls(margs)
        res = model.load_state_dict(sd, strict=strict)
        return model, list(res.missing_keys), list(res.unexpected_keys)

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.startswith(_PRETRAIN_PREFIXES)}
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._packed = None
        return out

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._ws = None
        return super()._apply(fn, *a, **k)

    def invalidate_packed_weights(self):
        """Call after mutating parameters in place (the packed copy is rebuilt on the next forward)."""
        self._packed = None

    def forward(self, source, padding_mask=None, mask=True, features_only=False):
        if mask or not features_only:
            raise NotImplementedError(
                "only the features_only / mask=False encoder forward is implemented "
                "(masking, quantizer and InfoNCE logits are pre-training paths)")
        x, pm = self._encode(source, padding_mask=padding_mask, layout=cabi.LAYOUT_BTD)
        return {"x": x, "padding_mask": pm}

    def extract_features(self, source, padding_mask, mask=False):
        res = self.forward(source, padding_mask, mask=mask, features_only=True)
        return res["x"], res["padding_mask"]

    # ---- C-ABI plumbing --------------------------------------------------------------------------
    def _device_dtype(self):
        p = self.layer_norm.weight
        if p.device.type != "cuda":
            raise RuntimeError("wav2vec-S B200 path: parameters must live on a CUDA device "
                               "(no CPU fallback exists); call .cuda() first")
        if p.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise RuntimeError(f"unsupported parameter dtype {p.dtype}: use float32, bfloat16 or float16")
        return p.device, p.dtype

    def _c_config(self, dtype, sin_rows):
        a = self.args
        c = cabi.Config()
        c.abi_version = cabi.W2VS_ABI_VERSION
        # fp16 models (`.half()`, what the reference trainer does under --fp16, trainer.py:86-90): fp16 in and out,
        # arithmetic of the bf16 path (w2vs_config.io_dtype)
        c.dtype = cabi.BF16 if dtype in (torch.bfloat16, torch.float16) else cabi.F32
        c.io_dtype = cabi.F16 if dtype == torch.float16 else 0
        spec = self.feature_extractor.conv_spec
        if len(spec) > cabi.W2VS_MAX_CONV:
            raise ValueError("too many conv layers")
        c.n_conv = len(spec)
        for i, (dim, k, s) in enumerate(spec):
            c.conv_dim[i], c.conv_kernel[i], c.conv_stride[i] = dim, k, s
        c.conv_bias = int(bool(a.conv_bias))
        c.extractor_mode = cabi.EXTRACTOR_LAYER_NORM if a.extractor_mode == "layer_norm" else cabi.EXTRACTOR_DEFAULT
        c.layer_norm_num = self.feature_extractor.layer_norm_num
        c.embed_dim, c.ffn_dim = a.encoder_embed_dim, a.encoder_ffn_embed_dim
        c.heads, c.layers = a.encoder_attention_heads, a.encoder_layers
        c.layer_norm_first = int(bool(a.layer_norm_first))
        c.pos_type = cabi.POS_CONV if a.pos_type == "conv" else cabi.POS_SIN
        c.conv_pos, c.conv_pos_groups = a.conv_pos, a.conv_pos_groups
        c.seq_multiple = a.required_seq_len_multiple
        c.sin_rows = sin_rows
        return c

    def _ref_tensors(self, sin_rows) -> List[torch.Tensor]:
        """Tensors in the order w2vs_weights_pack expects (include/w2vs.h)."""
        fe = self.feature_extractor
        out = []
        for i, blk in enumerate(fe.conv_layers):
            out.append(blk[0].weight)
            if fe.conv_bias:
                out.append(blk[0].bias)
            if fe.has_norm(i):
                norm = blk[2][1] if fe.mode == "layer_norm" else blk[2]
                out += [norm.weight, norm.bias]
        out += [self.layer_norm.weight, self.layer_norm.bias]
        if self.post_extract_proj is not None:
            out += [self.post_extract_proj.weight, self.post_extract_proj.bias]
        enc = self.encoder
        if enc.pos_type == "conv":
            pc = enc.pos_conv[0]
            out += [pc.bias, pc.weight_g, pc.weight_v]
        else:
            out.append(sinusoidal_table(sin_rows, enc.embedding_dim))
        for layer in enc.layers:
            sa = layer.self_attn
            for proj in (sa.q_proj, sa.k_proj, sa.v_proj, sa.out_proj):
                out += [proj.weight, proj.bias]
            out += [layer.self_attn_layer_norm.weight, layer.self_attn_layer_norm.bias,
                    layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias,
                    layer.final_layer_norm.weight, layer.final_layer_norm.bias]
        out += [enc.layer_norm.weight, enc.layer_norm.bias]
        return out

    def _ensure_packed(self, need_rows=0):
        device, dtype = self._device_dtype()
        sin_rows = max(self.max_positions + 2, need_rows)
        key = (str(device), dtype, sin_rows)
        if self._packed is not None and self._packed[0] == key:
            return self._ccfg, self._packed[1]
        lib = cabi.lib()
        ccfg = self._c_config(dtype, sin_rows)
        size = C.c_size_t()
        cabi.check(lib.w2vs_packed_weights_size(C.byref(ccfg), C.byref(size)), "w2vs_packed_weights_size")
        with torch.no_grad():
            tensors = [t.detach().to(device=device, dtype=torch.float32).contiguous()
                       for t in self._ref_tensors(sin_rows)]
        n = lib.w2vs_num_ref_tensors(C.byref(ccfg))
        assert n == len(tensors), (n, len(tensors))
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        packed = torch.empty(size.value, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            cabi.check(lib.w2vs_weights_pack(C.byref(ccfg), ptrs, n, packed.data_ptr(), size.value,
                                             C.c_void_p(stream)), "w2vs_weights_pack")
        self._packed, self._ccfg = (key, packed), ccfg
        return ccfg, packed

    def geometry(self, L, main_context=None, right_context=None):
        """Integer geometry (frames T, padded T', blocks, tokens M) for an L-sample input."""
        main = self.encoder.main_context if main_context is None else main_context
        rc = self.encoder.right_context if right_context is None else right_context
        ccfg = self._ccfg or self._c_config(torch.float32, self.max_positions + 2)
        g = cabi.Geometry()
        cabi.check(cabi.lib().w2vs_geometry_of(C.byref(ccfg), L, main, rc, C.byref(g)), "w2vs_geometry_of")
        return g

    def _workspace(self, ccfg, B, L, main, rc, device):
        size = C.c_size_t()
        cabi.check(cabi.lib().w2vs_get_workspace_size(C.byref(ccfg), B, L, main, rc, C.byref(size)),
                   "w2vs_get_workspace_size")
        if self._ws is None or self._ws.numel() < size.value or self._ws.device != device:
            self._ws = None
            self._ws = torch.empty(size.value, dtype=torch.uint8, device=device)
        return self._ws

    def _encode(self, source, padding_mask=None, lengths=None, mask_len=None, layout=cabi.LAYOUT_BTD,
                drop_tail=False, taps: Optional[Dict] = None, context=None):
        """One w2vs_encode call.  Padding is given either as the reference's bool ``padding_mask``
        [B, L'] or as ``lengths`` [B] (+ ``mask_len`` = width the reference mask would have had)."""
        device, dtype = self._device_dtype()
        if source.dim() != 2:
            raise ValueError("source must be [B, L]")
        if source.device != device:
            raise RuntimeError("source must be on the model's CUDA device")
        if source.dtype not in (torch.float32, torch.bfloat16, torch.int16):
            source = source.float()
        source = source.contiguous()
        B, L = source.shape
        main, rc = context if context is not None else self.encoder.pick_context()
        g = self.geometry(L, main, rc)
        ccfg, packed = self._ensure_packed(g.frames + 2)
        ws = self._workspace(ccfg, B, L, main, rc, device)
        D = self.args.encoder_embed_dim
        T = g.frames
        T_out = T - (rc if drop_tail else 0)
        if T_out < 0:
            raise ValueError("prefix shorter than right_context frames")
        out = torch.empty((B, T_out, D) if layout == cabi.LAYOUT_BTD else (T_out, B, D), dtype=dtype, device=device)
        out_mask = torch.empty((B, T_out), dtype=torch.bool, device=device)
        a = cabi.EncodeArgs()
        a.d_wav = source.data_ptr()
        a.wav_dtype = {torch.bfloat16: cabi.BF16, torch.int16: cabi.I16}.get(source.dtype, cabi.F32)
        a.wav_normalize = 1 if self.normalize_waveform else 0
        if self.normalize_waveform and lengths is None and padding_mask is not None:
            # the statistics need the valid length of each utterance: a length mask gives it, holes do not
            pmb = padding_mask.to(device=device, dtype=torch.bool)
            lengths = (~pmb).sum(1).to(torch.int32)
            if not torch.equal(pmb, torch.arange(pmb.size(1), device=device)[None, :] >= lengths[:, None]):
                raise ValueError("normalize_waveform needs a length-type padding mask")
            mask_len, padding_mask = pmb.size(1), None
        a.B, a.L = B, L
        keep = []
        if lengths is not None:
            lens = lengths.to(device=device, dtype=torch.int32).contiguous()
            keep.append(lens)
            a.d_lengths, a.mask_len = lens.data_ptr(), int(mask_len) if mask_len is not None else L
        elif padding_mask is not None:
            if padding_mask.dim() != 2 or padding_mask.size(0) != B:
                raise ValueError("padding_mask must be [B, L]")
            pm = padding_mask.to(device=device, dtype=torch.bool).contiguous()
            keep.append(pm)
            a.d_sample_pad_mask, a.mask_len = pm.data_ptr(), pm.size(1)
        a.main_ctx, a.right_ctx = main, rc
        a.out_layout = layout
        a.drop_tail_frames = rc if drop_tail else 0
        a.d_out, a.d_out_pad_mask = out.data_ptr(), out_mask.data_ptr()
        if taps is not None:
            M = g.tokens
            taps["conv_out"] = torch.empty((B, T, self.embed), dtype=torch.float32, device=device)
            taps["post_proj"] = torch.empty((B, T, D), dtype=torch.float32, device=device)
            taps["enc_in"] = torch.empty((B, M, D), dtype=torch.float32, device=device)
            taps["layers"] = torch.empty((self.args.encoder_layers, B, M, D), dtype=torch.float32, device=device)
            a.d_tap_conv_out, a.d_tap_post_proj = taps["conv_out"].data_ptr(), taps["post_proj"].data_ptr()
            a.d_tap_enc_in, a.d_tap_layers = taps["enc_in"].data_ptr(), taps["layers"].data_ptr()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            cabi.check(cabi.lib().w2vs_encode(C.byref(ccfg), packed.data_ptr(), C.byref(a), ws.data_ptr(),
                                              ws.numel(), C.c_void_p(stream)), "w2vs_encode")
        has_mask = lengths is not None or padding_mask is not None
        return out, (out_mask if has_mask else None)


class EncoderStream:
    """Incremental encoder state of B lock-step streams (w2vs_stream_*): conv carry buffers, the
    projected-frame buffer and the per-layer K/V cache of finalised main frames live on the device.

    ``step(new_samples, flush)`` feeds [B, n] new samples and returns the frames that became final,
    time-major [n_out, B, D].  flush: 0 none, 1 final (end of stream), 2 peek (also return the
    not-yet-final tail as the offline encoder would compute it on the current prefix)."""

    NONE, FINAL, PEEK = 0, 1, 2

    def __init__(self, model, B, max_frames, max_new_samples, main_context=None, right_context=None, step_impl=0):
        device, dtype = model._device_dtype()
        self.model, self.B, self.device, self.dtype = model, B, device, dtype
        self.main = model.encoder.main_context if main_context is None else main_context
        self.rc = model.encoder.right_context if right_context is None else right_context
        self.max_frames, self.max_new = int(max_frames), int(max_new_samples)
        ccfg, self.packed = model._ensure_packed(self.max_frames + 2)
        # step_impl (w2vs_config.stream_step_impl): 0 = automatic (one kernel of thread-block clusters per decision step
        # for one stream of a pre-LN bf16 model of an instantiated shape, else the operator chain), 1 = the
        # kernel-per-operator chain, 2 = the first persistent cooperative kernel (bf16, <= 32 tokens per step)
        self.ccfg = cabi.Config.from_buffer_copy(ccfg)
        self.ccfg.stream_step_impl = int(step_impl)
        lib = cabi.lib()
        hb, db, wb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        cabi.check(lib.w2vs_stream_state_size(C.byref(self.ccfg), B, self.max_frames, self.max_new, self.main,
                                              self.rc, C.byref(hb), C.byref(db), C.byref(wb)),
                   "w2vs_stream_state_size")
        self.host_state = C.create_string_buffer(hb.value)
        self.d_state = torch.empty(db.value, dtype=torch.uint8, device=device)
        self.ws = torch.empty(wb.value, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            cabi.check(lib.w2vs_stream_init(C.byref(self.ccfg), B, self.max_frames, self.max_new, self.main,
                                            self.rc, self.host_state, hb.value, self.d_state.data_ptr(), db.value,
                                            C.c_void_p(stream)), "w2vs_stream_init")
        self.D = model.args.encoder_embed_dim
        self.samples = 0
        self.out_cap = 0
        self._out = None

    def _grow(self, need_frames):
        """The stream outgrew its K/V cache: move it into buffers twice as large (w2vs_stream_grow) instead of
        failing -- the reference driver has no length limit other than --max-audio-positions."""
        new_max = max(2 * self.max_frames, int(need_frames))
        ccfg, self.packed = self.model._ensure_packed(new_max + 2)      # longer sinusoidal table if needed
        new_cfg = cabi.Config.from_buffer_copy(ccfg)
        new_cfg.stream_step_impl = self.ccfg.stream_step_impl
        lib = cabi.lib()
        hb, db, wb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        cabi.check(lib.w2vs_stream_state_size(C.byref(new_cfg), self.B, new_max, self.max_new, self.main, self.rc,
                                              C.byref(hb), C.byref(db), C.byref(wb)), "w2vs_stream_state_size")
        host_new = C.create_string_buffer(hb.value)
        d_new = torch.empty(db.value, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            cabi.check(lib.w2vs_stream_grow(C.byref(new_cfg), self.host_state, self.d_state.data_ptr(), new_max, host_new,
                                            hb.value, d_new.data_ptr(), db.value, C.c_void_p(stream)), "w2vs_stream_grow")
        # (the old device buffer is released in stream order by the caching allocator)
        self.ccfg, self.host_state, self.d_state, self.max_frames = new_cfg, host_new, d_new, new_max
        if self.ws.numel() < wb.value:
            self.ws = torch.empty(wb.value, dtype=torch.uint8, device=self.device)

    def step(self, new_samples=None, flush=0):
        n_new = 0 if new_samples is None else int(new_samples.size(1))
        if n_new and self.frames() + n_new // 320 + 2 > self.max_frames:
            self._grow(self.frames() + n_new // 320 + 2)
        src_ptr, wdt = None, cabi.F32
        if n_new:
            if new_samples.size(0) != self.B:
                raise ValueError("new_samples must be [B, n]")
            if new_samples.device != self.device:
                raise RuntimeError("new_samples must be on the model's CUDA device")
            if new_samples.dtype not in (torch.float32, torch.bfloat16, torch.int16):
                new_samples = new_samples.float()
            new_samples = new_samples.contiguous()
            src_ptr = new_samples.data_ptr()
            wdt = {torch.bfloat16: cabi.BF16, torch.int16: cabi.I16}.get(new_samples.dtype, cabi.F32)
        # every call can emit at most the frames the new samples complete plus the pending tail
        cap = n_new // 320 + 2 * (self.main + self.rc) + 2
        out = torch.empty((cap, self.B, self.D), dtype=self.dtype, device=self.device)
        n_out = C.c_int32(0)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            cabi.check(cabi.lib().w2vs_stream_step(C.byref(self.ccfg), self.packed.data_ptr(), self.host_state,
                                                   self.d_state.data_ptr(), src_ptr, wdt, n_new, int(flush),
                                                   out.data_ptr(), cap, C.byref(n_out), self.ws.data_ptr(),
                                                   self.ws.numel(), C.c_void_p(stream)), "w2vs_stream_step")
        self.samples += n_new
        return out[: n_out.value]

    def final_frames(self) -> int:
        """Frames of committed blocks (their outputs never change again)."""
        v = C.c_int32(0)
        cabi.check(cabi.lib().w2vs_stream_info(self.host_state, None, None, C.byref(v)), "w2vs_stream_info")
        return v.value

    def frames(self) -> int:
        v = C.c_int32(0)
        cabi.check(cabi.lib().w2vs_stream_info(self.host_state, None, C.byref(v), None), "w2vs_stream_info")
        return v.value


class BlockWiseWav2Vec2Model(Wav2VecSModel):
    """rain API (rain/layers/unidirect_w2v2_encoder.py:443-531): time-major output dict."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.max_positions = 2048           # rain :244

    def forward(self, source, padding_mask=None, incremental_state=None, finished=False, is_infer=False,
                src_lengths=None, mask_len=None):
        # The reference threads `incremental_state` through but never uses it (the code path is dead,
        # unidirect_w2v2_encoder.py:262-264) and re-encodes the whole prefix on every call.  Here a
        # dict passed as `incremental_state` together with is_infer=True switches to the cached
        # incremental path: `source` is still the whole prefix (as the reference driver passes it),
        # only the samples not seen before are encoded, and the return value is the same full-prefix
        # tensor the reference would have produced.
        if incremental_state is not None and is_infer:
            return self._forward_incremental(source, incremental_state, finished)
        drop = bool(is_infer and not finished and self.encoder.right_context > 0)
        ctx = (self.encoder.main_context, self.encoder.right_context)
        if src_lengths is not None:
            x, pm = self._encode(source, lengths=src_lengths, mask_len=mask_len, layout=cabi.LAYOUT_TBD,
                                 drop_tail=drop, context=ctx)
        else:
            x, pm = self._encode(source, padding_mask=padding_mask, layout=cabi.LAYOUT_TBD, drop_tail=drop,
                                 context=ctx)
        if pm is None:  # gen_block_atten_mask materialises an all-False mask (rain :80-81)
            pm = torch.zeros((x.size(1), x.size(0)), dtype=torch.bool, device=x.device)
        return {
            "encoder_out": [x],                  # T x B x C
            "encoder_padding_mask": [pm],        # B x T
            "encoder_embedding": [],
            "encoder_states": [],
            "src_tokens": [],
            "src_lengths": [],
            "dec1_state": [],
            "dec1_padding_mask": [],
        }

    def extract_features(self, source, padding_mask, mask=False):
        return self._encode(source, padding_mask=padding_mask, layout=cabi.LAYOUT_BTD,
                            context=(self.encoder.main_context, self.encoder.right_context))

    def open_stream(self, B=1, max_seconds=60.0, max_new_samples=4 * 5120 + 8000, step_impl=0):
        """Incremental encoder state for B lock-step live streams (see EncoderStream)."""
        max_frames = int(max_seconds * 16000) // 320 + 1
        return EncoderStream(self, B, max_frames, max_new_samples, step_impl=step_impl)

    def _forward_incremental(self, source, incremental_state, finished):
        st = incremental_state.get("w2vs_stream")
        B, L = source.shape
        if st is None:
            st = dict(stream=self.open_stream(B), frames=[])
            incremental_state["w2vs_stream"] = st
        stream, hist = st["stream"], st["frames"]
        if L < stream.samples or B != stream.B:
            raise ValueError("incremental forward needs a growing prefix of the same streams")
        pos = stream.samples
        while True:                       # feed at most max_new samples per call
            n = min(L - pos, stream.max_new)
            last = pos + n >= L
            flush = (EncoderStream.FINAL if finished else EncoderStream.PEEK) if last else EncoderStream.NONE
            before = stream.final_frames()
            chunk = source[:, pos:pos + n] if n else None
            if chunk is not None and chunk.device != stream.device:
                # host-resident prefix (what the SimulEval agent holds, rain/simul/transducer_searcher.py:728-731, where
                # the reference uploads the WHOLE prefix at every decision step): only the samples the device has
                # not seen cross the bus; conv carries, projected frames and K/V stay resident in the stream state
                chunk = chunk.to(stream.device, non_blocking=True)
            out = stream.step(chunk, flush)
            n_final = stream.final_frames() - before
            if n_final:
                hist.append(out[:n_final])   # frames of whole blocks never change again
            pos += n
            if last:
                break
        x = torch.cat(hist + [out[n_final:]], 0)   # + the tail as the reference computes it on this prefix
        rc = self.encoder.right_context
        if not finished and rc > 0:
            x = x[: max(x.size(0) - rc, 0)]
        pm = torch.zeros((B, x.size(0)), dtype=torch.bool, device=x.device)
        return {"encoder_out": [x], "encoder_padding_mask": [pm], "encoder_embedding": [], "encoder_states": [],
                "src_tokens": [], "src_lengths": [], "dec1_state": [], "dec1_padding_mask": []}


class OnlineW2V2TransformerEncoder(nn.Module):
    """rain/layers/unidirect_w2v2_encoder.py:534-607: checkpoint -> args -> model, length-based masks."""

    def __init__(self, args, wav2vec_ckpt=None):
        super().__init__()
        self.main_context = args.main_context
        self.right_context = args.right_context
        if wav2vec_ckpt is None:
            wav2vec_ckpt = torch.load(args.w2v2_model_path, map_location="cpu")
        if wav2vec_ckpt.get("args") is None:
            w2v2_args = argparse.Namespace(**wav2vec_ckpt["cfg"]["model"])
        else:
            w2v2_args = argparse.Namespace(**vars(wav2vec_ckpt["args"]))
            w2v2_args.extractor_mode = "layer_norm"
            w2v2_args.pos_type = "sin"
        w2v2_args.main_context = args.main_context
        w2v2_args.right_context = args.right_context
        w2v2_args.load_pretrained_model_from = ""
        self.w2v2_model = BlockWiseWav2Vec2Model.build_model(w2v2_args, task=None)
        self.w2v2_model.load_state_dict(wav2vec_ckpt["model"], strict=False)
        self.use_linear_layer = getattr(args, "use_linear_layer", False)
        self.encoder_proj = None
        if self.use_linear_layer and w2v2_args.encoder_embed_dim != args.encoder_embed_dim:
            self.encoder_proj = nn.Linear(w2v2_args.encoder_embed_dim, args.encoder_embed_dim)
        self.freeze_finetune_updates = getattr(args, "freeze_finetune_updates", -1)
        self.num_updates = 0

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates

    @property
    def init_frames(self):
        return self.main_context + self.right_context

    @property
    def step_frames(self):
        return self.main_context

    def forward(self, src_tokens, src_lengths, incremental_state=None, finished=False, is_infer=False):
        # lengths_to_padding_mask (fairseq/data/data_utils.py:528-532) builds a [B, max(len)] mask; the
        # kernel derives the same frame mask from (lengths, mask width) without materialising it.
        mask_len = int(torch.max(src_lengths).item())
        output = self.w2v2_model(src_tokens, None, incremental_state, finished, is_infer,
                                 src_lengths=src_lengths, mask_len=mask_len)
        if self.use_linear_layer and self.encoder_proj is not None:
            # unidirect_w2v2_encoder.py:590-594, through this library's GEMM (the step
            # right after the path stays on the same kernels and the same stream)
            from . import ops
            x = output["encoder_out"][0]
            T_, B_, D_ = x.shape
            # fp16 models run on the bf16 operand path (w2vs_config.io_dtype): so does this product
            gdt = torch.bfloat16 if x.dtype == torch.float16 else x.dtype
            w = self.encoder_proj.weight.detach().to(gdt).contiguous()
            y = ops.gemm(x.reshape(T_ * B_, D_).to(gdt), w, self.encoder_proj.bias.detach().float().contiguous(),
                         out_dtype=gdt).to(x.dtype) if T_ * B_ > 0 else x.new_zeros((0, w.size(0)))
            output = dict(output)
            output["encoder_out"] = [y.view(T_, B_, w.size(0))]
        return output

    def forward_torchscript(self, net_input: Dict[str, torch.Tensor]):
        return self.forward(src_tokens=net_input["src_tokens"], src_lengths=net_input["src_lengths"])

    def reorder_encoder_out(self, encoder_out, new_order):
        def sel(key, dim):
            v = encoder_out.get(key, [])
            return [v[0].index_select(dim, new_order)] if len(v) else []
        return {
            "encoder_out": sel("encoder_out", 1),
            "encoder_padding_mask": sel("encoder_padding_mask", 0),
            "encoder_embedding": sel("encoder_embedding", 0),
            "encoder_states": [s.index_select(1, new_order) for s in encoder_out.get("encoder_states", [])],
            "src_tokens": sel("src_tokens", 0),
            "src_lengths": sel("src_lengths", 0),
            "dec1_state": sel("dec1_state", 1),
            "dec1_padding_mask": sel("dec1_padding_mask", 0),
        }
