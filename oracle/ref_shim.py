"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference sources.

Executes the reference's own hot-path files under a stub ``fairseq`` namespace (``import
fairseq`` itself fails here: no omegaconf/hydra, py3.12).  No reference source is copied into
this repository: in the build container the files are loaded where they lie under
``/root/reference``; on the GPU box (where that path does not exist) the same modules are loaded
from ``oracle/_ref/`` -- marshalled code objects of the unmodified files, produced by ``oracle/build_ref.py``,
git-ignored, shipped with the snapshot like a built ``.so``.  Used for:

  * ``tests/golden/make_golden.py``  -- generate golden input/output vectors,
  * ``tests/test_oracle_vs_reference.py`` -- pin ``oracle/w2vs_oracle.py`` against the real
    reference (skipped when neither form of the reference is present),
  * ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` -- time the reference's own code on
    the host cores (``kind: "reference"``).

Loaded verbatim (reference paths relative to /root/reference):
  fairseq/fairseq/incremental_decoding_utils.py
  fairseq/fairseq/modules/{fairseq_dropout,quant_noise,multihead_attention,fp32_group_norm,
      layer_norm,transpose_last,same_pad,grad_multiply,gumbel_vector_quantizer,
      sinusoidal_positional_embedding}.py
  fairseq/fairseq/models/wav2vec/{utils,wav2vec2,wav2vec_S}.py
  rain/layers/unidirect_w2v2_encoder.py
Stubbed (each a few-line restatement of the cited function):
  fairseq.utils.index_put (utils.py:705-714), softmax (:487-491), make_positions (:250-260),
  get_activation_fn (:517-537 + modules/gelu.py:24-25), buffered_arange (:267-273),
  fairseq.modules.transformer_sentence_encoder.init_bert_params (:21-53),
  fairseq.data.data_utils.lengths_to_padding_mask (data_utils.py:528-532).
"""
import importlib.util
import os
import sys
import types
import argparse

import torch
import torch.nn as nn
import torch.nn.functional as F

import importlib.machinery

REF_ROOT = os.environ.get("W2VS_REFERENCE_ROOT", "/root/reference")
PYC_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_KEY = os.path.join("fairseq", "fairseq", "models", "wav2vec", "wav2vec_S.py")
_loaded = {}


def source_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, _KEY))


def compiled_available() -> bool:
    if not os.path.isfile(os.path.join(PYC_ROOT, _KEY[:-3] + ".code")):
        return False
    try:      # code objects are only valid for the interpreter version that made them
        return open(os.path.join(PYC_ROOT, "PYTHON_VERSION")).read().strip() == sys.version.split()[0]
    except OSError:
        return False


def available() -> bool:
    return source_available() or compiled_available()


def kind() -> str:
    """Where the reference modules come from: "source" (/root/reference), "compiled" (oracle/_ref) or "absent"."""
    return "source" if source_available() else ("compiled" if compiled_available() else "absent")


_ROOT = REF_ROOT if source_available() else PYC_ROOT
_EXT = ".py" if source_available() else ".code"
_FS = os.path.join(_ROOT, "fairseq", "fairseq")


def _mod(name, is_pkg=False):
    m = types.ModuleType(name)
    if is_pkg:
        m.__path__ = []
    sys.modules[name] = m
    return m


def _load(name, path):
    if _EXT == ".code":
        import marshal
        m = types.ModuleType(name)
        m.__file__ = path[:-3] + ".code"
        sys.modules[name] = m
        with open(m.__file__, "rb") as f:
            exec(marshal.loads(f.read()), m.__dict__)
        return m
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def _install():
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found: neither sources under {REF_ROOT} nor oracle/_ref (oracle/build_ref.py)")
    if "fairseq" in sys.modules and not getattr(sys.modules["fairseq"], "_w2vs_shim", False):
        raise RuntimeError("a real fairseq is already imported; shim refuses to shadow it")

    fairseq = _mod("fairseq", True)
    fairseq._w2vs_shim = True

    # ---- fairseq.utils (stubs of the five functions the path uses)
    utils = _mod("fairseq.utils")

    def index_put(tensor, indices, value):
        tensor[indices] = value
        return tensor

    def softmax(x, dim, onnx_trace=False):
        return F.softmax(x, dim=dim, dtype=torch.float32)

    def make_positions(tensor, padding_idx, onnx_trace=False):
        mask = tensor.ne(padding_idx).int()
        return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx

    def gelu(x):
        return F.gelu(x.float()).type_as(x)

    def get_activation_fn(activation):
        if activation == "gelu":
            return gelu
        if activation == "relu":
            return F.relu
        raise RuntimeError(f"activation {activation} not stubbed")

    def buffered_arange(max):
        return torch.arange(max)

    utils.index_put = index_put
    utils.softmax = softmax
    utils.make_positions = make_positions
    utils.get_activation_fn = get_activation_fn
    utils.get_available_activation_fns = lambda: ["relu", "gelu"]
    utils.buffered_arange = buffered_arange
    fairseq.utils = utils

    # ---- fairseq.dataclass
    dc = _mod("fairseq.dataclass", True)

    def ChoiceEnum(choices):
        return str

    class FairseqDataclass:
        pass

    dc.ChoiceEnum = ChoiceEnum
    dc.FairseqDataclass = FairseqDataclass
    dcu = _mod("fairseq.dataclass.utils")
    dcu.convert_namespace_to_omegaconf = lambda x: x

    # ---- fairseq.models
    models = _mod("fairseq.models", True)

    class BaseFairseqModel(nn.Module):
        def upgrade_state_dict_named(self, state_dict, name):
            return state_dict

    class FairseqEncoder(nn.Module):
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

        def set_num_updates(self, n):
            pass

    def register_model(name, dataclass=None):
        return lambda cls: cls

    def register_model_architecture(a, b):
        return lambda fn: fn

    models.BaseFairseqModel = BaseFairseqModel
    models.FairseqEncoder = FairseqEncoder
    models.register_model = register_model
    models.register_model_architecture = register_model_architecture

    # ---- fairseq.data
    data = _mod("fairseq.data", True)
    du = _mod("fairseq.data.data_utils")
    du.compute_mask_indices = None

    def lengths_to_padding_mask(lens):
        bsz, max_lens = lens.size(0), torch.max(lens).item()
        mask = torch.arange(max_lens).to(lens.device).view(1, max_lens)
        return mask.expand(bsz, -1) >= lens.view(bsz, 1).expand(-1, max_lens)

    du.lengths_to_padding_mask = lengths_to_padding_mask
    data.data_utils = du

    class Dictionary:
        pass

    data.Dictionary = Dictionary
    fairseq.options = _mod("fairseq.options")
    fairseq.checkpoint_utils = _mod("fairseq.checkpoint_utils")

    # ---- real reference files
    _load("fairseq.incremental_decoding_utils", os.path.join(_FS, "incremental_decoding_utils.py"))
    modules = _mod("fairseq.modules", True)
    order = ["fairseq_dropout", "quant_noise", "multihead_attention", "fp32_group_norm",
             "layer_norm", "transpose_last", "same_pad", "grad_multiply",
             "gumbel_vector_quantizer", "sinusoidal_positional_embedding"]
    for n in order:
        m = _load(f"fairseq.modules.{n}", os.path.join(_FS, "modules", n + ".py"))
        setattr(modules, n, m)
    modules.FairseqDropout = modules.fairseq_dropout.FairseqDropout
    modules.MultiheadAttention = modules.multihead_attention.MultiheadAttention
    modules.Fp32GroupNorm = modules.fp32_group_norm.Fp32GroupNorm
    modules.Fp32LayerNorm = modules.layer_norm.Fp32LayerNorm
    modules.LayerNorm = modules.layer_norm.LayerNorm
    modules.TransposeLast = modules.transpose_last.TransposeLast
    modules.SamePad = modules.same_pad.SamePad
    modules.GradMultiply = modules.grad_multiply.GradMultiply
    modules.GumbelVectorQuantizer = modules.gumbel_vector_quantizer.GumbelVectorQuantizer
    modules.SinusoidalPositionalEmbedding = (
        modules.sinusoidal_positional_embedding.SinusoidalPositionalEmbedding)
    modules.gelu = gelu

    tse = _mod("fairseq.modules.transformer_sentence_encoder")

    def init_bert_params(module):
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.bias is not None:
                module.bias.data.zero_()
        if isinstance(module, modules.MultiheadAttention):
            module.q_proj.weight.data.normal_(mean=0.0, std=0.02)
            module.k_proj.weight.data.normal_(mean=0.0, std=0.02)
            module.v_proj.weight.data.normal_(mean=0.0, std=0.02)

    tse.init_bert_params = init_bert_params

    w2v = _mod("fairseq.models.wav2vec", True)
    wdir = os.path.join(_FS, "models", "wav2vec")
    _load("fairseq.models.wav2vec.utils", os.path.join(wdir, "utils.py"))
    w2 = _load("fairseq.models.wav2vec.wav2vec2", os.path.join(wdir, "wav2vec2.py"))
    for k in ("Wav2Vec2Model", "TransformerEncoder", "TransformerSentenceEncoderLayer",
              "ConvFeatureExtractionModel", "EXTRACTOR_MODE_CHOICES",
              "MASKING_DISTRIBUTION_CHOICES", "LAYER_TYPE_CHOICES", "base_architecture"):
        setattr(w2v, k, getattr(w2, k))
    ws = _load("fairseq.models.wav2vec.wav2vec_S", os.path.join(wdir, "wav2vec_S.py"))
    rain = _load("w2vs_ref_rain_unidirect_w2v2_encoder",
                 os.path.join(_ROOT, "rain", "layers", "unidirect_w2v2_encoder.py"))
    _loaded.update(wav2vec2=w2, wav2vec_S=ws, rain=rain, modules=modules)
    return _loaded


def reference_namespace(cfg: dict) -> argparse.Namespace:
    """Namespace carrying every field the reference constructors read.

    ``cfg`` uses the reference's own field names (Wav2VecSConfig, wav2vec_S.py:43-311).
    """
    mods = _install()
    ns = argparse.Namespace(**cfg)
    mods["rain"].base_architecture(ns)  # fills all remaining defaults (rain :679-750)
    for k, v in dict(load_pretrained_model_from="", pos_type="sin", context_type="constant",
                     required_seq_len_multiple=2).items():
        if not hasattr(ns, k):
            setattr(ns, k, v)
    return ns


def build_fairseq_model(cfg: dict):
    """Reference ``Wav2VecSModel`` (wav2vec_S.py:314-332), eval mode, pre-training heads removed."""
    mods = _install()
    ns = reference_namespace(cfg)
    m = mods["wav2vec_S"].Wav2VecSModel(ns)
    m.remove_pretraining_modules()
    return m.eval()


def build_rain_model(cfg: dict):
    """Reference ``BlockWiseWav2Vec2Model`` (rain/layers/unidirect_w2v2_encoder.py:443-531)."""
    mods = _install()
    ns = reference_namespace(cfg)
    m = mods["rain"].BlockWiseWav2Vec2Model(ns)
    m.remove_pretraining_modules()
    return m.eval()


def lengths_to_padding_mask(lens):
    _install()
    return sys.modules["fairseq.data.data_utils"].lengths_to_padding_mask(lens)
