"""Generate golden vectors by running the UNMODIFIED reference (through oracle/ref_shim.py).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [case ...]
Writes tests/golden/<case>.npz holding outputs of the reference's own code
(Wav2VecSModel.extract_features, BlockWiseWav2Vec2Model.forward, and the prefix-recompute
streaming driver loop) for the seeded cases of oracle/cases.py.  Inputs/weights are pure
functions of the case entry, so only outputs are stored (fp16-free, fp32 arrays).
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, synth, cases  # noqa: E402
from oracle import w2vs_oracle as O  # noqa: E402

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))


def case_inputs(c):
    wav = synth.make_waveform(c["B"], c["L"], cases.XSEED)
    pm = None
    lens = None
    if c.get("ragged"):
        lens = synth.make_lengths(c["B"], c["L"], cases.LSEED)
        pm = O.lengths_to_padding_mask(lens)
        wav = wav.masked_fill(pm, 0.0)
    return wav, pm, lens


def run_case(name):
    c = cases.CASES[name]
    cfg = c["cfg"]
    sd = synth.make_state_dict(cfg, cases.WSEED)
    wav, pm, lens = case_inputs(c)
    api = c.get("api", "fairseq")
    out = {"cfg": json.dumps(cfg), "api": api}
    torch.manual_seed(0)
    with torch.no_grad():
        if api == "fairseq":
            m = ref_shim.build_fairseq_model(cfg)
            missing, unexpected = m.load_state_dict(sd, strict=False)
            assert not unexpected, unexpected
            assert all(k in ("mask_emb",) for k in missing), missing
            # per-stage taps through forward hooks on the reference modules
            taps = {}
            m.feature_extractor.register_forward_hook(lambda mod, i, o: taps.__setitem__("conv_out", o))
            if m.post_extract_proj is not None:
                m.post_extract_proj.register_forward_hook(lambda mod, i, o: taps.__setitem__("post_proj", o))
            m.encoder.layers[0].register_forward_hook(lambda mod, i, o: taps.__setitem__("layer0", o[0]))
            y, fm = m.extract_features(wav.clone(), pm)
            out["fmask"] = fm.numpy() if fm is not None else np.zeros((0,), dtype=bool)
            if c.get("compact"):
                # BASELINE-size case: the output only, every `compact`-th frame (see oracle/cases.py)
                out["y"] = y[:, ::c["compact"]].contiguous().numpy()
                out["y_shape"] = np.array(y.shape, dtype=np.int64)
            else:
                out["y"] = y.numpy()
                out["conv_out"] = taps["conv_out"].numpy()
                if "post_proj" in taps:
                    out["post_proj"] = taps["post_proj"].numpy()
                out["layer0"] = taps["layer0"].numpy()
        elif api == "rain":
            m = ref_shim.build_rain_model(cfg)
            missing, unexpected = m.load_state_dict(sd, strict=False)
            assert not unexpected and all(k in ("mask_emb",) for k in missing), (missing, unexpected)
            o = m(wav.clone(), pm, **c.get("kwargs", {}))
            out["y"] = o["encoder_out"][0].numpy()
            out["fmask"] = o["encoder_padding_mask"][0].numpy()
        elif api == "stream":
            # the reference's streaming driver: prefix recompute with is_infer=True
            # (rain/simul/transducer_agent.py:149-153, transducer_searcher.py:712-731)
            m = ref_shim.build_rain_model(cfg)
            m.load_state_dict(sd, strict=False)
            main, rc = cfg["main_context"], cfg["right_context"]
            L = wav.size(1)
            k, emitted, chunks, ns = 0, 0, [], []
            while True:
                n = O._samples_for_frames(cfg, main + rc + k * main)
                fin = n >= L
                n = min(n, L)
                o = m(wav[:, :n].clone(), None, None, fin, True)
                x = o["encoder_out"][0]
                chunks.append(x[emitted:].numpy())
                ns.append(n)
                emitted = x.size(0)
                if fin:
                    break
                k += 1
            out["y"] = np.concatenate(chunks, axis=0)
            out["chunk_sizes"] = np.array([c_.shape[0] for c_ in chunks], dtype=np.int64)
            out["prefix_samples"] = np.array(ns, dtype=np.int64)
            # offline output of the whole utterance, for the incremental == offline property
            off = m(wav.clone(), None, None, True, True)["encoder_out"][0]
            out["y_offline"] = off.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: y{tuple(out['y'].shape)} -> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.CASES)
    for n in names:
        run_case(n)
