"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np
import torch

from oracle import cases, synth
from oracle import w2vs_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["cfg"] = json.loads(str(g["cfg"]))
    g["api"] = str(g["api"])
    return g


def case_inputs(name):
    """(cfg, state_dict, wav, sample padding mask|None, lengths|None) -- same recipe as
    tests/golden/make_golden.py."""
    c = cases.CASES[name]
    sd = synth.make_state_dict(c["cfg"], cases.WSEED)
    wav = synth.make_waveform(c["B"], c["L"], cases.XSEED)
    pm, lens = None, None
    if c.get("ragged"):
        lens = synth.make_lengths(c["B"], c["L"], cases.LSEED)
        pm = O.lengths_to_padding_mask(lens)
        wav = wav.masked_fill(pm, 0.0)
    return c["cfg"], sd, wav, pm, lens


def valid_rel_err(y, ref, fmask_bt=None, time_first=False):
    """max|y-ref| / max|ref| over non-padded frames."""
    y = torch.as_tensor(y).float()
    ref = torch.as_tensor(ref).float()
    if time_first:
        y, ref = y.transpose(0, 1), ref.transpose(0, 1)
    if fmask_bt is not None and np.size(fmask_bt) > 0:
        keep = ~torch.as_tensor(fmask_bt).bool()
        y, ref = y[keep], ref[keep]
    return float((y - ref).abs().max() / ref.abs().max())
