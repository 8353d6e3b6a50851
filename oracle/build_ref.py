"""TEST INFRASTRUCTURE ONLY -- recipe that makes the reference's own implementation travel to the GPU box.

    python oracle/build_ref.py          (build container only: needs /root/reference)

Byte-compiles the UNMODIFIED reference source files of the hot path, where they lie under /root/reference, into
``oracle/_ref/`` (marshalled code objects, ``*.code`` -- no reference source text is copied into this repository; ``oracle/_ref/`` is
git-ignored like any other build output but is not gpurun-ignored, so it ships with the snapshot exactly like the
``.so`` this repository builds).  ``oracle/ref_shim.py`` loads the modules from /root/reference when that exists and
from these compiled files otherwise, which lets

  * ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` time the reference's own code on the box's host cores
    (``cpu_baseline.kind == "reference"``), and
  * ``tests/test_oracle_vs_reference.py`` re-pin the oracle against the live reference on the GPU box too.

The file list is the one SURVEY.md section 8(c) names (and ``ref_shim`` documents); the interpreter on the box is
the same image's Python, which is what the bytecode magic number requires (checked at load time).
"""
import marshal
import os
import sys

REF_ROOT = os.environ.get("W2VS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

FILES = [
    "fairseq/fairseq/incremental_decoding_utils.py",
    *[f"fairseq/fairseq/modules/{n}.py" for n in (
        "fairseq_dropout", "quant_noise", "multihead_attention", "fp32_group_norm", "layer_norm", "transpose_last",
        "same_pad", "grad_multiply", "gumbel_vector_quantizer", "sinusoidal_positional_embedding")],
    *[f"fairseq/fairseq/models/wav2vec/{n}.py" for n in ("utils", "wav2vec2", "wav2vec_S")],
    "rain/layers/unidirect_w2v2_encoder.py",
]


def build(verbose=True):
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"{REF_ROOT} not present: oracle/_ref can only be built in the build container")
    n = 0
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, rel[:-3] + ".code")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            # the code object keeps the reference's own path, so tracebacks show the file:line the oracle cites
            with open(src, "rb") as f:
                code = compile(f.read(), os.path.join("/root/reference", rel), "exec", dont_inherit=True, optimize=0)
            with open(dst, "wb") as f:
                f.write(marshal.dumps(code))
            n += 1
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + "\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference modules ({n} compiled now) for Python {sys.version.split()[0]}")
    return OUT


if __name__ == "__main__":
    build()
