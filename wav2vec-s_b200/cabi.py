"""ctypes binding of the C ABI declared in include/w2vs.h (libw2vs.so).

There is no fallback: if the library is missing or fails to load, importing the compute path
raises.  Build it with ``python wav2vec-s_b200/build.py`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# W2VS_LIBRARY selects another build of the same ABI (A/B timing of kernel variants on one GPU box)
LIB_PATH = os.environ.get("W2VS_LIBRARY") or os.path.join(HERE, "lib", "libw2vs.so")

W2VS_MAX_CONV = 8
W2VS_ABI_VERSION = 2
OK, INVALID_VALUE, UNSUPPORTED, WORKSPACE_TOO_SMALL, CUDA_ERROR = range(5)
F32, BF16, I16, F16 = 0, 1, 2, 3
EXTRACTOR_DEFAULT, EXTRACTOR_LAYER_NORM = 0, 1
POS_SIN, POS_CONV = 0, 1
LAYOUT_BTD, LAYOUT_TBD = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05_2CTA, GEMM_SKINNY = 0, 1, 3, 4
EPI_GELU = 1
EPI_SPLITK = 2    # in-place fp32 product of a few hundred rows may split K with L2 reductions (w2vs.h)


class W2vsError(RuntimeError):
    def __init__(self, status, what, detail):
        super().__init__(f"{what}: {detail or 'status %d' % status}")
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dtype", C.c_int32), ("n_conv", C.c_int32),
        ("conv_dim", C.c_int32 * W2VS_MAX_CONV), ("conv_kernel", C.c_int32 * W2VS_MAX_CONV),
        ("conv_stride", C.c_int32 * W2VS_MAX_CONV), ("conv_bias", C.c_int32),
        ("extractor_mode", C.c_int32), ("layer_norm_num", C.c_int32), ("embed_dim", C.c_int32),
        ("ffn_dim", C.c_int32), ("heads", C.c_int32), ("layers", C.c_int32),
        ("layer_norm_first", C.c_int32), ("pos_type", C.c_int32), ("conv_pos", C.c_int32),
        ("conv_pos_groups", C.c_int32), ("seq_multiple", C.c_int32), ("sin_rows", C.c_int32),
        ("stream_step_impl", C.c_int32), ("io_dtype", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class Geometry(C.Structure):
    _fields_ = [("frames", C.c_int32), ("frames_pad", C.c_int32), ("n_blocks", C.c_int32),
                ("tokens", C.c_int32), ("conv_len", C.c_int32 * W2VS_MAX_CONV),
                ("conv_rows", C.c_int32 * W2VS_MAX_CONV)]


class EncodeArgs(C.Structure):
    _fields_ = [
        ("d_wav", C.c_void_p), ("wav_dtype", C.c_int32), ("B", C.c_int32), ("L", C.c_int32),
        ("d_lengths", C.c_void_p), ("d_sample_pad_mask", C.c_void_p), ("mask_len", C.c_int32),
        ("main_ctx", C.c_int32), ("right_ctx", C.c_int32), ("out_layout", C.c_int32),
        ("drop_tail_frames", C.c_int32), ("d_out", C.c_void_p), ("d_out_pad_mask", C.c_void_p),
        ("d_tap_conv_out", C.c_void_p), ("d_tap_post_proj", C.c_void_p),
        ("d_tap_enc_in", C.c_void_p), ("d_tap_layers", C.c_void_p),
        ("wav_normalize", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


# name -> (restype, argtypes): every entry point include/w2vs.h declares
_P = C.POINTER
PROTOTYPES = {
    "w2vs_num_ref_tensors": (C.c_int32, [_P(Config)]),
    "w2vs_packed_weights_size": (C.c_int, [_P(Config), _P(C.c_size_t)]),
    "w2vs_weights_pack": (C.c_int, [_P(Config), _P(C.c_void_p), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "w2vs_geometry_of": (C.c_int, [_P(Config), C.c_int32, C.c_int32, C.c_int32, _P(Geometry)]),
    "w2vs_get_workspace_size": (C.c_int, [_P(Config), C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P(C.c_size_t)]),
    "w2vs_encode": (C.c_int, [_P(Config), C.c_void_p, _P(EncodeArgs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "w2vs_stream_state_size": (C.c_int, [_P(Config), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         _P(C.c_size_t), _P(C.c_size_t), _P(C.c_size_t)]),
    "w2vs_stream_init": (C.c_int, [_P(Config), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "w2vs_stream_step": (C.c_int, [_P(Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_void_p, C.c_int32, _P(C.c_int32), C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    "w2vs_stream_grow": (C.c_int, [_P(Config), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    "w2vs_stream_info": (C.c_int, [C.c_void_p, _P(C.c_int64), _P(C.c_int32), _P(C.c_int32)]),
    "w2vs_op_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_void_p]),
    "w2vs_op_layernorm": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "w2vs_op_attention": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "w2vs_status_string": (C.c_char_p, [C.c_int32]),
    "w2vs_last_error": (C.c_char_p, []),
    "w2vs_launch_count": (C.c_int64, [C.c_int32]),
    "w2vs_prof_enable": (None, [C.c_int32, C.c_void_p]),
    "w2vs_prof_collect": (C.c_int64, [C.c_char_p, C.c_int64]),
    "w2vs_debug_fused_trace": (C.c_int, [_P(C.c_uint64), C.c_int32]),
    "w2vs_debug_cluster_trace": (C.c_int, [_P(C.c_uint64), C.c_int32]),
    "w2vs_debug_fault_flags": (C.c_int, [_P(C.c_int32)]),
}

# kernel name -> class reported by bench.py
KERNEL_CLASS = {"gemm_tc_kernel": "gemm", "gemm_tc2_kernel": "gemm", "gemm_simt_kernel": "gemm_simt", "attn_mma_kernel": "attention", "attn_tc_kernel": "attention",
                "attn_simt_kernel": "attention", "conv0_kernel": "conv0", "stream_fused_kernel": "stream_fused",
                "stream_cluster_kernel": "stream_cluster", "conv_step_kernel": "conv_step", "feat_proj_kernel": "conv_step"}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python wav2vec-s_b200/build.py` "
                              "(there is no non-CUDA fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status, what):
    if status != OK:
        detail = lib().w2vs_last_error().decode(errors="replace")
        raise W2vsError(status, what, f"{lib().w2vs_status_string(status).decode()}: {detail}")


def launch_count(reset=False):
    return int(lib().w2vs_launch_count(1 if reset else 0))


def profile_step(fn, reps=1, detail=False):
    """Run ``fn`` (one step on the current stream) ``reps`` times with per-launch event timing;
    returns {class: {"ms": per-step device ms, "count": launches per step}}."""
    import torch
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    lib().w2vs_prof_enable(1, stream)
    for _ in range(reps):
        fn()
    need = lib().w2vs_prof_collect(None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    lib().w2vs_prof_collect(buf, len(buf))
    lib().w2vs_prof_enable(0, stream)
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.rsplit(" ", 2)
        cls = KERNEL_CLASS.get(name.split("[")[0], "rows")
        for key in (cls,) + ((name,) if detail else ()):
            d = out.setdefault(key, {"ms": 0.0, "count": 0})
            d["ms"] += float(ms) / reps
            d["count"] += int(cnt) // reps
    return out
