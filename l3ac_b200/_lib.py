"""ctypes binding of ``libl3ac_b200.so`` (the C ABI declared in ``include/l3ac_b200.h``).

There is no fallback: if the library is missing or a symbol is absent, importing the kernels raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libl3ac_b200.so"

F32, BF16, BF16X2 = 0, 1, 2
ACT_NONE, ACT_SNAKE, ACT_GEGLU, ACT_GELU, ACT_TANH = 0, 1, 2, 3, 4

_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class GemmDesc(C.Structure):
    """``l3ac_gemm_desc`` (include/l3ac_b200.h)."""
    _fields_ = [
        ("A", _p), ("W", _p), ("bias", _p), ("alpha", _p), ("scale", _p), ("shift", _p), ("residual", _p), ("out", _p),
        ("A_lo", _p), ("W_lo", _p), ("out_lo", _p),
        ("lda", _ll), ("ldr", _ll), ("ldo", _ll),
        ("B", _i), ("T", _i), ("K", _i), ("N", _i),
        ("taps", _i), ("tap_shift0", _i), ("tap_step", _i),
        ("act", _i), ("out_dtype", _i),
    ]


MAX_STAGES = 8


class CodecConfig(C.Structure):
    """``l3ac_codec_config`` (include/l3ac_b200.h, step-level interface)."""
    _fields_ = [
        ("feature_dim", _i), ("n_encoder_stages", _i), ("encoder_dims", _i * MAX_STAGES), ("encoder_depths", _i * MAX_STAGES),
        ("compress_rates", _i * MAX_STAGES), ("en_coder_depth", _i), ("en_coder_window_size", _i), ("en_coder_compress_rate", _i),
        ("en_coder_dynamic_pos", _i), ("n_levels", _i), ("levels", _i * 8), ("n_decoder_stages", _i),
        ("decoder_dims", _i * MAX_STAGES), ("decoder_depths", _i * MAX_STAGES), ("decode_rates", _i * MAX_STAGES),
        ("precision", _i),
    ]


class Tensor(C.Structure):
    """``l3ac_tensor``: a named host fp32 array."""
    _fields_ = [("name", C.c_char_p), ("data", _p), ("numel", _ll)]


# name -> (restype, argtypes); must list every symbol the header declares (tests check this against the header)
PROTOTYPES = {
    "l3ac_abi_version": (_i, []),
    "l3ac_error_string": (C.c_char_p, [_i]),
    "l3ac_stem": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "l3ac_stem_tc": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "l3ac_dwconv7_ln": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _i, _p]),
    "l3ac_dwconv_plan_create": (_i, [_i, _p, _p, _p, _p, _f, C.POINTER(_p)]),
    "l3ac_dwconv_plan_destroy": (_i, [_p]),
    "l3ac_dwconv7_ln_plan": (_i, [_p, _p, _i, _i, _p, _p]),
    "l3ac_layernorm": (_i, [_p, _ll, _i, _p, _p, _f, _p, _p, _i, _p]),
    "l3ac_split_bf16": (_i, [_p, _ll, _p, _p, _p]),
    "l3ac_snake": (_i, [_p, _ll, _i, _p, _p, _i, _p]),
    "l3ac_gemm_f32": (_i, [C.POINTER(GemmDesc), _p]),
    "l3ac_gemm_bf16_tc": (_i, [C.POINTER(GemmDesc), _p]),
    "l3ac_convunit_mlp_tc": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p]),
    "l3ac_convunit_mlp_tc_ch0": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p]),
    "l3ac_convunit_thin_f32": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "l3ac_convunit_thin_tc": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "l3ac_local_attention_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "l3ac_local_attention_tc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "l3ac_local_attention_umma": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "l3ac_rotary_pack": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p]),
    "l3ac_rotary_unpack": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "l3ac_fsq_quantize": (_i, [_p, _ll, _i, _p, _p, _p, _p, C.POINTER(_i), _i, _p, _p, _p, _p, _p]),
    "l3ac_fsq_quantize_latents": (_i, [_p, _ll, C.POINTER(_i), _i, _p, _p, _p, _p]),
    "l3ac_fsq_dequantize": (_i, [_p, _i, _ll, _i, _p, _p, C.POINTER(_i), _i, _p, _p]),
    "l3ac_upsample_linear_cn": (_i, [_p, _i, _i, _i, _i, _p, _p, _f, _p, _p]),
    "l3ac_updw_plan_create": (_i, [_i, _i, _p, _p, _f, _p, _p, _p, _p, _f, C.POINTER(_p)]),
    "l3ac_updw_plan_destroy": (_i, [_p]),
    "l3ac_upsample_cn_dwconv7_ln": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "l3ac_enhance_partials_floats": (_ll, [_i, _i]),
    "l3ac_enhance_stats": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "l3ac_enhance_apply": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "l3ac_enhup_plan_create": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, C.POINTER(_p)]),
    "l3ac_enhup_plan_destroy": (_i, [_p]),
    "l3ac_enhance_up": (_i, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "l3ac_tail_conv_tanh": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _p]),
    "l3ac_decoder_tail": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, C.POINTER(_i), _p, _p, _f, _p, _p]),
    "l3ac_convunit_plan_create": (_i, [_i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, C.POINTER(_p)]),
    "l3ac_convunit_plan_destroy": (_i, [_p]),
    "l3ac_convunit_umma": (_i, [_p, _p, _i, _i, _p, _p, _i, _p]),
    "l3ac_stem_plan_create": (_i, [_p, _p, _p, _p, _p, _p, _i, C.POINTER(_p)]),
    "l3ac_stem_plan_destroy": (_i, [_p]),
    "l3ac_stem_umma": (_i, [_p, _p, _i, _i, _p, _p]),
    "l3ac_tail_plan_create": (_i, [_p, _p, _p, _p, _p, _p, C.POINTER(_i), _p, _p, _f, _i, C.POINTER(_p)]),
    "l3ac_tail_plan_destroy": (_i, [_p]),
    "l3ac_decoder_tail_tc": (_i, [_p, _p, _i, _i, _p, _p]),
    "l3ac_decoder_tail_tc_split": (_i, [_p, _p, _i, _i, _p, _p]),
    # step-level interface (csrc/codec.cu)
    "l3ac_create": (_i, [C.POINTER(CodecConfig), C.POINTER(Tensor), _i, C.POINTER(_p)]),
    "l3ac_destroy": (_i, [_p]),
    "l3ac_last_error": (C.c_char_p, []),
    "l3ac_hop_length": (_i, [_p]),
    "l3ac_workspace_bytes": (_ll, [_p, _i, _i]),
    "l3ac_launch_count": (_ll, [_p]),
    "l3ac_encode": (_i, [_p, _p, _i, _i, _p, _ll, _p, _p, _p, _p]),
    "l3ac_decode": (_i, [_p, _p, _i, _p, _i, _i, _p, _ll, _p, _p]),
    "l3ac_quantize": (_i, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "l3ac_dequantize": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "l3ac_encode_host": (_i, [_p, _p, _i, _i, _p, _p]),
    "l3ac_decode_host": (_i, [_p, _p, _i, _i, _p]),
}

_lib = None


class L3acLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the CUDA library and bind every prototype.  Raises ``L3acLibraryError`` when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise L3acLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m l3ac_b200.build` (needs nvcc). "
            "l3ac_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise L3acLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().l3ac_error_string(code).decode()
        exc = ValueError if code in (-1, -2) else RuntimeError
        raise exc(f"{what} failed: {msg} (code {code})")
