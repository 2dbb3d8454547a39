"""ctypes binding of libparrot_b200.so (include/parrot_b200.h).

The product path has no fallback: if the shared library is missing, or a call fails, this module
raises - it never reroutes to PyTorch ops or to the CPU oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PB_LIB_VARIANT=NAME loads an experiment build (python -m parrot_tts_b200.build --variant NAME --define ...) instead of the product
_VARIANT = os.environ.get("PB_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "lib", f"libparrot_b200_{_VARIANT}.so" if _VARIANT else "libparrot_b200.so")

PB_OK, PB_ERR_INVALID, PB_ERR_INDEX, PB_ERR_CUDA, PB_ERR_STATE, PB_ERR_NCCL = 0, -1, -2, -3, -4, -5

PB_VOC_MAX_UPS, PB_VOC_MAX_RESK, PB_VOC_MAX_DIL = 8, 4, 4


class VocCfg(C.Structure):
    _fields_ = [
        ("model_in_dim", C.c_int32),
        ("upsample_initial_channel", C.c_int32),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * PB_VOC_MAX_UPS),
        ("upsample_kernel_sizes", C.c_int32 * PB_VOC_MAX_UPS),
        ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * PB_VOC_MAX_RESK),
        ("num_dilations", C.c_int32),
        ("resblock_dilation_sizes", (C.c_int32 * PB_VOC_MAX_DIL) * PB_VOC_MAX_RESK),
        ("num_embeddings", C.c_int32),
        ("embedding_dim", C.c_int32),
        ("num_speakers", C.c_int32),
        ("conv_mode", C.c_int32),
    ]


class TteCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "src_vocab_size", "src_pad_idx", "n_speaker", "d_model", "n_head", "n_enc_layer", "n_dec_layer",
        "conv_n_filter", "conv_k1", "conv_k2", "dp_n_filter", "dp_kernel", "n_codes", "max_len",
        "fold_projections", "gemm_mode")]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float

# name -> (restype, argtypes); mirrors include/parrot_b200.h one to one
SIGNATURES = {
    "pb_version": (C.c_int, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_voc_create": (C.c_int, [C.POINTER(VocCfg), C.POINTER(_P)]),
    "pb_voc_destroy": (None, [_P]),
    "pb_voc_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(_I64), _I32, _I32]),
    "pb_voc_finalize": (C.c_int, [_P, _P]),
    "pb_voc_forward": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P, _P]),
    "pb_voc_forward_pcm16": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P]),
    "pb_wav_to_pcm16": (C.c_int, [_P, _I64, _P, _P]),
    "pb_mel_spectrogram": (C.c_int, [_P, _I32, _I64, _P, _I32, _I32, _I32, _I32, _P, C.POINTER(_I32), _P]),
    "pb_pcm16_peak_normalize": (C.c_int, [_P, _I32, _I64, _P, _P, _P]),
    "pb_voc_last_launches": (_I64, [_P]),
    "pb_voc_memory_epoch": (_I64, [_P]),
    "pb_voc_check_ids": (C.c_int, [_P, _P]),
    "pb_voc_set_timing": (C.c_int, [_P, _I32]),
    "pb_voc_section_count": (C.c_int, [_P]),
    "pb_voc_section_ms": (C.c_int, [_P, C.POINTER(_F), _I32]),
    "pb_voc_arena_bytes": (C.c_int, [_P, C.POINTER(_I64)]),
    "pb_voc_arena_ptr": (_P, [_P]),
    "pb_voc_mark_finalized": (C.c_int, [_P]),
    "pb_conv1d": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _F, _F, _P]),
    "pb_conv_transpose1d": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _F, _P]),
    "pb_conv1d_tc": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _F, _F, _I32, _P]),
    "pb_conv_transpose1d_tc": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _F, _I32, _P]),
    "pb_tte_create": (C.c_int, [C.POINTER(TteCfg), C.POINTER(_P)]),
    "pb_tte_destroy": (None, [_P]),
    "pb_tte_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(_I64), _I32, _I32]),
    "pb_tte_finalize": (C.c_int, [_P, _P]),
    "pb_tte_encode": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, C.POINTER(_I32), _P]),
    "pb_tte_decode": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "pb_tte_last_launches": (_I64, [_P]),
    "pb_tte_arena_bytes": (C.c_int, [_P, C.POINTER(_I64)]),
    "pb_tte_arena_ptr": (_P, [_P]),
    "pb_tte_mark_finalized": (C.c_int, [_P]),
    "pb_layernorm": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P]),
    "pb_tte_fft_block": (C.c_int, [_P, _I32, _I32, _P, _P, _I32, _I32, _P, _P]),
    "pb_tte_duration": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _P, _P]),
    "pb_length_regulate": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "pb_linear": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
}

_lib = None


class ParrotB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (building is the job of parrot_tts_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ParrotB200Error(
            f"{LIB_PATH} not found: build it with `python -m parrot_tts_b200.build` "
            "(there is no PyTorch/CPU fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pb_version() != 3:
        raise ParrotB200Error(f"ABI version mismatch: library reports {lib.pb_version()}")
    _lib = lib
    return lib


def last_error():
    return load().pb_last_error().decode("utf-8", "replace")


def check(status):
    if status == PB_OK:
        return
    msg = last_error()
    if status == PB_ERR_INDEX:
        raise IndexError(msg)       # mirrors the reference's IndexError at modules/fft.py:18
    raise ParrotB200Error(f"libparrot_b200 error {status}: {msg}")


def ptr(t):
    """Device (or host) address of a tensor, or None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
