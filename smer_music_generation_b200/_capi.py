"""ctypes binding of libsmer_b200.so (declared in include/smer_b200.h).

The product path has no CPU or eager-PyTorch fallback: if the library is missing or a kernel
returns an error, a RuntimeError is raised (the reference's callers catch exceptions per batch,
train.py:917-926, so raising -- not aborting -- is the compatible behaviour).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsmer_b200.so")

F32, BF16 = 0, 1
EPI_RELU, EPI_ACCUM, EPI_ATOMIC, EPI_GATE = 1, 2, 4, 8
ST_PITCH, ST_REST, ST_SEP, ST_CONTINUE = 1, 2, 4, 8
SAMPLE_GREEDY, SAMPLE_MULTINOMIAL, SAMPLE_TOP_P, SAMPLE_TOP_K = 0, 1, 2, 3
XENT_MAX_SUMS = 16

_vp, _ll, _i, _f, _u64 = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_uint64


class AttnArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("v", _vp), ("o", _vp), ("dout", _vp), ("dq", _vp), ("dk", _vp), ("dv", _vp),
                ("dbq", _vp), ("dbk", _vp), ("dbv", _vp), ("ldq", _ll), ("ldk", _ll), ("ldv", _ll), ("ldo", _ll), ("lddo", _ll), ("lddq", _ll), ("lddk", _ll),
                ("lddv", _ll), ("lse", _vp), ("dsum", _vp), ("key_pad", _vp), ("kv_len", _vp), ("add_mask", _vp),
                ("ld_mask", _ll), ("B", _i), ("H", _i), ("Lq", _i), ("Lk", _i), ("dh", _i), ("dtype", _i),
                ("causal", _i), ("q_pos0", _i), ("scale", _f), ("dropout_p", _f), ("seed", _u64), ("site", _u64), ("dq_accum", _vp), ("cu_q", _vp), ("cu_k", _vp), ("q_rows", _ll), ("k_rows", _ll)]


class DecodeAttnArgs(C.Structure):
    _fields_ = [("q", _vp), ("new_k", _vp), ("new_v", _vp), ("k_cache", _vp), ("v_cache", _vp), ("out", _vp),
                ("kv_len", _vp), ("key_pad", _vp), ("workspace", _vp), ("ldq", _ll), ("ld_new", _ll), ("ldo", _ll),
                ("ld_cache", _ll), ("cache_stride", _ll), ("ld_pad", _ll), ("n_seq", _i), ("H", _i), ("dh", _i),
                ("cache_len", _i), ("splits", _i), ("dtype", _i), ("scale", _f), ("done", _vp)]


class SampleArgs(C.Structure):
    _fields_ = [("logits", _vp), ("ld", _ll), ("n_seq", _i), ("V", _i), ("mode", _i), ("temperature", C.c_double),
                ("top_p", C.c_double), ("top_k", _i), ("seed", _u64), ("seq_base", _ll), ("step_base", _u64),
                ("state", _vp), ("targets", _vp), ("nwd", _vp), ("max_spans", _i), ("raw_flags", _vp),
                ("raw_only_lo", _vp), ("raw_only_hi", _vp), ("tok_buf", _vp), ("cur_len", _vp), ("span_start", _vp),
                ("span_idx", _vp), ("fed_len", _vp), ("n_spans", _vp), ("done", _vp), ("gen_count", _vp), ("control_bitmap", _vp),
                ("max_len", _i), ("max_span", _i), ("out_token", _vp), ("out_probs", _vp), ("trace_masked", _vp), ("trace_span", _vp)]


_SIGS = {
    "smer_version": (C.c_int, []),
    "smer_last_error": (C.c_char_p, []),
    "smer_device_ok": (C.c_int, []),
    "smer_set_seed_device_ptr": (_i, [_vp]),
    "smer_embed_pe_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _u64, _u64, _vp]),
    "smer_embed_pe_packed": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _f, _f, _u64, _u64, _vp]),
    "smer_pack_rows": (_i, [_vp, _vp, _i, _i, _ll, _vp, _vp, _vp]),
    "smer_zero_tail_rows": (_i, [_vp, _ll, _ll, _vp, _vp]),
    "smer_embed_bwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _f, _f, _u64, _u64, _vp]),
    "smer_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _f, _f, _u64, _u64, _vp]),
    "smer_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _f, _u64, _u64, _vp]),
    "smer_gemm_simt": (_i, [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _ll, _i, _f, _u64,
                            _u64, _i, _vp]),
    "smer_gemm_bf16_tc": (_i, [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _ll, _i, _f, _u64,
                               _u64, _i, _vp, _vp]),
    "smer_colsum": (_i, [_vp, _i, _ll, _vp, _ll, _i, _vp]),
    "smer_adam_multi": (_i, [_vp, _i, _ll, _i, _f, _f, _f, _f, _f, _vp]),
    "smer_token_accuracy": (_i, [_vp, _ll, _vp, _vp, _i, _vp, _vp, _ll, _i, _vp]),
    "smer_attn_fwd_simt": (_i, [C.POINTER(AttnArgs), _vp]),
    "smer_attn_bwd_simt": (_i, [C.POINTER(AttnArgs), _vp]),
    "smer_attn_fwd_tc": (_i, [C.POINTER(AttnArgs), _vp]),
    "smer_attn_bwd_tc": (_i, [C.POINTER(AttnArgs), _vp]),
    "smer_attn_weights": (_i, [C.POINTER(AttnArgs), _vp, _ll, _vp]),
    "smer_xent_fwd": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _i, _vp, _vp, _ll, _i, _vp]),
    "smer_xent_denominator": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "smer_xent_bwd": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _ll, _ll, _i, _i, _f, _vp, _vp]),
    "smer_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _f, _f, _f, _f, _vp]),
    "smer_adam_step_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _f, _f, _f, _f, _f, _vp]),
    "smer_decode_attn_workspace_bytes": (_ll, [_i, _i, _i, _i]),
    "smer_decode_attn": (_i, [C.POINTER(DecodeAttnArgs), _vp]),
    "smer_decode_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "smer_decode_embed": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "smer_embed_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "smer_set_reserved_sms": (_i, [_i]),
    "smer_set_pdl": (_i, [_i]),
    "smer_decode_linear": (_i, [_vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp, _ll, _vp, _vp, _f, _vp]),
    "smer_sample_masked": (_i, [C.POINTER(SampleArgs), _vp]),
    "smer_cast2d": (_i, [_vp, _i, _ll, _vp, _i, _ll, _ll, _i, _i, _vp]),
    "smer_cast_f32_to_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "smer_kv_len_from_pad": (_i, [_vp, _vp, _i, _i, _vp]),
    "smer_classify_mask": (_i, [_vp, _ll, _i, _vp, _vp]),
}

EXPORTED = tuple(_SIGS)
_lib = None
_lock = threading.Lock()


def lib():
    """Loads the shared library once.  Raises if it is absent (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -m smer_music_generation_b200.build` "
                        "(the SMER B200 path has no CPU fallback)")
                # SMER_B200_LIB: another build of the same library (A/B timing of kernel variants on one box)
                l = C.CDLL(os.environ.get("SMER_B200_LIB") or LIB_PATH)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().smer_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"smer_b200 {what} failed (rc={rc}): {msg}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {d}")


def require_cuda_device() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("smer_b200: no CUDA device; this path runs only on a B200 (no CPU fallback)")
    if not lib().smer_device_ok():
        raise RuntimeError("smer_b200: the kernels are compiled for sm_100a only; this device is not a B200")
