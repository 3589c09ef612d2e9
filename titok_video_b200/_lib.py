"""ctypes binding of libtitok_b200.so (include/titok_b200.h).

The library is the product: there is no CPU or PyTorch fallback. If the shared object is missing the import of
this module raises, and every call checks the returned ttk_status and raises RuntimeError(ttk_strerror).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TTK_LIB_PATH") or os.path.join(_HERE, "lib", "libtitok_b200.so")  # override: kernel experiments


class TitokB200Error(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise TitokB200Error(
            f"{LIB_PATH} not found: the CUDA extension has not been built. "
            "Run `python titok_video_b200/build.py` (needs nvcc); there is no CPU fallback."
        )
    import torch  # noqa: F401  (loads libcudart.so.12 into the process: the library links the CUDA runtime dynamically)

    return ctypes.CDLL(LIB_PATH)


_lib = _load()

_vp, _i, _i64, _f = c_void_p, c_int, c_int64, c_float
_fp = POINTER(c_float)
_ip = POINTER(c_int32)

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/titok_b200.h exactly
SIGNATURES = {
    "ttk_strerror": [_i],
    "ttk_version": [],
    "ttk_fsq_fwd": [_vp, _vp, _vp, _i64, _i, _i, _fp, _fp, _fp, _fp, _ip, _ip, _vp],
    "ttk_fsq_bwd": [_vp, _vp, _vp, _i64, _i, _i, _fp, _fp, _fp, _fp, _ip, _ip, _vp],
    "ttk_fsq_indices_to_codes": [_vp, _i, _vp, _i, _i64, _i, _fp, _ip, _ip, _vp],
    "ttk_hist_u32": [_vp, _i64, _i, _vp, _vp],
    "ttk_codebook_stats": [_vp, _i, _vp, _vp],
    "ttk_vq_aug_dim": [_i],
    "ttk_vq_aug_rows": [_i, _i],
    "ttk_vq_prepare_codebook": [_vp, _i64, _i, _i, _vp, _i64, _vp],
    "ttk_vq_argmin": [_vp, _i64, _vp, _i64, _i64, _i, _i, _vp, _vp, _vp],
    "ttk_vq_gather_loss": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _vp, _i64, _vp, _vp],
    "ttk_vq_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _f, _f, _vp, _vp, _i64, _vp, _i64, _vp],
    "ttk_gemm_bf16": [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _i64, _vp, _i, _vp],
    "ttk_debug_set_trace": [_vp],
    "ttk_gemm_qkv_rope": [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _i64, _vp, _vp],
    "ttk_gemm_geglu": [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i64, _vp],
    "ttk_gemm_resid_norm256": [_vp, _i64, _vp, _i64, _i, _i, _vp, _i64, _i, _f, _vp, _vp, _vp, _vp, _i64, _vp],
    "ttk_attn_varlen_fwd": [_vp, _i64, _i, _i, _i, _vp, _i, _f, _vp, _i64, _vp, _vp],
    "ttk_rmsnorm_fwd": [_vp, _i64, _vp, _vp, _i64, _i, _i, _vp],
    "ttk_resid_norm": [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i64, _vp],
    "ttk_enc_embed": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp],
    "ttk_dec_embed": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp],
    "ttk_enc_head_fsq": [_vp, _i64, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _fp, _fp, _fp, _fp, _ip, _ip, _vp],
    "ttk_build_plan": [_vp, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ttk_build_plan_bucket": [_vp, _vp, _i64, _i64, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ttk_rope_table_gather": [_vp, _vp, _i, _vp, _i64, _vp],
    "ttk_clip_error": [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp],
    "ttk_normalize_u8": [_vp, _vp, _i64, _vp],
    "ttk_patchify": [_vp, _vp, _i, _i, _i, _i, _vp, _i64, _i64, _vp],
    "ttk_patchify_u8": [_vp, _vp, _i, _i, _i, _i, _vp, _i64, _i64, _vp],
    "ttk_unpatchify": [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _i64, _vp],
    # ---- training path (backward kernels)
    "ttk_attn_varlen_fwd_train": [_vp, _i64, _i, _i, _i, _vp, _i, _f, _vp, _i64, _vp, _vp, _vp, _vp],
    "ttk_attn_bwd_prep": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp],
    "ttk_attn_bwd_dkv": [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _vp, _i64, _vp],
    "ttk_attn_bwd_dq": [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _vp, _i64, _vp],
    "ttk_gemm_wgrad": [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i64, _vp],
    "ttk_rmsnorm_bwd": [_vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i64, _vp],
    "ttk_geglu_fwd": [_vp, _i64, _i, _vp, _i64, _i64, _vp],
    "ttk_geglu_bwd": [_vp, _i64, _i, _vp, _i64, _vp, _i64, _i64, _vp],
    "ttk_gather_rows": [_vp, _i64, _vp, _vp, _i64, _i64, _i, _vp],
    "ttk_scatter_rows": [_vp, _i64, _vp, _vp, _i64, _i64, _i, _vp],
    "ttk_colsum": [_vp, _i64, _i64, _i, _vp, _vp, _vp],
    "ttk_multi_cast": [_vp, _i, _i64, _vp],
    "ttk_head_bwd": [_vp, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "ttk_dec_in_bwd": [_vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "ttk_enc_embed_train": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp],
    "ttk_layers_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ttk_layer_fwd_latent": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ttk_layers_fwd_train": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "ttk_layers_bwd": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ttk_dec_embed_train": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _vp],
}
_RESTYPES = {"ttk_strerror": c_char_p}

for _name, _args in SIGNATURES.items():
    _fn = getattr(_lib, _name)  # AttributeError here == header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, c_int)


def strerror(status: int) -> str:
    return _lib.ttk_strerror(int(status)).decode()


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise TitokB200Error(f"{what or 'libtitok_b200'} failed: {strerror(status)} (status {status})")


def fn(name: str):
    return getattr(_lib, name)


LAUNCHES = 0  # kernels launched through call() since import (an entry point launches one kernel unless `launches` says otherwise)
_PROFILER = None  # optional object with begin(name) / end(name, token), see bench.py


def set_profiler(p) -> None:
    global _PROFILER
    _PROFILER = p


class LayersDesc(ctypes.Structure):
    """ttk_layers_desc of include/titok_b200.h."""
    _fields_ = [("M", c_int32), ("width", c_int32), ("gqa", c_int32), ("inner", c_int32), ("n_layers", c_int32),
                ("n_attn_work", c_int32), ("n_dkv_work", c_int32), ("n_dq_work", c_int32),
                ("alpha", c_float), ("softmax_scale", c_float),
                ("rope", c_void_p), ("attn_work", c_void_p), ("dkv_work", c_void_p), ("dq_work", c_void_p),
                ("weights", c_void_p), ("k_norm2", c_void_p)]


def profiling() -> bool:
    """True while a per-kernel profiler is installed: launch sequences then go kernel by kernel through call()."""
    return _PROFILER is not None


def call(name: str, *args, launches: int = 1, label: str = None) -> None:
    """Invoke an int-returning kernel entry point and raise on a non-zero status. `launches`: kernels the entry point
    enqueues (1 for the kernel entry points -- 2 for the attention forward without key norms --, more for the native sequencers).
    `label`: the name a per-kernel profiler files this launch under (default: the entry point's)."""
    global LAUNCHES
    LAUNCHES += launches
    if _PROFILER is None:
        check(getattr(_lib, name)(*args), name)
    else:
        tok = _PROFILER.begin(label or name)
        check(getattr(_lib, name)(*args), name)
        _PROFILER.end(label or name, tok)


def version() -> int:
    return _lib.ttk_version()


def float_array(values) -> "ctypes.Array":
    vals = [float(v) for v in values]
    return (c_float * len(vals))(*vals)


def int_array(values) -> "ctypes.Array":
    vals = [int(v) for v in values]
    return (c_int32 * len(vals))(*vals)
