"""ctypes binding of libmvlt_b200.so (the C ABI declared in include/mvlt_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent, importing/using the ops raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmvlt_b200.so")

_vp, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float

# name -> argtypes (all return int); mirrors include/mvlt_b200.h one to one
PROTOTYPES = {
    "mvlt_init": [],
    "mvlt_abi_version": [],
    "mvlt_gemm_bf16_tc": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_conv2d_nhwc_bf16_tc": [_vp, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_im2col_nhwc": [_vp, _i, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_stem_im2col_nchw": [_vp, _vp, _i, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_resnet_stem_tc": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "mvlt_maxpool_nhwc": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_swin_mlp_fused": [_vp, _ll, _vp, _vp, _f, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp],
    "mvlt_linear_residual_layernorm": [_vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _f, _vp, _ll, _vp, _ll, _i, _i, _i, _vp],
    "mvlt_linear_ln_resident_tiles": [],
    "mvlt_transpose_to_bf16": [_vp, _i, _ll, _vp, _ll, _ll, _i, _vp],
    "mvlt_layernorm_bwd_rows": [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp],
    "mvlt_colsum": [_vp, _i, _ll, _vp, _vp, _ll, _i, _vp],
    "mvlt_gelu_bwd": [_vp, _vp, _vp, _ll, _vp],
    "mvlt_joint_attention_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp],
    "mvlt_swin_ln_qkv": [_vp, _ll, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_swin_block_tail": [_vp, _vp, _ll, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp],
    "mvlt_ln_linear_bf16": [_vp, _ll, _vp, _vp, _f, _vp, _vp, _vp, _ll, _ll, _i, _i, _i, _vp],
    "mvlt_gemm_f32_simt": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _i, _i, _i, _i, _vp],
    "mvlt_layernorm_rows": [_vp, _i, _ll, _vp, _i, _ll, _vp, _vp, _ll, _i, _f, _i, _vp, _ll, _vp],
    "mvlt_patch_embed_ln": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "mvlt_patch_embed_ln_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _f, _vp, _i, _vp],
    "mvlt_layernorm_rows_winmajor": [_vp, _ll, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp],
    "mvlt_window_attention_tc": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "mvlt_joint_attention_tc": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp],
    "mvlt_patch_merge_ln": [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "mvlt_window_attention": [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "mvlt_joint_embed": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mvlt_mlm_ce_fused": [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _ll, _i, _i, _ll, _vp],
    "mvlt_vit_embed": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "mvlt_joint_attention": [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp],
    "mvlt_linear_small": [_vp, _i, _ll, _vp, _vp, _vp, _ll, _i, _i, _vp],
    "mvlt_softmax_rows": [_vp, _vp, _ll, _i, _vp],
    "mvlt_rank_first_positive": [_vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _vp],
    "mvlt_masked_ce_rows": [_vp, _ll, _vp, _vp, _ll, _i, _ll, _vp],
}

_lib = None
_lock = threading.Lock()
_initialised = False
SIZE_QUERIES = ("mvlt_mlm_ce_workspace_bytes", "mvlt_layernorm_bwd_workspace_bytes", "mvlt_colsum_workspace_bytes")


class MvltNativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library and type every prototype.  No compute, no CUDA context needed."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MvltNativeError(
                    f"{LIB_PATH} not found — build it with `python -m medical_vision_langauge_transformer_b200.build` "
                    "(or __graft_entry__.build()).  There is no CPU/PyTorch fallback for the MVLT hot path.")
            lib = C.CDLL(LIB_PATH)
            for name, argtypes in PROTOTYPES.items():
                fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
                fn.argtypes = argtypes
                fn.restype = C.c_int
            for name in SIZE_QUERIES:                   # long long f(long long rows, int cols): workspace sizes
                getattr(lib, name).argtypes = [_ll, _i]
                getattr(lib, name).restype = C.c_longlong
            _lib = lib
    return _lib


def ensure_init() -> C.CDLL:
    """load() + mvlt_init() once (needs a CUDA device; sets opt-in smem sizes and resolves the TMA encoder)."""
    global _initialised
    lib = load()
    if not _initialised:
        with _lock:
            if not _initialised:
                import torch
                if not torch.cuda.is_available():
                    raise MvltNativeError("mvlt_b200 kernels need a CUDA device (sm_100a); none is visible")
                torch.cuda.init()
                torch.zeros(1, device="cuda")  # make sure the primary context is current on this thread
                rc = lib.mvlt_init()
                if rc != 0:
                    raise MvltNativeError(f"mvlt_init failed with code {rc}")
                _initialised = True
    return lib


def check(rc: int, what: str):
    if rc != 0:
        kind = {-1: "invalid argument", -2: "unsupported shape", -3: "driver entry point"}.get(rc, f"cudaError {rc}")
        raise MvltNativeError(f"{what} failed: {kind} (code {rc})")
