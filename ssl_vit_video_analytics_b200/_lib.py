"""ctypes binding of libvmae_sm100a.so (the C ABI in include/vmae.h).

The product path has no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("VMAE_LIB_OVERRIDE") or (_PKG / "lib" / "libvmae_sm100a.so"))   # override: developer A/B builds only

DTYPE_BF16, DTYPE_F32 = 0, 1
MAJOR_K, MAJOR_MN = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_Q8, ACT_AUX_Q8 = 0, 1, 2, 3
DTYPE_U8 = 2

_p, _ll, _i, _f, _ull = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_ulonglong

# name -> argtypes (restype is int unless listed in _RESTYPES). Mirrors include/vmae.h one-to-one.
SIGNATURES = {
    "vmae_last_error": [],
    "vmae_abi_version": [],
    "vmae_device_sm_count": [],
    "vmae_gemm_bf16": [_p, _ll, _i, _p, _ll, _i, _i, _i, _i, _p, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p],
    "vmae_gemm_bf16_noise": [_p, _ll, _p, _ll, _i, _i, _i, _p, _p, _ll, _f, _ull, _ull, _p, _p],
    "vmae_linear_wgrad": [_p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _p, _i, _p],
    "vmae_attn_fwd": [_p, _p, _p, _i, _i, _i, _i, _f, _f, _ull, _p],
    "vmae_attn_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _ull, _p],
    "vmae_attn_dropout_mask": [_p, _i, _i, _i, _f, _ull, _p],
    "vmae_frames_to_clip": [_p, _p, _i, _i, _i, _i, _p, _p, _i, _p],
    "vmae_row_loss_grid": [_ll],
    "vmae_row_loss_fwd": [_p, _i, _p, _i, _ll, _i, _i, _i, _f, _p, _p, _p],
    "vmae_row_loss_bwd": [_p, _i, _p, _i, _ll, _i, _i, _i, _f, _p, _p, _p],
    "vmae_dropout_apply": [_p, _i, _p, _p, _i, _ll, _f, _ull, _p],
    "vmae_dropout_mask": [_p, _ll, _f, _ull, _p],
    "vmae_layernorm_fwd": [_p, _i, _p, _p, _p, _i, _p, _p, _ll, _i, _f, _p],
    "vmae_layernorm_bwd": [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "vmae_masked_mse_fwd": [_p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "vmae_masked_mse_finalize": [_p, _p, _p, _i, _p],
    "vmae_masked_mse_bwd": [_p, _p, _p, _p, _ll, _p],
    "vmae_mask_to_indices": [_p, _p, _i, _i, _i, _p],
    "vmae_gather_tubelets": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vmae_tubelet_embed_fwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vmae_tubelet_embed_wgrad": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vmae_decoder_input_fwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_decoder_input_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_pos_rows_bwd": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "vmae_bn_train_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _f, _f, _i, _p],
    "vmae_bn_apply": [_p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _p],
    "vmae_bn_train_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _p],
    "vmae_dwconv3x3_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_dwconv3x3_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_se_pool": [_p, _p, _p, _i, _i, _i, _f, _p],
    "vmae_se_scale": [_p, _p, _p, _p, _i, _i, _i, _f, _p],
    "vmae_im2col3x3": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_col2im3x3": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_im2col3x3_nchw_f32": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vmae_cast_f32_to_bf16": [_p, _p, _ll, _p],
    "vmae_cast_bf16_to_f32": [_p, _p, _ll, _p],
    "vmae_multi_cast_f32_to_bf16": [_p, _i, _p],
    "vmae_adamw_multi": [_p, _i, _p, _p],
    "vmae_adamw_chunk": [],
    "vmae_gelu_bwd": [_p, _p, _p, _ll, _p],
    "vmae_colsum_accum": [_p, _i, _ll, _p, _ll, _i, _p],
    "vmae_add_gaussian_noise": [_p, _p, _i, _p, _ll, _f, _ull, _ull, _p],
    "vmae_scale_f32": [_p, _p, _f, _i, _ll, _p],
    "vmae_mask_row_indices": [_p, _p, _p, _i, _i, _i, _i, _p],
    "vmae_mask_token_blend_fwd": [_p, _i, _p, _p, _p, _ll, _i, _p],
    "vmae_mask_token_blend_bwd": [_p, _i, _p, _p, _p, _ll, _i, _p],
    "vmae_gather_rows": [_p, _i, _p, _p, _ll, _i, _p],
    "vmae_scatter_rows": [_p, _i, _p, _p, _ll, _ll, _i, _p],
    "vmae_adaptive_avgpool_fwd": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_adaptive_avgpool_bwd": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vmae_meanpool_noise": [_p, _p, _p, _i, _i, _i, _f, _ull, _ull, _p],
    "vmae_softmax_ce_fwd": [_p, _i, _p, _p, _p, _p, _ll, _i, _p],
    "vmae_softmax_ce_bwd": [_p, _i, _p, _p, _p, _p, _ll, _i, _p],
}
_RESTYPES = {"vmae_last_error": C.c_char_p}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if the .so is absent and nvcc exists). Raises RuntimeError when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() and build_if_missing:
        from .build import build

        build()
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: the sm_100a kernels are required (no fallback path)")
    lib = C.CDLL(os.fspath(LIB_PATH))
    missing = [name for name in SIGNATURES if not hasattr(lib, name)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} does not export {missing}; rebuild with python -m ssl_vit_video_analytics_b200.build --force")
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vmae_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
