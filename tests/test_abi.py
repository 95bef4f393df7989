"""CPU: the C-ABI shared library loads, exports every symbol include/vmae.h declares, the ctypes table mirrors the
header, and argument validation fails loudly (no compute is launched: there is no GPU in this container)."""
import ctypes as C
import os
import re

import pytest

from ssl_vit_video_analytics_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vmae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(?:int|const char\*)\s+(vmae_[a-z0-9_]+)\s*\(", src)


def test_header_declares_expected_entry_points():
    names = header_functions()
    assert len(names) == len(set(names)) and len(names) >= 20
    for must in ("vmae_gemm_bf16", "vmae_attn_fwd", "vmae_attn_bwd", "vmae_layernorm_fwd", "vmae_layernorm_bwd", "vmae_masked_mse_fwd",
                 "vmae_gather_tubelets", "vmae_decoder_input_fwd", "vmae_add_gaussian_noise", "vmae_scale_f32", "vmae_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(os.fspath(_lib.LIB_PATH)) if _lib.LIB_PATH.exists() else _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/vmae.h but not exported"


def test_ctypes_table_mirrors_header():
    assert sorted(_lib.SIGNATURES) == sorted(header_functions())
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vmae.h")).read(), flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else len(params.split(","))
        assert n == len(argtypes), (name, n, len(argtypes))


def test_argument_validation_fails_loudly():
    lib = _lib.load()
    assert lib.vmae_abi_version() == 3
    rc = lib.vmae_layernorm_fwd(None, 1, None, None, None, 0, None, None, 4, 384, 1e-5, None)
    assert rc == -1 and b"null" in lib.vmae_last_error()
    rc = lib.vmae_gemm_bf16(None, 0, 0, None, 0, 0, 128, 64, 64, None, None, 0, None, 0, None, 0, None, 0, 0, 0, 0, 1, None)
    assert rc == -1
    with pytest.raises(RuntimeError, match="failed"):
        _lib.check(rc, "vmae_gemm_bf16")
    rc = lib.vmae_attn_fwd(C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 1, 16, 1, 48, 0.1, 0.0, 0, None)
    assert rc == -2 and b"head_dim" in lib.vmae_last_error()          # unsupported shape: no fallback


def test_ops_refuse_cpu_tensors():
    import torch

    from ssl_vit_video_analytics_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.LinearFn.apply(torch.randn(4, 8), torch.randn(8, 8), None, 0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.masked_mse(torch.randn(1, 4, 192), torch.randn(1, 3, 1, 16, 16), torch.ones(1, 1, 4, dtype=torch.bool), 1, 8)
