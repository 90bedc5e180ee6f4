"""ctypes binding of libtae_b200.so (the C ABI declared in include/tae_b200.h).

The library is built in-tree by tae_b200/build.py.  There is no CPU fallback: if the shared library is missing
or the device is not an sm_100 part, the ops raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libtae_b200.so"

TAE_OK = 0
EPI_BF16, EPI_BF16_GELU, EPI_F32_RESID, EPI_F32_ACC, EPI_BF16_DGELU, EPI_BF16_ROWDOT = range(6)


class TaeError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of `tae_gemm_args` (include/tae_b200.h)."""

    _fields_ = [
        ("A", C.c_void_p),
        ("B", C.c_void_p),
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("lda", C.c_int32),
        ("ldb", C.c_int32),
        ("a_mn_major", C.c_int32),
        ("b_mn_major", C.c_int32),
        ("epilogue", C.c_int32),
        ("out", C.c_void_p),
        ("ldo", C.c_int32),
        ("out2", C.c_void_p),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p),
        ("ldr", C.c_int32),
        ("resid_rows", C.c_int32),
        ("aux", C.c_void_p),
        ("ldaux", C.c_int32),
        ("beta", C.c_int32),
        ("splits", C.c_int32),
        ("colsum_partials", C.c_void_p),
        ("rowdot", C.c_void_p),
        ("rowdot_tokens", C.c_int32),
    ]


_vp, _i32, _f32, _sz = C.c_void_p, C.c_int32, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/tae_b200.h declares
PROTOTYPES = {
    "tae_version": (C.c_int, []),
    "tae_build_fingerprint": (C.c_char_p, []),
    "tae_last_error_string": (C.c_char_p, []),
    "tae_device_check": (C.c_int, []),
    "tae_num_sms": (C.c_int, []),
    "tae_launch_count": (C.c_uint64, []),
    "tae_set_dynamic_scheduling": (C.c_int, [C.c_int]),
    "tae_gemm": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "tae_layernorm_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "tae_layernorm_bwd_num_partials": (C.c_int, [_i32, _i32]),
    "tae_layernorm_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "tae_layernorm_bwd_finalize": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp]),
    "tae_attention_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_attention_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_attention_bwd_delta": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_im2col_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "tae_patchify": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_unpatchify": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_mse_loss": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "tae_colsum_workspace_floats": (C.c_size_t, [_i32, _i32]),
    "tae_colsum_bf16": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "tae_colsum_f32": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _vp]),
    "tae_batch_sum_f32": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "tae_adamw_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _sz, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _vp, _vp, _vp]),
    "tae_adamw_hyper": (C.c_int, [_f32, _f32, _f32, _f32, _f32, _i32, _f32, C.POINTER(C.c_float)]),
    "tae_adamw_step_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "tae_cast_f32_to_bf16": (C.c_int, [_vp, _vp, _sz, _vp]),
    "tae_grad_stats": (C.c_int, [_vp, _sz, _vp, _vp, _vp]),
    "tae_patchify_c": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tae_unpatchify_c": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tae_token_mean_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "tae_token_mean_bwd_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    # fp32 ("no autocast") mode
    "tae_split3_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "tae_bias_act_f32": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp]),
    "tae_gelu_bwd_f32": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "tae_add_f32": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "tae_layernorm_fwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "tae_layernorm_bwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "tae_attention_fwd_f32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_attention_bwd_f32_workspace_floats": (C.c_size_t, [_i32, _i32, _i32]),
    "tae_attention_bwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tae_im2col_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "tae_mse_loss_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed and possible) the shared library and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("TAE_B200_LIB")  # developer A/B builds (tae_b200.build --variant); never a fallback
    if override:
        path = Path(override).resolve(strict=True)
    else:
        if build_if_missing:
            # no-op when the in-tree library matches the sources' fingerprint; rebuilds (nvcc) when csrc/ changed
            from . import build as _build

            _build.build()
        if not LIB_PATH.exists():
            raise TaeError(f"{LIB_PATH} is missing: build it with `python -m tae_b200.build` (no CPU fallback exists)")
        path = LIB_PATH
    lib = C.CDLL(os.fspath(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise TaeError(f"{path} does not export {name}: it was built from other sources — rebuild it "
                           "(`python -m tae_b200.build --force`, variants: tools/build_variants.sh)")
        fn.restype = res
        fn.argtypes = args
    # a library built from other sources may disagree about struct layouts and prototypes: refuse it
    from . import build as _build

    built_from, have = lib.tae_build_fingerprint().decode(), _build.fingerprint()
    if built_from != have and os.environ.get("TAE_ALLOW_STALE_LIB") != "1":
        raise TaeError(f"{path} was built from sources with fingerprint {built_from[:12]}, the checkout has {have[:12]}: "
                       "rebuild it (`python -m tae_b200.build --force`), or set TAE_ALLOW_STALE_LIB=1 to load it anyway")
    _lib = lib
    return lib


def last_error() -> str:
    return load().tae_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != TAE_OK:
        raise TaeError(f"{what or 'tae call'} failed with code {rc}: {last_error()}")


_device_ok = None


def require_device() -> None:
    """Fail loudly unless a B200-class (sm_100) device is current."""
    global _device_ok
    if _device_ok:
        return
    import torch

    if not torch.cuda.is_available():
        raise TaeError("tae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    check(load().tae_device_check(), "tae_device_check")
    _device_ok = True
