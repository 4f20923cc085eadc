"""ctypes binding of libspegnet_b200.so (C-ABI declared in include/spegnet_b200.h).

There is no fallback: if the shared library is missing or the device is not a B200, every entry
point raises.  Two variants of the same sources are built in-tree by ``python -m spegnet_b200.build``
(or ``__graft_entry__.build()``): ``libspegnet_b200_fp16.so`` and ``libspegnet_b200_bf16.so``; they differ
only in the 16-bit storage type of activations / weights (csrc/half16.cuh).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPEGNET_B200_LIBDIR: load the two libraries from another directory (same-box A/B of two builds; development only)
_LIBDIR = os.environ.get("SPEGNET_B200_LIBDIR", _HERE)
LIB_PATHS = {"fp16": os.path.join(_LIBDIR, "libspegnet_b200_fp16.so"),
             "bf16": os.path.join(_LIBDIR, "libspegnet_b200_bf16.so")}
DEFAULT_DTYPE = "fp16"

SPG_OK = 0
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
H16, F32 = 0, 1
BF16 = H16  # legacy alias


class Epilogue(C.Structure):
    """Mirror of ``spg_epilogue_t``."""

    _fields_ = [
        ("bias", C.c_void_p), ("act", C.c_int), ("residual", C.c_void_p), ("res_rows", C.c_int),
        ("out", C.c_void_p), ("out_dtype", C.c_int), ("head_w", C.c_void_p), ("head_b", C.c_float),
        ("head_out", C.c_void_p),
        ("ln_fold_rec", C.c_void_p), ("ln_fold_cw", C.c_void_p), ("ln_cols", C.c_int), ("ln_eps", C.c_float),
        ("ln_emit_rec", C.c_void_p), ("ln_prev_rec", C.c_void_p), ("ln_emit_out", C.c_void_p),
        ("ln_apply_gamma", C.c_void_p), ("ln_apply_beta", C.c_void_p), ("ln_apply_out", C.c_void_p),
    ]


class Launch(C.Structure):
    """Mirror of ``spg_launch_t``: the per-call launch descriptor (stream + SPG_LAUNCH_* flags)."""

    _fields_ = [("stream", C.c_void_p), ("flags", C.c_uint)]


LAUNCH_PDL, LAUNCH_REVERSE = 1, 2

_P, _I, _F, _LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes (restype is always int unless listed in _SPECIAL)
SIGNATURES = {
    "spg_linear_h16": [_P, _P, _I, _I, _I, C.POINTER(Epilogue), _P],
    "spg_conv3x3_h16": [_P, _P, _I, _I, _I, _I, _I, C.POINTER(Epilogue), _P],
    "spg_conv3x3_up2_h16": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "spg_up2_border_gather_h16": [_P, _P, _I, _I, _I, _I, _P],
    "spg_layernorm_f32_h16": [_P, _P, _P, _P, _I, _I, _F, _P],
    "spg_layernorm_matched_f32_h16": [_P, _P, _P, _P, _I, _I, _F, _P],
    "spg_copy_grid_h16": [_P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "spg_patchify_7x7s4": [_P, _P, _I, _I, _P],
    "spg_maxpool2x2_f32": [_P, _P, _I, _I, _I, _I, _P],
    "spg_cast_f32_h16": [_P, _P, _LL, _P],
    "spg_window_attention_h16": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "spg_window_attention_tc_h16": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "spg_upsample_concat_h16": [_P, _I, _I, _I, _P, _I, _I, _I, _P, _I, _I, _I, _P],
    "spg_fusion_combine": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "spg_row_sums_h16": [_P, _P, _I, _I, _I, _I, _P],
    "spg_pooled_mlp": [_P, _I, _I, _P, _P, _I, _P, _P, _I, _I, _P],
    "spg_scale_channels_h16": [_P, _P, _I, _I, _I, _P],
    "spg_easpp_branches": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, C.POINTER(C.c_int), _P],
    "spg_nhwc_h16_to_nchw_f32": [_P, _P, _I, _I, _I, _P],
    "spg_mask_stats_u8": [_P, _P, _P, _P, _I, _I, _I, _P],
    "spg_preprocess_rgb_u8": [_P, _I, _I, _P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, C.c_size_t, _P],
    "spg_resize_bilinear_f32": [_P, _I, _I, _I, _P, _I, _I, _I, _P],
    "spg_sod_gt_prepare_u8": [_P, _I, _I, _I, _P, _P, _P, C.c_size_t, _P],
    "spg_sod_scores_u8": [_P, _P, _P, _P, _I, _I, _I, _P, _P, C.c_size_t, _P],
    "spg_device_check": [],
    "spg_version": [],
    "spg_half_is_fp16": [],
}
_SPECIAL = {
    "spg_last_error": ([], C.c_char_p),
    "spg_launch_count": ([], C.c_longlong),
    "spg_launch_count_reset": ([], None),
    "spg_sod_workspace_bytes": ([_I, _I, _I], C.c_size_t),
    "spg_preprocess_workspace_bytes": ([_I, _I, _I], C.c_size_t),
}
EXPORTS = sorted(list(SIGNATURES) + list(_SPECIAL))

_libs = {}
_lock = threading.Lock()


class SpgError(RuntimeError):
    pass


def dtype_name(torch_dtype) -> str:
    """'fp16' / 'bf16' for the torch 16-bit dtypes (the library variant is chosen by operand dtype)."""
    name = str(torch_dtype)
    if name.endswith("float16") and "bfloat16" not in name:
        return "fp16"
    if name.endswith("bfloat16"):
        return "bf16"
    raise ValueError(f"spegnet_b200 kernels take fp16 or bf16 operands, got {torch_dtype}")


def load(dtype: str = DEFAULT_DTYPE) -> C.CDLL:
    """Load (once) and return the shared library built for `dtype` ('fp16' | 'bf16');
    raises RuntimeError when it has not been built.  There is no fallback."""
    lib = _libs.get(dtype)
    if lib is not None:
        return lib
    with _lock:
        if dtype not in _libs:
            path = LIB_PATHS[dtype]
            if not os.path.exists(path):
                raise RuntimeError(
                    f"{path} not found: the sm_100a extension is not built (run `python -m spegnet_b200.build`). "
                    "spegnet_b200 has no CPU or PyTorch fallback.")
            try:
                import torch  # noqa: F401  (makes torch's bundled libcudart resolvable before dlopen)
            except Exception:  # pragma: no cover
                pass
            lib = C.CDLL(path, mode=C.RTLD_LOCAL)
            for name, args in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = C.c_int
            for name, (args, res) in _SPECIAL.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = res
            if bool(lib.spg_half_is_fp16()) != (dtype == "fp16"):
                raise RuntimeError(f"{path} was built for the other 16-bit type")
            _libs[dtype] = lib
    return _libs[dtype]


def check(rc: int, what: str, dtype: str = DEFAULT_DTYPE) -> None:
    if rc != SPG_OK:
        msg = load(dtype).spg_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise SpgError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    """Kernels launched so far by every loaded library variant."""
    return sum(int(lib.spg_launch_count()) for lib in _libs.values())


def reset_launch_count() -> None:
    for lib in _libs.values():
        lib.spg_launch_count_reset()
