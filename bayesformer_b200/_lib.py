"""ctypes binding of libbayesformer_b200.so (the C ABI in include/bayesformer_b200.h).

There is deliberately no fallback: if the shared library is missing, cannot be
loaded, or a call returns a non-zero status, an exception is raised.  The CUDA
path is the product; nothing here routes work to the CPU.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

ABI_VERSION = 4

OK, ERR_INVALID_ARG, ERR_NOT_DEVICE_PTR, ERR_WORKSPACE, ERR_CUDA, ERR_UNSUPPORTED = range(6)

MODE = {"max": 0, "margin": 1, "entropy": 2}
SCHEME = {"greedy": 0, "sampling": 1, "dropout": 2}
COMPUTE = {"fp32": 0, "fp16": 1, "bf16": 2}

SITE_EMBED = 1 << 0
SITE_LAYER_OUT = 1 << 1
SITE_ATTN_RESID = 1 << 2
SITE_FFN_IN = 1 << 3
SITE_FFN_RESID = 1 << 4
SITE_HEAD_OUT = 1 << 5
SITE_HEAD_IN = 1 << 6    # Head.bayes: per-head masks on the Q / K / V projection inputs (fp32 path only)
SITES_PER_LAYER = 5


class Transformer(C.Structure):
    """struct bfb200_transformer"""

    _fields_ = [
        ("ntoken", C.c_int32),
        ("d_model", C.c_int32),
        ("n_heads", C.c_int32),
        ("d_ff", C.c_int32),
        ("n_layers", C.c_int32),
        ("pe_len", C.c_int32),
        ("p_drop", C.c_float),
        ("site_flags", C.c_uint32),
        ("embedding", C.c_void_p),
        ("pe", C.c_void_p),
        ("w_qkv", C.c_void_p),
        ("w_out", C.c_void_p),
        ("ln1_w", C.c_void_p),
        ("ln1_b", C.c_void_p),
        ("ff1_w", C.c_void_p),
        ("ff1_b", C.c_void_p),
        ("ff2_w", C.c_void_p),
        ("ff2_b", C.c_void_p),
        ("ln2_w", C.c_void_p),
        ("ln2_b", C.c_void_p),
        ("head_w", C.c_void_p),
        ("head_b", C.c_void_p),
        ("ln_eps", C.c_float),
        ("compute", C.c_int32),
        ("fused_image", C.c_void_p),
    ]


class Masks(C.Structure):
    """struct bfb200_masks"""

    _fields_ = [("seed", C.c_uint64), ("bits", C.c_void_p), ("max_len", C.c_int32)]


class McArgs(C.Structure):
    """struct bfb200_mc_args"""

    _fields_ = [
        ("prompts", C.c_void_p),
        ("n_items", C.c_int64),
        ("prompt_len", C.c_int32),
        ("new_tokens", C.c_int32),
        ("n_draws", C.c_int32),
        ("scheme", C.c_int32),
        ("mode", C.c_int32),
        ("temperature", C.c_float),
        ("index_base", C.c_int64),
        ("masks", Masks),
        ("forced_tokens", C.c_void_p),
        ("step_scores", C.c_void_p),
        ("pool_scores", C.c_void_p),
        ("logp_out", C.c_void_p),
        ("tokens_out", C.c_void_p),
    ]


class TrainSegment(C.Structure):
    """struct bfb200_train_segment"""

    _fields_ = [("offset", C.c_int32), ("count", C.c_int32), ("ptr", C.c_void_p)]


class TrainArgs(C.Structure):
    """struct bfb200_train_args"""

    _fields_ = [
        ("ntoken", C.c_int32), ("d_model", C.c_int32), ("n_heads", C.c_int32), ("d_ff", C.c_int32),
        ("n_layers", C.c_int32), ("pe_len", C.c_int32),
        ("batch", C.c_int32), ("seq_len", C.c_int32), ("prompt_len", C.c_int32),
        ("n_segments", C.c_int32), ("step", C.c_int32),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("p_train", C.c_float), ("p_bayes", C.c_float), ("ln_eps", C.c_float),
        ("seed", C.c_uint64),
        ("segments", C.c_void_p), ("tokens", C.c_void_p), ("pe", C.c_void_p),
        ("flat", C.c_void_p), ("grad", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("scratch", C.c_void_p), ("loss_out", C.c_void_p),
    ]


_P = C.c_void_p
# every symbol include/bayesformer_b200.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "bfb200_abi_version": (C.c_int, []),
    "bfb200_last_error": (C.c_char_p, []),
    "bfb200_compute_uncertainty": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "bfb200_acquire": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "bfb200_topk_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "bfb200_topk": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, _P, C.c_size_t, _P]),
    "bfb200_merge_topk": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_size_t, _P]),
    "bfb200_unpack_keys": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "bfb200_mlp_mc_score": (
        C.c_int,
        [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_float, C.c_int32, C.c_uint64, C.c_int64,
         _P, _P, _P, _P, _P, _P],
    ),
    "bfb200_vmlp_mc_score": (
        C.c_int,
        [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P],
    ),
    "bfb200_transformer_workspace_bytes": (C.c_size_t, [C.POINTER(Transformer), C.c_int64, C.c_int32]),
    "bfb200_transformer_forward": (
        C.c_int,
        [C.POINTER(Transformer), _P, C.c_int32, C.c_int64, C.POINTER(Masks), C.c_uint32, C.c_int64, _P, _P,
         C.c_size_t, _P],
    ),
    "bfb200_mc_workspace_bytes": (C.c_size_t, [C.POINTER(Transformer), C.POINTER(McArgs), C.c_int64]),
    "bfb200_mc_score_transformer": (C.c_int, [C.POINTER(Transformer), C.POINTER(McArgs), _P, C.c_size_t, _P]),
    "bfb200_fused_image_bytes": (C.c_size_t, [C.POINTER(Transformer)]),
    "bfb200_compute_path": (C.c_int, [C.POINTER(Transformer)]),
    "bfb200_fused_pack": (C.c_int, [C.POINTER(Transformer), _P, C.c_size_t, _P]),
    "bfb200_linear_tc": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, _P]),
    "bfb200_attention_tc": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "bfb200_step_mean": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    "bfb200_train_state_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "bfb200_train_scratch_floats": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bfb200_train_flat_offset": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bfb200_train_step": (C.c_int, [C.POINTER(TrainArgs), _P]),
    "bfb200_launch_count": (C.c_int64, []),
    "bfb200_reset_launch_count": (None, []),
    "bfb200_profile_begin": (C.c_int, [_P]),
    "bfb200_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_LIB_PATH = Path(__file__).resolve().parent / "libbayesformer_b200.so"
_lib = None


def library_path() -> Path:
    return Path(os.environ.get("BAYESFORMER_B200_LIB", _LIB_PATH))


def load() -> C.CDLL:
    """Load the shared library once; raise RuntimeError when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bayesformer_b200/csrc`). bayesformer_b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(str(path))
    except OSError as exc:  # pragma: no cover - depends on the box
        raise RuntimeError(f"cannot load {path}: {exc}") from exc
    for name, (restype, argtypes) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise RuntimeError(f"{path} does not export {name}; rebuild the library") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bfb200_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.bfb200_abi_version()} != expected {ABI_VERSION}")
    _lib = lib
    return lib


_PY_ERRORS = {
    ERR_INVALID_ARG: ValueError,
    ERR_NOT_DEVICE_PTR: RuntimeError,
    ERR_WORKSPACE: RuntimeError,
    ERR_CUDA: RuntimeError,
    ERR_UNSUPPORTED: NotImplementedError,
}


def check(status: int) -> None:
    """Turn a BFB200_* status into a Python exception (ValueError for bad arguments, as the
    reference raises for unknown `mode` / `compute`; RuntimeError otherwise)."""
    if status == OK:
        return
    msg = load().bfb200_last_error().decode("utf-8", "replace")
    raise _PY_ERRORS.get(status, RuntimeError)(msg)
