"""ctypes binding of libvldd_b200.so (the C ABI in include/vldd_b200.h).

The product path has no CPU fallback: if the shared library is missing and cannot be built, or a call is made
without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import build as _build

_LIB = None
_P = C.c_void_p

# name -> (restype, argtypes); kept in sync with include/vldd_b200.h (tests/test_abi.py parses the header)
_SIGS = {
    "vldd_version": (C.c_int, []),
    "vldd_last_error": (C.c_char_p, []),
    "vldd_kernel_launch_count": (C.c_ulonglong, []),
    "vldd_flat_sgd_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "vldd_match_loss_scratch_bytes": (C.c_size_t, []),
    "vldd_match_loss_fwd": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P]),
    "vldd_match_loss_bwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P]),
    "vldd_momentum_sgd": (C.c_int, [_P, _P, _P, C.c_float, C.c_float, C.c_int, C.c_int64, _P]),
    "vldd_ranks_from_scores": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "vldd_ranks_cols": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "vldd_rank_best_gt": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "vldd_rank_count": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "vldd_recall_counts": (C.c_int, [_P, C.c_int, _P, _P]),
    "vldd_sim_scores": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    "vldd_topk_fill": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "vldd_sim_rank_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vldd_sim_rank": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vldd_sim_rank_fused_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vldd_sim_rank_fused": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, C.c_int, _P, _P, _P,
                                      C.c_size_t, _P]),
    "vldd_sim_rank_fused_candidates": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, C.c_int, C.c_int, _P, _P,
                                                 _P, C.c_size_t, _P]),
    "vldd_sim_rank_fused_count": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, C.c_int, C.c_int, _P, _P, _P,
                                            C.c_size_t, _P]),
    "vldd_itm_eval_host": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "vldd_proj_head_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vldd_proj_head_forward": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "vldd_contrastive_step_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vldd_contrastive_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P,
                                        C.c_size_t, _P]),
    "vldd_clip_loss": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vldd_infonce_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vldd_infonce_grad": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vldd_infonce_hvp": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vldd_nearest_rows_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vldd_nearest_rows": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "vldd_bench_skinny_gemm_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vldd_bench_skinny_gemm": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P]),
    "vldd_unrolled_match_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vldd_unrolled_match": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_float, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vldd_match_final": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "vldd_outer_update": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, _P, _P, _P, C.c_int64, C.c_float, _P, _P, _P, _P, _P,
                                    C.c_float, C.c_float, C.c_int, C.c_float, _P, _P, _P]),
    "vldd_dropout_masks": (C.c_int, [_P, C.c_int64, C.c_float, _P, C.c_int, _P]),
}


class VlddError(RuntimeError):
    pass


def header_symbols() -> list[str]:
    """Function names declared in include/vldd_b200.h."""
    with open(os.path.join(_build.INCLUDE, "vldd_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vldd_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load (building if needed) the shared library.  Raises if it cannot be produced."""
    global _LIB
    if _LIB is None:
        # always go through build_library: it compares the digest of csrc/ + include/ with the one the .so was built
        # from and rebuilds on a mismatch, so a stale library is never loaded silently (cheap: hashes ~30 files)
        path = _build.build_library(force=False)
        handle = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)          # AttributeError if the .so is stale / missing a symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vldd_last_error().decode("utf-8", "replace")
        raise VlddError(f"{what} failed (code {rc}): {msg}")
