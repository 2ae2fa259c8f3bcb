"""ctypes binding of libvlgae_b200.so (the C ABI declared in include/vlgae_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  If it is missing the
import of any operator fails loudly -- there is deliberately no fallback path.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvlgae_b200.so")

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/vlgae_b200.h declares
PROTOTYPES = {
    "vlgae_version": (c_int, []),
    "vlgae_last_error": (ctypes.c_char_p, []),
    "vlgae_dmv_set_schedule": (c_int, [c_int]),
    "vlgae_dmv_set_linear_max_len": (c_int, [c_int]),
    "vlgae_dmv_set_profile_buffer": (c_int, [c_void_p]),
    "vlgae_dmv_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vlgae_dmv_inside_outside": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_dmv_viterbi": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_dmv_parse": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_dmv_parse_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlgae_dmv_parse_host_async": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlgae_dmv_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                                c_void_p]),
    "vlgae_deptree_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vlgae_deptree": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "vlgae_align_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "vlgae_align_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                   c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "vlgae_align_logits_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                            c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_align_reduce_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "vlgae_align_max_over_factors": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_align_maxima": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_align_diagonal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                     c_void_p]),
    "vlgae_grounding_ce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vlgae_topk_rows": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "vlgae_align_max_over_factors_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vlgae_dmv_scores_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vlgae_dmv_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                 c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlgae_dmv_scores_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_size_t, c_void_p]),
    "vlgae_word_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    "vlgae_word_attention_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlgae_vis_factors": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                  c_void_p]),
    "vlgae_vis_factors_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                           c_void_p, c_void_p, c_void_p]),
    "vlgae_scale_rows": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "vlgae_microbench_mufu": (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
    "vlgae_microbench_fp32": (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class VlgaeError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the CUDA library once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VlgaeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "vlgae_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vlgae_last_error().decode("utf-8", "replace")
        raise VlgaeError(f"{what} failed (code {rc}): {msg}")
