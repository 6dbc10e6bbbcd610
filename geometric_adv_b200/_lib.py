"""ctypes binding of libga_b200.so (C ABI in include/ga_b200.h).

The library is the product; there is no fallback.  If it is missing or cannot be
loaded every op raises -- a silent CPU / eager path would void the parity claims.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libga_b200.so")

GA_OK = 0
GA_ERR_INVALID_ARGUMENT = -1
GA_ERR_UNSUPPORTED = -2
GA_ERR_NO_DEVICE = -3
GA_MODE_CPU_EXACT = 0
GA_MODE_GPU_REF = 1

_i = C.c_int
_p = C.c_void_p
_ll = C.POINTER(C.c_longlong)

# name -> (restype, argtypes); must list every symbol of include/ga_b200.h
SIGNATURES = {
    "ga_version": (_i, []),
    "ga_last_error": (C.c_char_p, []),
    "ga_launch_count": (C.c_longlong, []),
    "ga_last_kernel": (C.c_char_p, []),
    "ga_debug_host_streamed": (C.c_int, []),
    "ga_check_nn_distance": (_i, [_i, _ll, _i, _ll]),
    "ga_check_nn_distance_grad": (_i, [_i, _ll, _i, _ll, _i, _ll, _i, _ll, _i, _ll, _i, _ll]),
    "ga_check_selection_sort": (_i, [_i, _i, _ll]),
    "ga_check_group_point": (_i, [_i, _ll, _i, _ll]),
    "ga_nn_distance_fwd": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p]),
    "ga_nn_distance_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "ga_nn_distance_fwd_ws": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, C.c_size_t, _p]),
    "ga_nn_distance_fwd_host": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i]),
    "ga_nn_distance_bwd": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ga_nn_distance_fwd_bwd": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "ga_nn_distance_bwd_host": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ga_nn_distance_fwd_bwd_host": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i]),
    "ga_chamfer_per_cloud": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "ga_chamfer_loss_terms": (_i, [_i, _i, _i, _p, _p, _p, _p, _p]),
    "ga_chamfer_all_pairs": (_i, [_i, _i, _p, _i, _i, _p, _i, _p]),
    "ga_chamfer_all_pairs_directed": (_i, [_i, _i, _p, _i, _i, _p, _i, _p]),
    "ga_symmetrize_rows": (_i, [_i, _i, _i, _p, _p, _p]),
    "ga_sort_dist_mat": (_i, [_i, _i, _p, _i, _p, _i, _p, _p]),
    "ga_knn": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ga_knn_host": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "ga_selection_sort": (_i, [_i, _i, _i, _i, _p, _p, _p, _p]),
    "ga_group_point": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ga_knn_dists": (_i, [_i, _i, _i, _p, _p, _p]),
    "ga_knn_dists_host": (_i, [_i, _i, _i, _p, _p]),
    "ga_split_by_threshold": (_i, [_i, _i, _p, _p, C.c_float, _p, _p, _p, _p, _p]),
    "ga_set_tuning": (_i, [_i, _i]),
    "ga_probe_fp32_peak": (_i, [_i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p]),
    "ga_probe_launch_floor": (_i, [_i, C.POINTER(C.c_float), _p]),
    "ga_debug_mma_filter": (_i, [_i, _i, _p, _p, _p, _p]),
    "ga_debug_umma_filter": (_i, [_i, _i, _p, _p, _p, _p]),
    "ga_debug_ws_trace": (_i, [_p]),
}


class GaError(RuntimeError):
    """CUDA / library failure."""


_lib = None


def load():
    """Load libga_b200.so once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GaError(
            "geometric_adv_b200: %s is missing. Build it with `make -C geometric_adv_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().ga_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a return code to the exception the reference's framework would raise:
    argument errors -> ValueError (tf.errors.InvalidArgumentError is-a ValueError-like
    user error), everything else -> GaError."""
    if rc == GA_OK:
        return
    msg = last_error()
    if rc == GA_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if rc == GA_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise GaError("libga_b200 failed with code %d: %s" % (rc, msg))


def dims(shape):
    return (C.c_longlong * max(1, len(shape)))(*[int(s) for s in shape])
