"""ctypes binding of libphylok.so (declared in include/phylok.h).  There is no fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PK_LIB", os.path.join(_HERE, "libphylok.so"))   # PK_LIB: kernel-variant experiments

PK_OK = 0
PK_ERR_INVALID, PK_ERR_PARSE, PK_ERR_CUDA, PK_ERR_NOMEM, PK_ERR_NODEVICE = -1, -2, -3, -4, -5
PK_NORM = {"none": 0, None: 0, "mean": 1, "median": 2}
PK_PREP_ZERO_ROOT, PK_PREP_LADDERIZE, PK_PREP_KAMPHIR = 1, 2, 3


class pk_params(ctypes.Structure):
    _fields_ = [("decay", ctypes.c_double), ("gauss", ctypes.c_double), ("sigma", ctypes.c_double),
                ("precision", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class pk_tree_info(ctypes.Structure):
    _fields_ = [("n_internal", ctypes.c_int32), ("n_tips", ctypes.c_int32), ("n_levels", ctypes.c_int32),
                ("n_classes", ctypes.c_int32), ("n_states", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("height", ctypes.c_double), ("scale", ctypes.c_double)]


_P = ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)
_I32, _I64, _SZ, _D = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_double

# name -> (restype, argtypes); every symbol include/phylok.h declares
SIGNATURES = {
    "pk_abi_version": (ctypes.c_int, []),
    "pk_last_error": (ctypes.c_char_p, []),
    "pk_device_count": (ctypes.c_int, []),
    "pk_tree_from_newick": (ctypes.c_int, [ctypes.c_char_p, _SZ, ctypes.c_int, ctypes.c_int, ctypes.c_int, _PP]),
    "pk_trees_from_newick": (ctypes.c_int, [ctypes.c_char_p, _SZ, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.POINTER(_PP), ctypes.POINTER(_I32)]),
    "pk_trees_from_newick_list": (ctypes.c_int, [_P, _P, _I32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(_PP), ctypes.POINTER(_I32)]),
    "pk_tree_array_free": (None, [_PP]),
    "pk_tree_from_arrays": (ctypes.c_int, [_I32, _P, _P, _P, _I32, _PP]),
    "pk_tree_free": (None, [_P]),
    "pk_tree_set_heights": (ctypes.c_int, [_P, _P, _I32, _D, _D]),
    "pk_tree_get_info": (ctypes.c_int, [_P, ctypes.POINTER(pk_tree_info)]),
    "pk_tree_export": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "pk_tree_node_heights": (ctypes.c_int, [_P, _P]),
    "pk_pair_counts": (ctypes.c_int, [_P, _P, _P]),
    "pk_kcoal": (ctypes.c_int, [_P, _P, ctypes.POINTER(_D)]),
    "pk_kcoal_batch": (ctypes.c_int, [_P, _I32, _P, _P]),
    "pk_ctx_create": (ctypes.c_int, [ctypes.c_int, _PP]),
    "pk_ctx_destroy": (None, [_P]),
    "pk_ctx_set_stream": (ctypes.c_int, [_P, _P]),
    "pk_ctx_sync": (ctypes.c_int, [_P]),
    "pk_ctx_configure": (ctypes.c_int, [_P, _I32, _I32, _I32]),
    "pk_ctx_set_option": (ctypes.c_int, [_P, ctypes.c_char_p, _I32]),
    "pk_ctx_last_stats": (ctypes.c_int, [_P, ctypes.POINTER(_D), ctypes.POINTER(_I64), ctypes.POINTER(_I32), _P]),
    "pk_ctx_nonfinite": (ctypes.c_int, [_P, ctypes.POINTER(_I64), ctypes.c_int]),
    "pk_forest_upload": (ctypes.c_int, [_P, _P, _I32, _I32, _PP]),
    "pk_forest_upload_part": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _I32, _PP]),
    "pk_forest_device_range": (ctypes.c_int, [_P, _I32, _I32, _PP, ctypes.POINTER(_I64)]),
    "pk_forest_free": (None, [_P]),
    "pk_forest_bytes": (ctypes.c_int, [_P, ctypes.POINTER(_I64)]),
    "pk_forest_pair_stats": (ctypes.c_int, [_P, _P, _P, _I64, _P]),
    "pk_forest_gram_stats": (ctypes.c_int, [_P, _I32, _I32, _I32, _P]),
    "pk_kernel_pairs": (ctypes.c_int, [_P, _P, _I32, _P, _I32, _P, _I32, ctypes.POINTER(pk_params), _P]),
    "pk_forest_kernel_pairs": (ctypes.c_int, [_P, _P, _P, _P, _I32, ctypes.POINTER(pk_params), _P]),
    "pk_forest_kernel_pairs_host": (ctypes.c_int, [_P, _P, _P, _P, _I32, ctypes.POINTER(pk_params), _P]),
    "pk_score_newick_batch": (ctypes.c_int, [_P, _P, _P, _P, _I32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(pk_params), ctypes.POINTER(_D), _D, _P, _P, _P, _P, _P,
                                             ctypes.POINTER(_D)]),
    "pk_score_batch": (ctypes.c_int, [_P, _P, _P, _I32, ctypes.POINTER(pk_params), ctypes.POINTER(_D), _P, _P, _P]),
    "pk_gram": (ctypes.c_int, [_P, _P, _I32, ctypes.POINTER(pk_params), ctypes.c_int, _P]),
    "pk_forest_gram_part": (ctypes.c_int, [_P, _P, ctypes.POINTER(pk_params), ctypes.c_int, _I32, _I32, _I32, _P, _I64, _P]),
    "pk_gram_part_pairs": (ctypes.c_int, [_I32, _I32, _I32, _I32, ctypes.POINTER(_I64)]),
    "pk_gram_part_tiles": (ctypes.c_int, [_I32, _I32, _I32, _I32, _P, ctypes.POINTER(_I64)]),
    "pk_forest_gram_part_tiles": (ctypes.c_int, [_P, _I32, _I32, _I32, _P, ctypes.POINTER(_I64)]),
    "pk_host_alloc": (ctypes.c_int, [_P, _SZ, _PP]),
    "pk_host_free": (ctypes.c_int, [_P, _P]),
    "pk_device_alloc": (ctypes.c_int, [_P, _SZ, _PP]),
    "pk_device_free": (ctypes.c_int, [_P, _P]),
    "pk_ipc_export": (ctypes.c_int, [_P, _P, _P]),
    "pk_ipc_open": (ctypes.c_int, [_P, _P, _PP]),
    "pk_ipc_close": (ctypes.c_int, [_P, _P]),
    "pk_memcpy_d2h": (ctypes.c_int, [_P, _P, _P, _SZ]),
    "pk_memcpy_h2d": (ctypes.c_int, [_P, _P, _P, _SZ]),
    "pk_memset_d": (ctypes.c_int, [_P, _P, ctypes.c_int, _SZ]),
}

_lib = None


class PhyloKError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PhyloKError(
                "kamphir_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or kamphir_b200/csrc/build.sh); there is no CPU fallback for the kernel path." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    """Status code -> exception, following the reference's failure modes (SURVEY.md 8b 'Errors')."""
    if rc == PK_OK:
        return
    msg = lib().pk_last_error().decode("utf-8", "replace")
    if rc in (PK_ERR_INVALID, PK_ERR_PARSE):
        raise ValueError(msg)
    if rc == PK_ERR_NOMEM:
        raise MemoryError(msg)
    raise PhyloKError(msg)
