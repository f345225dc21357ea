"""ctypes binding of libsevero_b200.so (the C ABI declared in include/severo_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (``make -C severo.jl_b200/csrc``).
There is no CPU fallback: if the library is missing, or no CUDA device is visible, every compute
call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char, c_char_p, c_double, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsevero_b200.so")

SVB_I32, SVB_I64, SVB_F32, SVB_F64 = 0, 1, 2, 3
SVB_OK, SVB_EDIM, SVB_ENOCONV, SVB_ENOMEM, SVB_ENULLSPACE, SVB_EARG = 0, -1, -2, -3, -4, -5
SVB_ECUDA, SVB_ENCCL = -100, -101
NORM_LOGNORMALIZE, NORM_RELATIVECOUNTS = 0, 1
METRIC_EUCLIDEAN, METRIC_COSINE = 0, 1
K_CLASSES = ("spmv_fwd", "spmv_adj", "reorth", "restart", "vector", "comm")


class SeveroB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[svb {code}] {msg}")
        self.code = code


_p64 = POINTER(c_int64)
_pf64 = POINTER(c_double)
_pint = POINTER(c_int)
_h = c_void_p
_ph = POINTER(c_void_p)

# name -> (restype, argtypes). Must list every symbol of include/severo_b200.h (tests check this).
SIGNATURES = {
    "svb_init": (c_int, [c_int]),
    "svb_shutdown": (c_int, []),
    "svb_last_error": (c_char_p, []),
    "svb_version": (c_char_p, []),
    "svb_set_stream": (c_int, [c_void_p]),
    "svb_synchronize": (c_int, []),
    "svb_device_info": (c_int, [_pint, _p64, _pint, _pint]),
    "svb_profile_enable": (c_int, [c_int]),
    "svb_profile_reset": (c_int, []),
    "svb_profile_get": (c_int, [_pf64, _p64, _pf64]),
    "svb_launch_count": (c_int64, []),
    "svb_launch_count_reset": (c_int, []),
    "svb_comm_unique_id": (c_int, [c_void_p]),
    "svb_comm_init": (c_int, [c_int, c_int, c_void_p]),
    "svb_comm_destroy": (c_int, []),
    "svb_comm_info": (c_int, [_pint, _pint]),
    "svb_comm_allreduce_f64": (c_int, [c_void_p, c_int64]),
    "svb_init_devices": (c_int, [c_int, c_void_p]),
    "svb_devices_info": (c_int, [_pint, c_void_p, _pint]),
    "svb_shutdown_devices": (c_int, []),
    "svb_irlba_csc_devices": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int64, c_int64,
                                      c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, _p64, _p64]),
    "svb_pca_counts_devices": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_double, c_double,
                                       c_int64, c_int64, c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, _p64, _p64]),
    "svb_csc_upload": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, _ph]),
    "svb_csc_upload_lognorm_moments": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_double,
                                               c_void_p, c_void_p, _ph]),
    "svb_matrix_free": (c_int, [_h]),
    "svb_matrix_info": (c_int, [_h, _p64, _p64, _p64, _pint]),
    "svb_matrix_download": (c_int, [_h, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "svb_column_subset": (c_int, [_h, c_void_p, c_int64, c_int, _ph]),
    "svb_row_slice": (c_int, [_h, c_int64, c_int64, _ph]),
    "svb_transpose": (c_int, [_h, _ph]),
    "svb_filter_counts": (c_int, [_h, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, _ph]),
    "svb_normalize": (c_int, [_h, c_int, c_double, c_int, _ph]),
    "svb_normalize_libsize": (c_int, [_h, c_void_p, c_int, c_double, c_int, _ph]),
    "svb_row_sums": (c_int, [_h, c_void_p]),
    "svb_mean_var": (c_int, [_h, c_void_p, c_void_p]),
    "svb_welford_carry": (c_int, [_h, c_void_p, c_void_p, c_void_p]),
    "svb_stdvar_clipped": (c_int, [_h, c_void_p, c_void_p, c_double, c_void_p]),
    "svb_scale": (c_int, [_h, c_double, c_int, _ph, c_void_p]),
    "svb_scale_with_moments": (c_int, [_h, c_void_p, c_void_p, c_double, c_int, _ph, c_void_p]),
    "svb_operator_create": (c_int, [_h, c_void_p, c_int, _ph]),
    "svb_operator_create_ex": (c_int, [_h, c_void_p, c_int, c_int, _ph]),
    "svb_operator_create_dense": (c_int, [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int, _ph]),
    "svb_operator_create_counts": (c_int, [_h, c_void_p, c_double, c_void_p, c_void_p, c_double, c_int, c_void_p, _ph]),
    "svb_operator_counts_info": (c_int, [_h, _pint, _p64, _p64, _p64, _p64, _p64]),
    "svb_operator_counts_stream": (c_int, [_h, c_int, c_void_p, c_void_p]),
    "svb_operator_counts_layout": (c_int, [_h, _pint, _pf64, _pint, _pint, _pf64]),
    "svb_operator_free": (c_int, [_h]),
    "svb_operator_info": (c_int, [_h, _p64, _p64, _p64, _pint, _pint, _pint]),
    "svb_mul": (c_int, [_h, c_char, c_double, c_void_p, c_double, c_void_p, c_int64]),
    "svb_mul_device": (c_int, [_h, c_char, c_double, c_void_p, c_double, c_void_p]),
    "svb_irlba": (c_int, [_h, c_int64, c_int64, c_int64, c_int64, c_double, c_double, c_void_p, c_void_p,
                          c_void_p, c_void_p, _p64, _p64]),
    "svb_irlba_solve": (c_int, [_h, c_int64, c_int64, c_int64, c_int64, c_double, c_double, c_void_p, c_void_p,
                                c_void_p, c_void_p, _ph]),
    "svb_spmspv": (c_int, [_h, c_void_p, c_void_p, c_int64, c_int, c_double, c_double, c_void_p]),
    "svb_spgemm_dense": (c_int, [_h, c_int, _h, c_double, c_double, c_void_p, c_int64]),
    "svb_gram": (c_int, [_h, c_void_p]),
    "svb_tssvd": (c_int, [_h, c_int64, c_int64, c_int64, c_double, c_void_p, _ph]),
    "svb_knn": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "svb_knn_result": (c_int, [_h, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "svb_jaccard_index": (c_int, [_h, c_int64, c_double, c_int, _ph]),
    "svb_result_info": (c_int, [_h, _p64, _p64, _p64, _p64, _p64, _pint]),
    "svb_result_download": (c_int, [_h, c_void_p, c_void_p, c_void_p, c_int]),
    "svb_result_free": (c_int, [_h]),
    "svb_synth_counts": (c_int, [c_int64, c_int64, c_int64, c_int64, c_double, c_int64, c_double, c_uint64, _ph]),
    "svb_synth_normal": (c_int, [c_int64, c_uint64, c_void_p]),
}

_lib = None
_initialised_device = None


def load():
    """Load the shared library (no CUDA call yet). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SeveroB200Error(SVB_ECUDA, f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                             "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != SVB_OK:
        msg = load().svb_last_error()
        raise SeveroB200Error(rc, msg.decode() if msg else "unknown error")
    return rc


def init(device=None):
    """svb_init on ``device`` (default: LOCAL_RANK or 0). Idempotent per process."""
    global _initialised_device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device is None:
        check(lib.svb_init(int(device)))
        _initialised_device = int(device)
    elif _initialised_device != int(device):
        raise SeveroB200Error(SVB_EARG, f"library already initialised on device {_initialised_device}")
    return lib


def lib():
    """Initialised library handle (initialises on the default device on first use)."""
    return init(_initialised_device)


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(c_void_p)
