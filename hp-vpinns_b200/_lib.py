"""ctypes binding of libhpv.so (include/hpv.h).  The library is the product's only compute path: if it is
missing, or there is no B200, every call raises -- there is no CPU fallback here."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# HPV_LIB selects another build of the same library (tuning experiments: build.py --lib PATH)
LIB_PATH = os.environ.get("HPV_LIB") or os.path.join(HERE, "libhpv.so")

c_int, c_double, c_void_p, c_float = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_float
P_int, P_double, P_float = ctypes.POINTER(c_int), ctypes.POINTER(c_double), ctypes.POINTER(c_float)

# name -> (restype, argtypes): every symbol include/hpv.h declares
SIGNATURES = {
    "hpv_abi_version": (c_int, []),
    "hpv_device_count": (c_int, []),
    "hpv_create": (c_int, [ctypes.POINTER(c_void_p), c_int]),
    "hpv_destroy": (None, [c_void_p]),
    "hpv_last_error": (ctypes.c_char_p, [c_void_p]),
    "hpv_set_stream": (c_int, [c_void_p, c_void_p]),
    "hpv_sync": (c_int, [c_void_p]),
    "hpv_set_network": (c_int, [c_void_p, c_int, P_int, c_int, c_int]),
    "hpv_num_params": (c_int, [c_void_p]),
    "hpv_set_params": (c_int, [c_void_p, P_double, c_int, c_double]),
    "hpv_get_params": (c_int, [c_void_p, P_double, c_int, P_double]),
    "hpv_set_quadrature": (c_int, [c_void_p, c_int, P_double, P_double]),
    "hpv_set_test_tables": (c_int, [c_void_p, c_int, P_double, P_double, P_double, P_double]),
    "hpv_set_form": (c_int, [c_void_p, c_int, c_int, c_double]),
    "hpv_set_elements": (c_int, [c_void_p, c_int, P_double, P_double, P_int, c_int, c_int, P_double]),
    "hpv_update_rhs_f32": (c_int, [c_void_p, c_void_p]),
    "hpv_project_field": (c_int, [c_void_p, P_double, c_int, c_int, c_double, c_int, c_int, P_double]),
    "hpv_varloss_forward": (c_int, [c_void_p, P_double, c_void_p, P_double]),
    "hpv_forward_async": (c_int, [c_void_p]),
    "hpv_varloss_backward": (c_int, [c_void_p, P_double, c_int, P_double]),
    "hpv_net_u": (c_int, [c_void_p, c_int, P_double, P_double, P_double, P_double]),
    "hpv_set_point_loss": (c_int, [c_void_p, c_int, c_int, P_double, P_double, P_double, P_double, c_double]),
    "hpv_point_loss_forward": (c_int, [c_void_p, c_int, P_double, P_double]),
    "hpv_configure_training": (c_int, [c_void_p, c_double, ctypes.c_uint, c_int, c_double, c_double, c_double, c_double]),
    "hpv_loss_and_grad": (c_int, [c_void_p]),
    "hpv_reduce_buffer": (c_int, [c_void_p, ctypes.POINTER(c_void_p), P_int]),
    "hpv_adam_step": (c_int, [c_void_p]),
    "hpv_peer_export": (c_int, [c_void_p, c_int, ctypes.c_char_p]),
    "hpv_peer_connect": (c_int, [c_void_p, c_int, c_int, ctypes.c_char_p]),
    "hpv_read_losses": (c_int, [c_void_p, P_double, c_int]),
    "hpv_read_grad": (c_int, [c_void_p, P_double, c_int, P_double]),
    "hpv_read_losses_and_grad": (c_int, [c_void_p, P_double, c_int, P_double, c_int, P_double]),
    "hpv_reset_optimizer": (c_int, [c_void_p]),
    "hpv_train_steps": (c_int, [c_void_p, c_int, P_double]),
    "hpv_launch_count": (ctypes.c_longlong, [c_void_p]),
    "hpv_kernel_info": (c_int, [c_void_p, P_int, c_int]),
    "hpv_probe_fp32_peak": (c_int, [c_void_p, c_int, P_double]),
    "hpv_time_kernel": (c_int, [c_void_p, c_int, c_int, P_double]),
}

_lib = None


class HpvError(RuntimeError):
    pass


def load():
    """Load libhpv.so and declare the prototypes.  Raises HpvError when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HpvError("libhpv.so is not built: run `python hp-vpinns_b200/build.py` (or __graft_entry__.build())")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(P_double)


def iptr(a):
    return None if a is None else a.ctypes.data_as(P_int)


def f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
