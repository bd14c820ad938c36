"""ctypes binding of libcilrs_b200.so (the C-ABI declared in include/cilrs_b200.h).

There is no fallback: if the library is missing or a call returns a non-zero status, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CILRS_B200_LIB selects an instrumented debug build of the same library (tools/); never a different implementation
LIB_PATH = os.environ.get("CILRS_B200_LIB") or os.path.join(_HERE, "libcilrs_b200.so")
_lib = None


class ConvDesc(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int), ("in_h", ctypes.c_int), ("in_w", ctypes.c_int), ("in_c", ctypes.c_int),
                ("out_c", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int), ("stride", ctypes.c_int),
                ("pad", ctypes.c_int)]


class FlatConvArgs(ctypes.Structure):
    """cilrs_flat_conv_args (include/cilrs_b200.h)"""
    _P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    _fields_ = [("batch", _I), ("H", _I), ("W", _I), ("in_c", _I), ("out_c", _I), ("dgrad", _I), ("flags", _I),
                ("x", _P), ("w", _P), ("y", _P), ("scale", _P), ("bias", _P), ("residual", _P), ("mask", _P), ("mask_bits", _P),
                ("gamma", _P), ("beta", _P), ("running_mean", _P), ("running_var", _P), ("num_batches_tracked", _P),
                ("vec", _P), ("momentum", _F), ("eps", _F), ("update_running", _I),
                ("y1", _P), ("vec1", _P), ("bred1", _P), ("dgamma1", _P), ("dbeta1", _P),
                ("y2", _P), ("vec2", _P), ("bred2", _P), ("dgamma2", _P), ("dbeta2", _P),
                ("partials_ws", _P), ("counter_ws", _P)]


def lib():
    """Load the shared library (building it is `__graft_entry__.build()` / `python build.py`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "cilrs_b200: %s not found - the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'`). There is no CPU fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cilrs_status_string.restype = ctypes.c_char_p
        _lib.cilrs_status_string.argtypes = [ctypes.c_int]
        for name in ("cilrs_conv_packed_weight_bytes", "cilrs_conv_stats_bytes", "cilrs_stem_packed_weight_bytes",
                     "cilrs_model_workspace_bytes", "cilrs_conv_flat_workspace_floats", "cilrs_wgrad_flat_workspace_bytes",
                     "cilrs_jpeg_desc_bytes", "cilrs_jpeg_table_set_bytes", "cilrs_jpeg_plane_bytes", "cilrs_augment_param_bytes"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = ctypes.c_size_t
    return _lib


def _arg(a):
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_longlong(a) if abs(a) > 0x7FFFFFFF else ctypes.c_int(a)
    if isinstance(a, float):
        return ctypes.c_float(a)
    if isinstance(a, ctypes.Structure):
        return ctypes.byref(a)
    return a


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Call an int-returning entry point on the current CUDA stream argument list; raise on non-zero status."""
    fn = getattr(lib(), name)
    status = fn(*[_arg(a) for a in args])
    if status != 0:
        raise RuntimeError("cilrs_b200.%s failed: %s" % (name, lib().cilrs_status_string(status).decode()))


def query(name, *args):
    return getattr(lib(), name)(*[_arg(a) for a in args])
