"""ctypes binding of libherald_b200.so — the counterpart of python/hetu/_base.py:66-95.

The reference loads ``build/lib/libc_runtime_api.so`` as ``_LIB`` and asserts ``ret == 0`` on
plumbing calls; here every call (op calls included) is checked and a failure raises with the
library's message.  A missing library is an ImportError: there is no fallback path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libherald_b200.so")


class HeraldError(RuntimeError):
    pass


def _load_lib():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "herald_b200: %s is missing — build it with `make lib` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, ctypes.RTLD_GLOBAL)
    lib.HBGetLastError.restype = ctypes.c_char_p
    lib.HBVersion.restype = ctypes.c_char_p
    lib.HBKernelLaunchCount.restype = ctypes.c_uint64
    return lib


_LIB = _load_lib()


def check_call(ret):
    """Raise when a C API call failed (the reference asserts ret == 0, _base.py:84-95)."""
    if ret != 0:
        raise HeraldError(_LIB.HBGetLastError().decode("utf-8", "replace"))


def c_array(ctype, values):
    return (ctype * len(values))(*values)


def kernel_launch_count():
    """Kernels launched by this library in this process."""
    return int(_LIB.HBKernelLaunchCount())


def version():
    return _LIB.HBVersion().decode()


def set_hot_threshold(rows):
    """Ids occurring more than `rows` times in a batch take the column-split reduce path."""
    check_call(_LIB.HBSetHotThreshold(ctypes.c_uint(int(rows))))
